"""Scalar restatement of csrc/tb_wcs.cuh (world -> pixel) for the HOST-SIDE set-up of
``toast_b200.wcs.create_wcs`` (CRPIX and the image shape need a few world2pix evaluations before
any sample is projected).  Same operations in the same order as the device code; the per-sample
work itself only runs on the GPU (``tb_pixels_wcs``)."""

import math

from .wcs import D2R, R2D, _acosd, _atan2d, _sincosd


def _asind(v):
    if v <= -1.0 and v + 1.0 > -1e-10:
        return -90.0
    if v == 0.0:
        return 0.0
    if v >= 1.0 and v - 1.0 < 1e-10:
        return 90.0
    return math.asin(v) * R2D


def _tand(a):
    r = math.fmod(a, 360.0)
    if r == 0.0 or abs(r) == 180.0:
        return 0.0
    if r in (45.0, 225.0):
        return 1.0
    if r in (-135.0, -315.0):
        return -1.0
    return math.tan(a * D2R)


def _wrap(phi):
    if phi > 180.0:
        return phi - 360.0
    if phi < -180.0:
        return phi + 360.0
    return phi


def sph_s2x(eul, lng, lat):
    if eul[4] == 0.0:
        if eul[1] == 0.0:
            dphi = math.fmod(eul[2] - 180.0 - eul[0], 360.0)
            return _wrap(math.fmod(lng + dphi, 360.0)), lat
        dphi = math.fmod(eul[2] + eul[0], 360.0)
        return _wrap(math.fmod(dphi - lng, 360.0)), -lat
    dlng = lng - eul[0]
    sinlng, coslng = _sincosd(dlng)
    sinlat, coslat = _sincosd(lat)
    coslat3, coslat4 = coslat * eul[3], coslat * eul[4]
    sinlat3, sinlat4 = sinlat * eul[3], sinlat * eul[4]
    x = sinlat4 - coslat3 * coslng
    if abs(x) < 1.0e-5:
        x = -_sincosd(lat + eul[1])[1] + coslat3 * (1.0 - coslng)
    y = -coslat * sinlng
    if x != 0.0 or y != 0.0:
        dphi = _atan2d(y, x)
    else:
        dphi = dlng - 180.0 if eul[1] < 90.0 else -dlng
    phi = _wrap(math.fmod(eul[2] + dphi, 360.0))
    if math.fmod(dlng, 180.0) == 0.0:
        theta = lat + coslng * eul[1]
        if theta > 90.0:
            theta = 180.0 - theta
        if theta < -90.0:
            theta = -180.0 - theta
    else:
        z = sinlat3 + coslat4 * coslng
        if abs(z) > 0.99:
            a = abs(_acosd(math.sqrt(x * x + y * y)))
            theta = -a if z < 0.0 else a
        else:
            theta = _asind(z)
    return phi, theta


def prj_s2x(proj, lam, phi, theta):
    if proj == "CAR":
        return phi, theta
    if proj == "CEA":
        return phi, (R2D / lam) * _sincosd(theta)[0]
    if proj == "MER":
        if theta <= -90.0:
            return None
        return phi, R2D * math.log(_tand((theta + 90.0) / 2.0))
    if proj == "SFL":
        return phi * _sincosd(theta)[1], theta
    if proj == "TAN":
        s, c = _sincosd(theta)
        if s <= 0.0:
            return None
        r = R2D * c / s
    else:  # ZEA
        r = 2.0 * R2D * _sincosd((90.0 - theta) / 2.0)[0]
    sp, cp = _sincosd(phi)
    return r * sp, -r * cp


def world2pix_scalar(wcs, lng, lat):
    phi, theta = sph_s2x(wcs.euler, lng, lat)
    xy = prj_s2x(wcs.proj, wcs.cea_lambda, phi, theta)
    if xy is None:
        return float("nan"), float("nan")
    return ((xy[0] / wcs.cdelt[0] + wcs.crpix[0]) - 1.0,
            (xy[1] / wcs.cdelt[1] + wcs.crpix[1]) - 1.0)
