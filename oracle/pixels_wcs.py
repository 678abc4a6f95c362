"""oracle/pixels_wcs.py -- TEST INFRASTRUCTURE ONLY (never imported by toast_b200).

numpy restatement of ``PixelsWCS`` (/root/reference/src/toast/ops/pixels_wcs.py:39-662) for the
hot-path parity tests:

* ``quat_to_lonlat_deg``   pointing_utils.py:16-67 (center_offset_lonlat, no moving centre) on top
                           of qa_to_iso (_libtoast/math_qarray.cpp:694-783) and np.degrees
* ``create_wcs``           pixels_wcs.py:190-345
* ``world2pix`` / ``pix2world``  what ``astropy.wcs.WCS.wcs_world2pix / wcs_pix2world(..., 0)`` do
                           for the six projections the operator offers, through WCSLIB

PARITY STATUS: the TOAST half (quaternion -> lon / lat, the operator logic, around(), the
``>= n_pix`` rule) follows the reference source line by line.  The WCSLIB half is a third-party
dependency (astropy's bundled WCSLIB; astropy is not installed in this image and is not vendored
under /root/reference): it is restated from its published algorithm -- Calabretta & Greisen 2002,
A&A 395, 1077 (Paper II) sections 2.2-2.4 (celestial <-> native rotation, default LONPOLE /
LATPOLE), 5.1.3 TAN, 5.1.7 ZEA, 5.2.2 CEA, 5.2.3 CAR, 5.2.4 MER, 5.3.1 SFL -- and wcstrig.c's
degree trigonometry.  It is **unpinned against WCSLIB itself**; it is pinned against (i) closed-form
projection formulas, (ii) the forward / inverse round trip, and (iii) the reference's own test
(tests/ops_pointing_wcs.py:45-78, 165-215: a boresight aimed at every pixel centre must hit every
pixel exactly once).
"""

import numpy as np

D2R = np.pi / 180.0
R2D = 180.0 / np.pi
PROJECTIONS = ("CAR", "CEA", "MER", "SFL", "TAN", "ZEA")


# ---- wcstrig.c ------------------------------------------------------------------------------
def sincosd(a):
    a = np.asarray(a, dtype=np.float64)
    s, c = np.sin(a * D2R), np.cos(a * D2R)
    ex = np.fmod(a, 90.0) == 0.0
    i = np.abs(np.floor(a / 90.0 + 0.5).astype(np.int64)) % 4
    s = np.where(ex, np.choose(i, [0.0, np.where(a > 0, 1.0, -1.0), 0.0,
                                   np.where(a > 0, -1.0, 1.0)]), s)
    c = np.where(ex, np.choose(i, [1.0, 0.0, -1.0, 0.0]), c)
    return s, c


def atan2d(y, x):
    y, x = np.broadcast_arrays(np.asarray(y, dtype=np.float64), np.asarray(x, dtype=np.float64))
    r = np.arctan2(y, x) * R2D
    r = np.where((y == 0.0) & (x >= 0.0), 0.0, r)
    r = np.where((y == 0.0) & (x < 0.0), 180.0, r)
    r = np.where((y != 0.0) & (x == 0.0) & (y > 0.0), 90.0, r)
    r = np.where((y != 0.0) & (x == 0.0) & (y < 0.0), -90.0, r)
    return r


def asind(v):
    v = np.asarray(v, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        r = np.arcsin(np.clip(v, -1.0, 1.0)) * R2D
    r = np.where(v == 0.0, 0.0, r)
    return r


def acosd(v):
    v = np.asarray(v, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        r = np.arccos(np.clip(v, -1.0, 1.0)) * R2D
    r = np.where(v == 0.0, 90.0, r)
    return r


def tand(a):
    a = np.asarray(a, dtype=np.float64)
    r = np.tan(a * D2R)
    m = np.fmod(a, 360.0)
    r = np.where((m == 0.0) | (np.abs(m) == 180.0), 0.0, r)
    r = np.where((m == 45.0) | (m == 225.0), 1.0, r)
    r = np.where((m == -135.0) | (m == -315.0), -1.0, r)
    return r


# ---- celset: Euler angles for the default LONPOLE / LATPOLE ----------------------------------
def celestial_euler(proj, crval):
    lng0, lat0 = float(crval[0]), float(crval[1])
    zen = proj in ("TAN", "ZEA")
    theta0 = 90.0 if zen else 0.0
    phip = 0.0 if lat0 >= theta0 else 180.0
    if zen:
        lngp, latp = lng0, lat0
    else:
        slat0, clat0 = (float(v) for v in sincosd(lat0))
        sphip, cphip = (float(v) for v in sincosd(phip))
        # theta0 = 0: x = cos(phi_p), y = 0, z = 1; latp = u +- v with u = atan2d(0, x)
        u = float(atan2d(0.0, cphip))
        v = float(acosd(slat0))
        cands = []
        for lp in (u + v, u - v):
            if lp > 180.0:
                lp -= 360.0
            elif lp < -180.0:
                lp += 360.0
            cands.append(lp)
        ok = [lp for lp in cands if abs(lp) < 90.0 + 5e-9]
        latp = min(ok, key=lambda lp: abs(90.0 - lp))      # closest to LATPOLE = 90
        latp = max(-90.0, min(90.0, latp))
        z = np.cos(latp * D2R) * clat0
        if abs(z) < 5e-9:
            lngp = lng0 + phip - 180.0 if latp > 0 else lng0 - phip
        else:
            x = (0.0 - np.sin(latp * D2R) * slat0) / z
            y = sphip * 1.0 / clat0
            lngp = lng0 - float(atan2d(y, x))
        if lng0 >= 0.0:
            if lngp < 0.0:
                lngp += 360.0
            elif lngp > 360.0:
                lngp -= 360.0
        else:
            if lngp > 0.0:
                lngp -= 360.0
            elif lngp < -360.0:
                lngp += 360.0
    e1 = 90.0 - latp
    s, c = (float(v) for v in sincosd(e1))
    return np.array([lngp, e1, phip, c, s])


class Wcs:
    def __init__(self, proj, crval, cdelt, is_azimuth=False):
        assert proj in PROJECTIONS
        self.proj = proj
        self.crval = np.array(crval, dtype=np.float64)
        self.cdelt = np.array(cdelt, dtype=np.float64)
        self.crpix = np.zeros(2)
        self.lam = np.cos(np.deg2rad(self.crval[1])) ** 2 if proj == "CEA" else 1.0
        self.euler = celestial_euler(proj, self.crval)
        self.shape = (0, 0)
        self.is_azimuth = is_azimuth


def _wrap180(phi):
    phi = np.where(phi > 180.0, phi - 360.0, phi)
    return np.where(phi < -180.0, phi + 360.0, phi)


def sph_s2x(eul, lng, lat):
    """celestial -> native (sphs2x)."""
    lng, lat = np.asarray(lng, dtype=np.float64), np.asarray(lat, dtype=np.float64)
    if eul[4] == 0.0:
        if eul[1] == 0.0:
            dphi = np.fmod(eul[2] - 180.0 - eul[0], 360.0)
            return _wrap180(np.fmod(lng + dphi, 360.0)), lat.copy()
        dphi = np.fmod(eul[2] + eul[0], 360.0)
        return _wrap180(np.fmod(dphi - lng, 360.0)), -lat
    dlng = lng - eul[0]
    sinlng, coslng = sincosd(dlng)
    sinlat, coslat = sincosd(lat)
    coslat3, coslat4 = coslat * eul[3], coslat * eul[4]
    sinlat3, sinlat4 = sinlat * eul[3], sinlat * eul[4]
    x = sinlat4 - coslat3 * coslng
    x = np.where(np.abs(x) < 1.0e-5, -sincosd(lat + eul[1])[1] + coslat3 * (1.0 - coslng), x)
    y = -coslat * sinlng
    dphi = np.where((x != 0.0) | (y != 0.0), atan2d(y, x),
                    dlng - 180.0 if eul[1] < 90.0 else -dlng)
    phi = _wrap180(np.fmod(eul[2] + dphi, 360.0))
    z = sinlat3 + coslat4 * coslng
    a = np.abs(acosd(np.sqrt(x * x + y * y)))
    theta = np.where(np.abs(z) > 0.99, np.where(z < 0.0, -a, a), asind(z))
    t2 = lat + coslng * eul[1]
    t2 = np.where(t2 > 90.0, 180.0 - t2, t2)
    t2 = np.where(t2 < -90.0, -180.0 - t2, t2)
    theta = np.where(np.fmod(dlng, 180.0) == 0.0, t2, theta)
    return phi, theta


def sph_x2s(eul, phi, theta):
    """native -> celestial (sphx2s), the general rotation (used by the round-trip KAT only)."""
    phi, theta = np.asarray(phi, dtype=np.float64), np.asarray(theta, dtype=np.float64)
    dphi = phi - eul[2]
    sinphi, cosphi = sincosd(dphi)
    sinthe, costhe = sincosd(theta)
    x = sinthe * eul[4] - costhe * eul[3] * cosphi
    y = -costhe * sinphi
    dlng = np.where((x != 0.0) | (y != 0.0), atan2d(y, x), dphi + 180.0)
    lng = eul[0] + dlng
    z = sinthe * eul[3] + costhe * eul[4] * cosphi
    a = np.abs(acosd(np.sqrt(x * x + y * y)))
    lat = np.where(np.abs(z) > 0.99, np.where(z < 0.0, -a, a), asind(z))
    return lng, lat


def prj_s2x(w, phi, theta):
    p = w.proj
    with np.errstate(divide="ignore", invalid="ignore"):
        if p == "CAR":
            return phi, theta, np.ones_like(phi, dtype=bool)
        if p == "CEA":
            return phi, (R2D / w.lam) * sincosd(theta)[0], np.ones_like(phi, dtype=bool)
        if p == "MER":
            ok = theta > -90.0
            return phi, R2D * np.log(tand((theta + 90.0) / 2.0)), ok
        if p == "SFL":
            return phi * sincosd(theta)[1], theta, np.ones_like(phi, dtype=bool)
        sp, cp = sincosd(phi)
        if p == "TAN":
            s, c = sincosd(theta)
            ok = s > 0.0
            r = R2D * c / np.where(ok, s, 1.0)
        else:
            r = 2.0 * R2D * sincosd((90.0 - theta) / 2.0)[0]
            ok = np.ones_like(phi, dtype=bool)
        return r * sp, -r * cp, ok


def prj_x2s(w, x, y):
    p = w.proj
    if p == "CAR":
        return x, y
    if p == "CEA":
        return x, asind(y * w.lam / R2D)
    if p == "MER":
        return x, 2.0 * np.arctan(np.exp(y / R2D)) * R2D - 90.0
    if p == "SFL":
        c = sincosd(y)[1]
        return x / c, y
    r = np.sqrt(x * x + y * y)
    phi = np.where(r == 0.0, 0.0, atan2d(x, -y))
    if p == "TAN":
        return phi, atan2d(R2D, r)
    return phi, 90.0 - 2.0 * asind(r / (2.0 * R2D))


def world2pix(w, lng, lat):
    """wcs_world2pix([...], 0): fractional (col, row) and a validity mask."""
    phi, theta = sph_s2x(w.euler, lng, lat)
    x, y, ok = prj_s2x(w, phi, theta)
    col = (x / w.cdelt[0] + w.crpix[0]) - 1.0
    row = (y / w.cdelt[1] + w.crpix[1]) - 1.0
    return col, row, ok


def pix2world(w, col, row):
    x = (np.asarray(col, dtype=np.float64) + 1.0 - w.crpix[0]) * w.cdelt[0]
    y = (np.asarray(row, dtype=np.float64) + 1.0 - w.crpix[1]) * w.cdelt[1]
    phi, theta = prj_x2s(w, x, y)
    return sph_x2s(w.euler, phi, theta)


def create_wcs(proj="CAR", center_deg=None, bounds_deg=None, res_deg=None, dims=None,
               is_azimuth=False):
    """pixels_wcs.py:190-345."""
    if center_deg is not None:
        crval = np.array(center_deg, dtype=np.float64)
    else:
        lon_min, lon_max, lat_min, lat_max = bounds_deg
        crval = np.array([0.5 * (lon_min + lon_max), 0.5 * (lat_min + lat_max)])
    if center_deg is not None or res_deg is not None:
        cdelt = np.array([-res_deg[0], res_deg[1]])
    else:
        lon_min, lon_max, lat_min, lat_max = bounds_deg
        cdelt = np.array([-(lon_max - lon_min) / dims[0], (lat_max - lat_min) / dims[1]])
    w = Wcs(proj, crval, cdelt, is_azimuth)
    if dims is not None:
        shape = (int(dims[1]), int(dims[0]))
    else:
        lon_min, lon_max, lat_min, lat_max = bounds_deg
        c0, r0, _ = world2pix(w, lon_min, lat_min)
        c1, r1, _ = world2pix(w, lon_max, lat_max)
        n_col, n_row = int(abs(float(c1) - float(c0))), int(abs(float(r1) - float(r0)))
        n_col += n_col % 2
        n_row += n_row % 2
        shape = (n_row, n_col)
    oc, orow, _ = world2pix(w, crval[0], crval[1])
    off = np.array([float(oc), float(orow)])
    c_row, c_col = 0.5 * np.array(shape, dtype=np.float64) + 0.5 + off
    w.crpix = np.array([c_col, c_row])
    w.shape = shape
    return w, shape


# ---- the TOAST half ---------------------------------------------------------------------------
def quat_to_lonlat_deg(quats, is_azimuth=False):
    """qa_to_iso (math_qarray.cpp:738-772) -> to_lonlat_angles (qarray.py:511-534) ->
    center_offset_lonlat (pointing_utils.py:35-66, center_offset None, degrees=True)."""
    q = np.asarray(quats, dtype=np.float64)
    norm = np.zeros(q.shape[:-1])
    for j in range(4):
        norm = norm + q[..., j] * q[..., j]
    norm = 1.0 / np.sqrt(norm)
    x, y, z, w = (q[..., j] * norm for j in range(4))
    xw, yw = w * x, w * y
    x2, y2 = -x * x, -y * y
    xz, yz = x * z, y * z
    dx = 2 * (yw + xz) + 0.0
    dy = 2 * (yz - xw) + 0.0
    dz = 2 * (x2 + y2) + 1.0
    eps = np.finfo(np.float64).eps
    pole = np.abs(np.abs(dz) - 1.0) < eps
    with np.errstate(invalid="ignore"):
        theta = np.where(pole, np.where(dz >= 0.0, 0.0, np.pi), np.pi / 2 - np.arcsin(dz))
    phi = np.where(pole, 0.0, np.arctan2(dy, dx))
    lon = phi.copy()
    lat = 0.5 * np.pi - theta
    if is_azimuth:
        lon = 2 * np.pi - lon
    lon = np.where(lon >= 2 * np.pi, lon - 2 * np.pi, lon)
    lon = np.where(lon < 0, lon + 2 * np.pi, lon)
    return np.degrees(lon), np.degrees(lat)


def pixels_wcs(w, quats, flags=None, flag_mask=0):
    """pixels_wcs.py:590-616 for one detector: pixel numbers (int64), the fractional coordinates
    (for boundary-aware comparisons)."""
    lon, lat = quat_to_lonlat_deg(quats, w.is_azimuth)
    dcol, drow, ok = world2pix(w, lon, lat)
    dcol_s = np.where(ok, dcol, 0.0)
    drow_s = np.where(ok, drow, 0.0)
    col = np.around(dcol_s).astype(np.int64)
    row = np.around(drow_s).astype(np.int64)
    pix = col + row * w.shape[1]
    n_pix = w.shape[0] * w.shape[1]
    bad = (pix >= n_pix) | ~ok
    if flags is not None:
        bad |= (np.asarray(flags) & flag_mask) != 0
    pix[bad] = -1
    return pix, dcol, drow
