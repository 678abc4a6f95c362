"""Flat-sky projection set-up for ``ops.PixelsWCS``: the host-side half of what astropy / WCSLIB
do for the reference (``ops/pixels_wcs.py:190-345``: ``create_wcs``), restated from Calabretta &
Greisen 2002 (FITS WCS Paper II) and WCSLIB's ``celset``:

* CRVAL from the centre or the bounding box, CDELT = (-res_lon, res_lat) or from bounds / dims,
* the image shape from ``dims`` or from the projected bounding-box corners (made even),
* CRPIX so that CRVAL lands on the centre of the image (``0.5 * shape + 0.5 + off``),
* the Euler angles of the celestial -> native rotation with the default LONPOLE / LATPOLE:
  zenithal projections (TAN, ZEA) put the native pole at CRVAL (phi_p = 180); the cylindrical
  and pseudo-cylindrical ones (CAR, CEA, MER, SFL) put the native origin (0, 0) there
  (phi_p = 0 for CRVAL2 >= 0, else 180; the celestial pole of the native system follows from
  Paper II eq. 8-10 as WCSLIB solves them).

The per-sample arithmetic lives in ``csrc/tb_wcs.cuh`` (host / device) and in
``oracle/pixels_wcs.py`` (numpy, test infrastructure); this module only produces the numbers they
consume.  No astropy objects: the projection is described by plain attributes (``ctype``,
``crval``, ``cdelt``, ``crpix``, ``pv``) that a maintainer can copy into a ``astropy.wcs.WCS``.
"""

import math

import numpy as np

from . import lib as L

PROJECTIONS = {"CAR": 0, "CEA": 1, "MER": 2, "SFL": 3, "TAN": 4, "ZEA": 5}
_COORD = {"AZEL": ("TLON", "TLAT"), "EQU": ("RA--", "DEC-"), "GAL": ("GLON", "GLAT"),
          "ECL": ("ELON", "ELAT")}
D2R = math.pi / 180.0
R2D = 180.0 / math.pi


def _sincosd(a):
    if math.fmod(a, 90.0) == 0.0:
        i = abs(int(math.floor(a / 90.0 + 0.5))) % 4
        return [(0.0, 1.0), (1.0 if a > 0 else -1.0, 0.0), (0.0, -1.0),
                (-1.0 if a > 0 else 1.0, 0.0)][i]
    return math.sin(a * D2R), math.cos(a * D2R)


def _acosd(v):
    if v >= 1.0 and v - 1.0 < 1e-10:
        return 0.0
    if v == 0.0:
        return 90.0
    if v <= -1.0 and v + 1.0 > -1e-10:
        return 180.0
    return math.acos(v) * R2D


def _atan2d(y, x):
    if y == 0.0:
        return 0.0 if x >= 0.0 else 180.0
    if x == 0.0:
        return 90.0 if y > 0.0 else -90.0
    return math.atan2(y, x) * R2D


def celestial_euler(proj, crval):
    """WCSLIB celset for the default LONPOLE / LATPOLE: Euler angles (lng_p, 90 - lat_p, phi_p,
    cos, sin) of the rotation from celestial to native spherical coordinates."""
    lng0, lat0 = float(crval[0]), float(crval[1])
    zenithal = proj in ("TAN", "ZEA")
    theta0 = 90.0 if zenithal else 0.0
    phi0 = 0.0
    phip = phi0 + (0.0 if lat0 >= theta0 else 180.0)   # default LONPOLE
    latpole = 90.0
    if zenithal:
        lngp, latp = lng0, lat0
    else:
        slat0, clat0 = _sincosd(lat0)
        sphip, cphip = _sincosd(phip - phi0)
        sthe0, cthe0 = _sincosd(theta0)
        x = cthe0 * cphip
        y = sthe0
        z = math.sqrt(x * x + y * y)
        slz = slat0 / z
        u = _atan2d(y, x)
        v = _acosd(slz)
        latp1, latp2 = u + v, u - v
        for k in range(2):
            pass
        def norm(a):
            if a > 180.0:
                a -= 360.0
            elif a < -180.0:
                a += 360.0
            return a
        latp1, latp2 = norm(latp1), norm(latp2)
        # the solution closer to LATPOLE among the valid ones
        if abs(latpole - latp1) < abs(latpole - latp2):
            latp = latp1 if abs(latp1) < 90.0 + 5e-9 else latp2
        else:
            latp = latp2 if abs(latp2) < 90.0 + 5e-9 else latp1
        if abs(latp) < 90.0 + 5e-9:
            latp = max(-90.0, min(90.0, latp))
        z = math.cos(latp * D2R) * clat0
        if abs(z) < 5e-9:
            if abs(clat0) < 5e-9:
                lngp = lng0
            elif latp > 0.0:
                lngp = lng0 + phip - phi0 - 180.0
            else:
                lngp = lng0 - phip + phi0
        else:
            xx = (sthe0 - math.sin(latp * D2R) * slat0) / z
            yy = sphip * cthe0 / clat0
            if xx == 0.0 and yy == 0.0:
                raise RuntimeError("PixelsWCS: degenerate native pole")
            lngp = lng0 - _atan2d(yy, xx)
        # same sign as the longitude of the fiducial point
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
    s, c = _sincosd(e1)
    return np.array([lngp, e1, phip, c, s], dtype=np.float64)


class FlatWCS:
    """The projection parameters PixelsWCS works with (the subset of astropy.wcs.WCS that
    ``create_wcs`` sets: ops/pixels_wcs.py:268-345)."""

    def __init__(self, coord, proj, crval, cdelt):
        if proj not in PROJECTIONS:
            raise ValueError(f"Invalid WCS projection name '{proj}'")
        if coord not in _COORD:
            raise RuntimeError(f"Unsupported coordinate frame '{coord}'")
        self.coord, self.proj = coord, proj
        self.ctype = [f"{_COORD[coord][0]}-{proj}", f"{_COORD[coord][1]}-{proj}"]
        self.crval = np.array(crval, dtype=np.float64)
        self.cdelt = np.array(cdelt, dtype=np.float64)
        self.crpix = np.zeros(2)
        self.pv = []
        self.cea_lambda = 1.0
        if proj == "CEA":
            self.cea_lambda = math.cos(math.radians(self.crval[1])) ** 2
            self.pv = [(2, 1, self.cea_lambda)]
        self.euler = celestial_euler(proj, self.crval)
        self.shape = (0, 0)
        self.is_azimuth = coord == "AZEL"

    def world2pix(self, lon_deg, lat_deg):
        """wcs_world2pix(..., origin=0) through the SAME host/device code the kernel runs
        (csrc/tb_wcs.cuh compiled for the host is test infrastructure; here the numbers come from
        the small scalar restatement in this module so that set-up needs no device)."""
        from ._wcs_host import world2pix_scalar

        return world2pix_scalar(self, float(lon_deg), float(lat_deg))

    def desc(self):
        d = L.tb_wcs_desc()
        d.projection = PROJECTIONS[self.proj]
        d.is_azimuth = 1 if self.is_azimuth else 0
        for k in range(5):
            d.euler[k] = float(self.euler[k])
        for k in range(2):
            d.crpix[k] = float(self.crpix[k])
            d.cdelt[k] = float(self.cdelt[k])
        d.cea_lambda = float(self.cea_lambda)
        d.n_col, d.n_row = int(self.shape[1]), int(self.shape[0])
        return d


def create_wcs(coord="EQU", proj="CAR", center_deg=None, bounds_deg=None, res_deg=None,
               dims=None):
    """ops/pixels_wcs.py:190-345 (PixelsWCS.create_wcs).  Returns (FlatWCS, (n_row, n_col))."""
    if center_deg is not None:
        if bounds_deg is not None:
            raise RuntimeError("PixelsWCS: only one of center and bounds should be set.")
        if res_deg is None or dims is None:
            raise RuntimeError("PixelsWCS: when center is set, both resolution and dimensions"
                               " are required.")
        crval = np.array(center_deg, dtype=np.float64)
    else:
        if bounds_deg is None:
            raise RuntimeError("PixelsWCS: when center is not specified, bounds required.")
        lon_min, lon_max, lat_min, lat_max = bounds_deg
        crval = np.array([0.5 * (lon_min + lon_max), 0.5 * (lat_min + lat_max)])
        if res_deg is not None and dims is not None:
            raise RuntimeError("PixelsWCS: when using bounds, only one of resolution or"
                               " dimensions must be specified.")
    if center_deg is not None or res_deg is not None:
        cdelt = np.array([-res_deg[0], res_deg[1]])
    else:
        lon_min, lon_max, lat_min, lat_max = bounds_deg
        n_col, n_row = dims
        cdelt = np.array([-(lon_max - lon_min) / n_col, (lat_max - lat_min) / n_row])
    wcs = FlatWCS(coord, proj, crval, cdelt)
    if dims is not None:
        n_col, n_row = dims
        shape = (int(n_row), int(n_col))
    else:
        lon_min, lon_max, lat_min, lat_max = bounds_deg
        col_min, row_min = wcs.world2pix(lon_min, lat_min)
        col_max, row_max = wcs.world2pix(lon_max, lat_max)
        n_col = int(abs(col_max - col_min))
        n_row = int(abs(row_max - row_min))
        n_col += n_col % 2
        n_row += n_row % 2
        shape = (n_row, n_col)
    off = wcs.world2pix(crval[0], crval[1])
    c_row = 0.5 * shape[0] + 0.5 + off[0]
    c_col = 0.5 * shape[1] + 0.5 + off[1]
    # (the reference unpacks `0.5 * shape + 0.5 + off` with shape = (n_row, n_col) and
    # off = (col, row): rows pair with the column offset and vice versa; off is 0 - crpix + ... =
    # (-1, -1) before CRPIX is set, so the cross pairing is harmless and is reproduced)
    wcs.crpix = np.array([c_col, c_row], dtype=np.float64)
    wcs.shape = shape
    return wcs, shape
