"""PixelsWCS (/root/reference/src/toast/ops/pixels_wcs.py:39-662).

CPU part: the numpy oracle (oracle/pixels_wcs.py) against closed-form projection formulas, its
own inverse, and the reference test's pixel-centre KAT; the product's projection set-up
(toast_b200/wcs.py) against the oracle's; the product's per-sample code (csrc/tb_wcs.cuh compiled
for the host) against the oracle.  GPU part: the CUDA kernel and the operator against the oracle.

WCSLIB itself is not available here (astropy is not in the image): see the parity note at the top
of oracle/pixels_wcs.py."""

import ctypes as ct

import numpy as np
import pytest

import helpers as H
from helpers import S
from oracle import pixels_wcs as OW
from toast_b200 import wcs as W

CENTERS = [(130.0, -40.0), (130.0, 0.0), (180.0, -40.0), (180.0, 0.0), (40.0, 35.0)]
DIMS = (100, 50)   # tests/ops_pointing_wcs.py:35


def _pixel_centre_quats(w):
    """tests/ops_pointing_wcs.py:45-78: boresight aimed at every pixel centre."""
    n_row, n_col = w.shape
    cols, rows = np.meshgrid(np.arange(n_col), np.arange(n_row))
    lng, lat = OW.pix2world(w, cols.ravel().astype(float), rows.ravel().astype(float))
    theta = np.pi / 2 - np.radians(lat)
    return H.iso_quat(theta, np.radians(lng), np.zeros_like(theta)), cols.ravel() + rows.ravel() * n_col


@pytest.mark.parametrize("proj", OW.PROJECTIONS)
@pytest.mark.parametrize("center", CENTERS)
def test_oracle_round_trip_and_pixel_centre_kat(proj, center):
    for res in (0.02, 0.012, 1.0):
        w, shape = OW.create_wcs(proj, center_deg=center, res_deg=(res, res), dims=DIMS)
        assert shape == (DIMS[1], DIMS[0])
        quats, expect = _pixel_centre_quats(w)
        pix, dcol, drow = OW.pixels_wcs(w, quats)
        np.testing.assert_array_equal(pix, expect)       # every pixel hit exactly once
        assert np.max(np.abs(dcol - np.round(dcol))) < 1e-6


def test_oracle_against_closed_forms():
    rng = np.random.default_rng(0)
    for c in [(130.0, -40.0), (20.0, 50.0)]:
        a = c[0] + rng.uniform(-1, 1, 1000)
        d = c[1] + rng.uniform(-0.5, 0.5, 1000)
        a0, d0, ar, dr = np.radians(c[0]), np.radians(c[1]), np.radians(a), np.radians(d)
        # TAN: the gnomonic standard coordinates (xi, eta)
        w, _ = OW.create_wcs("TAN", center_deg=c, res_deg=(0.02, 0.02), dims=DIMS)
        D = np.sin(d0) * np.sin(dr) + np.cos(d0) * np.cos(dr) * np.cos(ar - a0)
        xi = np.cos(dr) * np.sin(ar - a0) / D
        eta = (np.cos(d0) * np.sin(dr) - np.sin(d0) * np.cos(dr) * np.cos(ar - a0)) / D
        dc, drw, ok = OW.world2pix(w, a, d)
        assert ok.all()
        assert np.max(np.abs(dc - ((np.degrees(xi) / w.cdelt[0] + w.crpix[0]) - 1))) < 1e-9
        assert np.max(np.abs(drw - ((np.degrees(eta) / w.cdelt[1] + w.crpix[1]) - 1))) < 1e-9
        # ZEA: r = 2 sin(angular distance / 2), same position angle as TAN
        w, _ = OW.create_wcs("ZEA", center_deg=c, res_deg=(0.02, 0.02), dims=DIMS)
        dist = np.arccos(np.clip(D, -1, 1))
        scale = np.where(dist > 0, 2 * np.sin(dist / 2) / np.tan(dist), 1.0)
        dc, drw, _ = OW.world2pix(w, a, d)
        assert np.max(np.abs(dc - ((np.degrees(xi * scale) / w.cdelt[0] + w.crpix[0]) - 1))) < 1e-7
        assert np.max(np.abs(drw - ((np.degrees(eta * scale) / w.cdelt[1] + w.crpix[1]) - 1))) < 1e-7
    # equatorial reference point: CAR is the identity, SFL scales longitude by cos(lat), CEA /
    # MER have their textbook y
    c = (130.0, 0.0)
    a = c[0] + rng.uniform(-1, 1, 500)
    d = rng.uniform(-0.5, 0.5, 500)
    for proj, fx, fy in (("CAR", lambda a, d: a - c[0], lambda a, d: d),
                         ("SFL", lambda a, d: (a - c[0]) * np.cos(np.radians(d)), lambda a, d: d),
                         ("CEA", lambda a, d: a - c[0], lambda a, d: np.degrees(np.sin(np.radians(d)))),
                         ("MER", lambda a, d: a - c[0],
                          lambda a, d: np.degrees(np.log(np.tan(np.pi / 4 + np.radians(d) / 2))))):
        w, _ = OW.create_wcs(proj, center_deg=c, res_deg=(0.02, 0.02), dims=DIMS)
        dc, drw, _ = OW.world2pix(w, a, d)
        assert np.max(np.abs(dc - ((fx(a, d) / w.cdelt[0] + w.crpix[0]) - 1))) < 1e-9, proj
        assert np.max(np.abs(drw - ((fy(a, d) / w.cdelt[1] + w.crpix[1]) - 1))) < 1e-9, proj


def test_create_wcs_matches_oracle_and_validates():
    for proj in OW.PROJECTIONS:
        for c in CENTERS:
            wo, so = OW.create_wcs(proj, center_deg=c, res_deg=(0.02, 0.03), dims=DIMS)
            wp, sp = W.create_wcs("EQU", proj, center_deg=c, res_deg=(0.02, 0.03), dims=DIMS)
            assert so == sp
            np.testing.assert_array_equal(wo.euler, wp.euler)
            np.testing.assert_array_equal(wo.crpix, wp.crpix)
            np.testing.assert_array_equal(wo.cdelt, wp.cdelt)
        b = (120.0, 140.0, -35.0, -25.0)                 # tests/ops_pointing_wcs.py:131
        wo, so = OW.create_wcs(proj, bounds_deg=b, res_deg=(0.01, 0.01))
        wp, sp = W.create_wcs("EQU", proj, bounds_deg=b, res_deg=(0.01, 0.01))
        assert so == sp and so[0] % 2 == 0 and so[1] % 2 == 0
        np.testing.assert_array_equal(wo.crpix, wp.crpix)
        wo, so = OW.create_wcs(proj, bounds_deg=b, dims=DIMS)
        wp, sp = W.create_wcs("EQU", proj, bounds_deg=b, dims=DIMS)
        np.testing.assert_array_equal(wo.cdelt, wp.cdelt)
    assert wp.ctype == ["RA---ZEA", "DEC--ZEA"]   # (PROJECTIONS ends with ZEA)
    with pytest.raises(RuntimeError):
        W.create_wcs("EQU", "CAR", center_deg=(1, 2), bounds_deg=(0, 1, 0, 1), res_deg=(1, 1),
                     dims=DIMS)
    with pytest.raises(RuntimeError):
        W.create_wcs("EQU", "CAR", center_deg=(1, 2), res_deg=(1, 1))
    with pytest.raises(RuntimeError):
        W.create_wcs("EQU", "CAR")
    with pytest.raises(ValueError):
        W.create_wcs("EQU", "XYZ", center_deg=(1, 2), res_deg=(1, 1), dims=DIMS)
    with pytest.raises(RuntimeError):
        W.create_wcs("J2000", "CAR", center_deg=(1, 2), res_deg=(1, 1), dims=DIMS)


def _random_quats(n, center, spread_deg, rng):
    lon = np.radians(center[0] + rng.uniform(-spread_deg, spread_deg, n))
    lat = np.radians(center[1] + rng.uniform(-spread_deg / 2, spread_deg / 2, n))
    return H.iso_quat(np.pi / 2 - lat, lon, rng.uniform(0, 2 * np.pi, n))


def _same_pixels(pix, ref, dcol, drow):
    """Identical, except where a fractional coordinate sits within 1e-9 of a rounding boundary
    (sin / cos / atan2 differ by an ulp between libm implementations)."""
    diff = pix != ref
    if not diff.any():
        return
    edge = (np.abs(np.abs(dcol - np.floor(dcol)) - 0.5) < 1e-9) | \
        (np.abs(np.abs(drow - np.floor(drow)) - 0.5) < 1e-9)
    assert np.all(edge[diff]), f"{int(diff.sum())} pixels differ away from rounding boundaries"


@pytest.mark.parametrize("proj", OW.PROJECTIONS)
@pytest.mark.parametrize("azel", [False, True])
def test_device_code_on_the_host_matches_oracle(proj, azel):
    lib = H.host_math_lib()
    rng = np.random.default_rng(3)
    for c in CENTERS[:3]:
        wo, shape = OW.create_wcs(proj, center_deg=c, res_deg=(0.02, 0.02), dims=DIMS,
                                  is_azimuth=azel)
        cc = (360.0 - c[0], c[1]) if azel else c
        quats = np.ascontiguousarray(np.vstack([_random_quats(4000, cc, 1.5, rng),
                                                _pixel_centre_quats(wo)[0][:500]
                                                if not azel else _random_quats(10, cc, 1, rng)]))
        n = len(quats)
        pix = np.zeros(n, dtype=np.int64)
        dcol, drow = np.zeros(n), np.zeros(n)
        proj_code = W.PROJECTIONS[proj]
        lib.tbw_quat2pix(ct.c_int64(n), quats.ctypes.data_as(ct.c_void_p), ct.c_int(proj_code),
                         ct.c_int(int(azel)), wo.euler.ctypes.data_as(ct.c_void_p),
                         wo.crpix.ctypes.data_as(ct.c_void_p),
                         wo.cdelt.ctypes.data_as(ct.c_void_p), ct.c_double(wo.lam),
                         ct.c_int64(shape[1]), ct.c_int64(shape[0]),
                         pix.ctypes.data_as(ct.c_void_p), dcol.ctypes.data_as(ct.c_void_p),
                         drow.ctypes.data_as(ct.c_void_p))
        ref, rc, rr = OW.pixels_wcs(wo, quats)
        assert np.mean(ref >= 0) > 0.3
        _same_pixels(pix, ref, rc, rr)
        ok = ref >= 0
        assert np.max(np.abs(dcol[ok] - rc[ok])) < 1e-9 and np.max(np.abs(drow[ok] - rr[ok])) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("proj", OW.PROJECTIONS)
def test_gpu_kernel_matches_oracle(proj):
    import torch

    from toast_b200 import kernels as K

    rng = np.random.default_rng(5)
    n_det, n_samp = 3, 5000
    for c in CENTERS[:3]:
        wo, shape = OW.create_wcs(proj, center_deg=c, res_deg=(0.02, 0.02), dims=DIMS)
        wp, _ = W.create_wcs("EQU", proj, center_deg=c, res_deg=(0.02, 0.02), dims=DIMS)
        quats = np.ascontiguousarray(_random_quats(n_det * n_samp, c, 1.5, rng)
                                     .reshape(n_det, n_samp, 4))
        kat, expect = _pixel_centre_quats(wo)
        quats[0, :len(kat)] = kat
        flags = (rng.random(n_samp) < 0.05).astype(np.uint8)
        idx = np.arange(n_det, dtype=np.int32)
        iv = S.make_intervals([(0, 2000), (2100, n_samp)])
        hits = np.zeros(4, dtype=np.uint8)
        n_pix_submap = (shape[0] * shape[1] + 3) // 4
        pix = torch.full((n_det, n_samp), -7, dtype=torch.int64, device="cuda")
        K.pixels_wcs(wp, idx, torch.from_numpy(quats).cuda(), torch.from_numpy(flags).cuda(), 1,
                     idx, pix, iv, hits, n_pix_submap)
        got = pix.cpu().numpy()
        assert np.all(got[:, 2000:2100] == -7)           # outside the view: untouched
        hit_ref = np.zeros(4, dtype=np.uint8)
        for d in range(n_det):
            ref, rc, rr = OW.pixels_wcs(wo, quats[d], flags, 1)
            for a, b in ((0, 2000), (2100, n_samp)):
                _same_pixels(got[d, a:b], ref[a:b], rc[a:b], rr[a:b])
                good = ref[a:b][ref[a:b] >= 0]
                hit_ref[good // n_pix_submap] = 1
        np.testing.assert_array_equal(hits, hit_ref)
        unflagged = flags[:len(kat)] == 0
        unflagged[2000:2100] = False
        np.testing.assert_array_equal(got[0, :len(kat)][unflagged[:len(kat)]],
                                      expect[unflagged[:len(kat)]])


@pytest.mark.gpu
def test_gpu_operator_pixel_centre_kat():
    """tests/ops_pointing_wcs.py:165-215 through the operator mirror: one boresight sample per
    pixel centre -> every pixel of the projection is hit exactly once per detector."""
    from toast_b200 import ops
    from toast_b200.data import Data, observation_from_synthetic

    for proj in OW.PROJECTIONS:
        wo, shape = OW.create_wcs(proj, center_deg=(130.0, -40.0), res_deg=(0.02, 0.02), dims=DIMS)
        bore, expect = _pixel_centre_quats(wo)
        n_samp = len(bore)
        obs = S.make_observation("c1", n_det=2, n_samp=n_samp, with_signal=False, flags=False)
        obs["boresight"] = np.ascontiguousarray(bore)
        obs["focalplane"] = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (2, 1))   # on the boresight
        obs["intervals"] = S.make_intervals([(0, n_samp)])
        data = Data()
        data.obs.append(observation_from_synthetic(obs))
        dp = ops.PointingDetectorSimple(view="scanning", shared_flags="flags", shared_flag_mask=1)
        pix = ops.PixelsWCS(detector_pointing=dp, projection=proj, auto_bounds=False,
                            center=(130.0, -40.0), resolution=(0.02, 0.02), dimensions=DIMS,
                            submaps=10, create_dist="pixel_dist")
        pix.apply(data)
        assert pix.wcs_shape == (DIMS[1], DIMS[0])
        p = data.obs[0].detdata["pixels"].data
        for d in range(2):
            np.testing.assert_array_equal(p[d], expect)
            assert np.array_equal(np.bincount(p[d], minlength=n_samp), np.ones(n_samp, dtype=int))
        dist = data["pixel_dist"]
        assert dist.n_pix == DIMS[0] * DIMS[1] and dist.n_local_submap == 10
        assert dist.wcs_shape == (DIMS[1], DIMS[0])


def _host_pixels_wcs(wcs, quat_index, quats, shared_flags, shared_flag_mask, pixel_index, pixels,
                     intervals, hit_submaps, n_pix_submap, use_accel=False, stream=None):
    """kernels.pixels_wcs with the device code run on the host (tb_wcs.cuh through
    tests/csrc/host_math.cpp): per detector and view, as tb_pixels_wcs lays the work out."""
    lib = H.host_math_lib()
    d = wcs.desc()
    euler = np.array(list(d.euler), dtype=np.float64)
    crpix = np.array(list(d.crpix), dtype=np.float64)
    cdelt = np.array(list(d.cdelt), dtype=np.float64)
    q, p = np.asarray(quats), np.asarray(pixels)
    for qi, pi in zip(quat_index, pixel_index):
        for iv in intervals:
            a, b = int(iv["first"]), int(iv["last"])
            n = b - a
            qq = np.ascontiguousarray(q[qi, a:b])
            out = np.zeros(n, dtype=np.int64)
            dc, dr = np.zeros(n), np.zeros(n)
            lib.tbw_quat2pix(ct.c_int64(n), qq.ctypes.data_as(ct.c_void_p), ct.c_int(d.projection),
                             ct.c_int(d.is_azimuth), euler.ctypes.data_as(ct.c_void_p),
                             crpix.ctypes.data_as(ct.c_void_p), cdelt.ctypes.data_as(ct.c_void_p),
                             ct.c_double(d.cea_lambda), ct.c_int64(d.n_col), ct.c_int64(d.n_row),
                             out.ctypes.data_as(ct.c_void_p), dc.ctypes.data_as(ct.c_void_p),
                             dr.ctypes.data_as(ct.c_void_p))
            if shared_flags is not None and len(shared_flags) == q.shape[1]:
                out[(np.asarray(shared_flags)[a:b] & shared_flag_mask) != 0] = -1
            p[pi, a:b] = out
            if hit_submaps is not None:
                hit_submaps[out[out >= 0] // n_pix_submap] = 1


def test_operator_pixel_centre_kat_on_the_host(monkeypatch):
    """The reference's own test of PixelsWCS (tests/ops_pointing_wcs.py:165-215) through the
    operator mirror with the kernel's code run on the host: a boresight aimed at every pixel
    centre hits every pixel of the projection exactly once, for all six projections; the
    operator's views, flags, pixel distribution and submaps as on the device."""
    import fake_device
    from toast_b200 import ops
    from toast_b200.data import Data, observation_from_synthetic
    import toast_b200.ops.pixels_wcs as PW

    fake_device.install_operator_kernels(monkeypatch)      # pointing_detector -> oracle
    monkeypatch.setattr(PW.K, "pixels_wcs", _host_pixels_wcs)
    for proj in OW.PROJECTIONS:
        wo, shape = OW.create_wcs(proj, center_deg=(130.0, -40.0), res_deg=(0.02, 0.02), dims=DIMS)
        bore, expect = _pixel_centre_quats(wo)
        n_samp = len(bore)
        obs = S.make_observation("c1", n_det=2, n_samp=n_samp, with_signal=False, flags=False)
        obs["boresight"] = np.ascontiguousarray(bore)
        obs["focalplane"] = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (2, 1))   # on the boresight
        obs["intervals"] = S.make_intervals([(0, n_samp)])
        data = Data()
        data.obs.append(observation_from_synthetic(obs))
        dp = ops.PointingDetectorSimple(view="scanning", shared_flags="flags", shared_flag_mask=1)
        pix = ops.PixelsWCS(detector_pointing=dp, projection=proj, auto_bounds=False,
                            center=(130.0, -40.0), resolution=(0.02, 0.02), dimensions=DIMS,
                            submaps=10, create_dist="pixel_dist")
        pix.apply(data)
        assert pix.wcs_shape == (DIMS[1], DIMS[0])
        p = data.obs[0].detdata["pixels"].data
        for d in range(2):
            np.testing.assert_array_equal(p[d], expect)
            assert np.array_equal(np.bincount(p[d], minlength=n_samp), np.ones(n_samp, dtype=int))
        dist = data["pixel_dist"]
        assert dist.n_pix == DIMS[0] * DIMS[1] and dist.n_local_submap == 10
        assert dist.wcs_shape == (DIMS[1], DIMS[0])


@pytest.mark.parametrize("proj", ["CAR", "TAN"])
def test_operator_auto_bounds_cover_the_scan(monkeypatch, proj):
    """pixels_wcs.py:436-489 (auto_bounds, the operator's default): the projection is sized from
    the boresight track plus the field of view, so every unflagged detector sample of the scan
    lands inside the image; the bounds contain the detectors' own lon / lat range."""
    import fake_device
    from toast_b200 import ops
    from toast_b200.data import Data, observation_from_synthetic
    import toast_b200.ops.pixels_wcs as PW

    fake_device.install_operator_kernels(monkeypatch)
    monkeypatch.setattr(PW.K, "pixels_wcs", _host_pixels_wcs)
    obs = S.make_observation("c2", n_det=6, n_samp=20000, with_signal=False)
    data = Data()
    data.obs.append(observation_from_synthetic(obs))
    dp = ops.PointingDetectorSimple(view="scanning", shared_flags="flags", shared_flag_mask=1)
    pix = ops.PixelsWCS(detector_pointing=dp, projection=proj, field_of_view=11.0,
                        resolution=(0.05, 0.05), submaps=8, create_dist="pixel_dist")
    assert pix.auto_bounds
    pix.apply(data)
    assert not pix.auto_bounds and len(pix.bounds) == 4
    lon_min, lon_max, lat_min, lat_max = pix.bounds
    assert lon_max - lon_min > 10.0 and lat_max - lat_min > 10.0      # track + field of view
    ob = data.obs[0]
    p = ob.detdata["pixels"].data
    good = np.zeros(obs["n_samp"], dtype=bool)
    for iv in obs["intervals"]:
        good[int(iv["first"]):int(iv["last"])] = True
    good &= (obs["shared_flags"] & 1) == 0
    assert good.sum() > 1000
    assert np.all(p[:, good] >= 0), "an unflagged sample fell outside the auto-sized image"
    flagged_in_view = np.zeros(obs["n_samp"], dtype=bool)
    for iv in obs["intervals"]:
        flagged_in_view[int(iv["first"]):int(iv["last"])] = True
    flagged_in_view &= (obs["shared_flags"] & 1) != 0
    assert np.all(p[:, flagged_in_view] == -1)
    n_row, n_col = pix.wcs_shape
    assert p.max() < n_row * n_col
    # the detectors' own directions lie inside the bounds
    quats = ob.detdata[dp.quats].data
    lon, lat = OW.quat_to_lonlat_deg(quats[:, good].reshape(-1, 4))
    lon = np.where(lon < lon_min - 180.0, lon + 360.0, lon)
    lon = np.where(lon > lon_max + 180.0, lon - 360.0, lon)
    assert lon.min() >= lon_min - 1e-9 and lon.max() <= lon_max + 1e-9
    assert lat.min() >= lat_min - 1e-9 and lat.max() <= lat_max + 1e-9
    dist = data["pixel_dist"]
    assert dist.n_pix == n_row * n_col and 1 <= dist.n_local_submap <= 8


def test_scan_range_refuses_a_track_through_the_pole():
    """pointing_utils.py:121-130: same refusal, same message, as the reference."""
    from toast_b200.ops.pixels_wcs import scan_range_lonlat_deg

    lat = np.radians(np.linspace(80.0, 88.0, 50))
    bore = H.iso_quat(np.pi / 2 - lat, np.zeros(50), np.zeros(50))
    with pytest.raises(RuntimeError, match="includes the zenith"):
        scan_range_lonlat_deg(bore, None, 0, 6.0, False)
    lo, hi, la0, la1 = scan_range_lonlat_deg(bore, None, 0, 2.0, False)   # 88 + 1 < 90: fine
    assert la1 <= 89.0 + 1e-9 and la0 >= 79.0 - 1e-9
