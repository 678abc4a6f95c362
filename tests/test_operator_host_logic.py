"""The operator mirror's HOST LOGIC in the CPU suite (views, detector rows, flags, `ensure()` /
skip, pixel distribution, Pipeline): the `_libtoast` kernels the operators call are routed to the
oracle (tests/fake_device.py::install_operator_kernels -- it has the reference's signatures), the
expectations are formed by calling the oracle directly on the raw arrays, as the GPU tests do with
the real kernels (tests/test_gpu_ops.py)."""

import numpy as np

import fake_device
from helpers import O, S, assert_close_norm
from toast_b200 import ops
from toast_b200.data import Data, observation_from_synthetic
from toast_b200.templates import Offset


def _data(name, n_det, n_samp, nside=64):
    obs = S.make_observation(name, n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03)
    data = Data()
    data.obs.append(observation_from_synthetic(obs))
    return obs, data


def test_operator_chain_host_logic(monkeypatch):
    fake_device.install_operator_kernels(monkeypatch)
    n_det = 4
    obs, data = _data("c2", n_det, 12000)
    ob = data.obs[0]
    pb = O.build_problem(obs, O)
    dp = ops.PointingDetectorSimple(view="scanning", shared_flags="flags", shared_flag_mask=1)
    pix = ops.PixelsHealpix(detector_pointing=dp, nside=obs["nside"], nest=obs["nest"],
                            create_dist="pixel_dist")
    wts = ops.StokesWeights(detector_pointing=dp, mode="IQU")
    ops.Pipeline(operators=[pix, wts]).apply(data, use_accel=False)
    np.testing.assert_array_equal(ob.detdata["pixels"].data, pb.pixels)
    np.testing.assert_array_equal(ob.detdata["weights"].data, pb.weights)
    dist = data["pixel_dist"]
    np.testing.assert_array_equal(dist.local_submaps, pb.local_submaps)
    np.testing.assert_array_equal(dist.global_submap_to_local, pb.global2local)
    # `exists => skip` (pixels_healpix.py:215-243)
    ob.detdata["pixels"].data[0, :10] = -5
    pix.apply(data)
    assert np.all(ob.detdata["pixels"].data[0, :10] == -5)
    ob.detdata["pixels"].data[:] = pb.pixels

    idx = np.arange(n_det, dtype=np.int32)
    build = ops.BuildNoiseWeighted(pixel_dist="pixel_dist", zmap="zmap", view="scanning",
                                   det_flags="flags", det_flag_mask=1, shared_flags="flags",
                                   shared_flag_mask=1)
    ops.Pipeline(operators=[build]).apply(data, use_accel=False)
    z_ref = np.zeros((pb.n_local_submap, pb.n_pix_submap, 3))
    O.build_noise_weighted(pb.global2local, z_ref, idx, pb.pixels, idx, pb.weights, idx,
                           obs["signal"], idx, obs["det_flags"], pb.det_scale, 1, pb.intervals,
                           obs["shared_flags"], 1, False)
    np.testing.assert_array_equal(data["zmap"].data, z_ref)

    before = ob.detdata["signal"].data.copy()
    scan = ops.ScanMap(pixels="pixels", weights="weights", map_key="zmap", view="scanning",
                       subtract=True)
    ops.Pipeline(operators=[scan]).apply(data, use_accel=False)
    d_ref = obs["signal"].copy()
    O.scan_map(pb.global2local, pb.n_pix_submap, z_ref, d_ref, idx, pb.pixels, idx, pb.weights,
               idx, pb.intervals, 1.0, False, True, False, False)
    np.testing.assert_array_equal(ob.detdata["signal"].data, d_ref)
    scan.subtract = False
    scan.apply(data)
    assert_close_norm(ob.detdata["signal"].data, before, rtol=1e-12, what="scan round trip")
    ob.detdata["signal"].data[:] = before

    ops.NoiseWeight(noise_model="noise_model", view="scanning").apply(data)
    d_ref = obs["signal"].copy()
    O.noise_weight(d_ref, idx, pb.intervals, pb.det_scale, False)
    np.testing.assert_array_equal(ob.detdata["signal"].data, d_ref)


def test_template_matrix_offset_host_logic(monkeypatch):
    """tests/template_offset.py:26-92 through TemplateMatrix / templates.Offset: layout, flags,
    variance, project(add(1)) = samples per step."""
    fake_device.install_operator_kernels(monkeypatch)
    obs, data = _data("c2", 4, 12000)
    ob = data.obs[0]
    tmpl = Offset(name="baselines", step_time=obs["step_time"], times="times",
                  noise_model="noise_model", det_flags="flags", det_flag_mask=1)
    tmat = ops.TemplateMatrix(templates=[tmpl], amplitudes="amps", view="scanning",
                              det_data="signal", det_flags="flags", det_flag_mask=1)
    tmat.transpose = True
    tmat.apply(data)
    amps = data["amps"]["baselines"]
    nav, det_start, n_amp = O.offset_layout(4, obs["intervals"], obs["step_length"])
    assert amps.n_local == n_amp
    ref = np.zeros(n_amp)
    aflags = amps.local_flags.copy()
    for d in range(4):
        O.template_offset_project_signal(d, obs["signal"], d, obs["det_flags"], 1,
                                         obs["step_length"], int(det_start[d]), nav, ref, aflags,
                                         obs["intervals"])
    np.testing.assert_array_equal(amps.local, ref)
    sf = (obs["det_flags"] & 1).astype(np.uint8)
    var_ref, fl_ref = O.offset_variance(4, obs["n_samp"], obs["intervals"], obs["step_length"],
                                        nav, obs["detweight"], sf, 1)
    np.testing.assert_array_equal(aflags, fl_ref)
    np.testing.assert_array_equal(tmpl._offsetvar, var_ref)
    ob.detdata["signal"].data[:] = 0
    amps.local[:] = 1.0
    amps.local_flags[:] = 0
    tmat.transpose = False
    tmat.apply(data)
    out = amps.duplicate()
    out.reset()
    data["amps2"] = type(data["amps"])()
    data["amps2"]["baselines"] = out
    tmpl.det_flags = None
    tmat2 = ops.TemplateMatrix(templates=[tmpl], amplitudes="amps2", view="scanning",
                               det_data="signal", transpose=True)
    tmat2._initialized = True
    tmat2.apply(data)
    lens = np.concatenate([
        np.minimum(obs["step_length"],
                   int(v["last"] - v["first"]) - obs["step_length"] * np.arange(na))
        for v, na in zip(obs["intervals"], nav)])
    np.testing.assert_array_equal(out.local, np.tile(lens, 4).astype(np.float64))


def test_covariance_operators_and_binmap_host_logic(monkeypatch):
    """CovarianceAndHits / BuildHitMap / BuildInverseCovariance / BinMap / covariance_rcond
    (mapmaker_utils.py:114-515, 1131-1270; mapmaker_binning.py:27-294; covariance.py:20-306):
    the operators' glue around the accumulation, inversion and apply kernels."""
    from toast_b200.covariance import covariance_rcond

    fake_device.install_operator_kernels(monkeypatch)
    n_det = 6
    obs, data = _data("c2", n_det, 12000)
    pb = O.build_problem(obs, O, rcond_threshold=1.0e-3)
    dp = ops.PointingDetectorSimple(view="scanning", shared_flags="flags", shared_flag_mask=1)
    pix = ops.PixelsHealpix(detector_pointing=dp, nside=obs["nside"], nest=obs["nest"],
                            create_dist="pixel_dist")
    wts = ops.StokesWeights(detector_pointing=dp, mode="IQU")
    ops.Pipeline(operators=[pix, wts]).apply(data)

    cah = ops.CovarianceAndHits(pixel_dist="pixel_dist", view="scanning", hits="hits",
                                inverse_covariance="invcov", covariance="cov", rcond="rcond",
                                rcond_threshold=1.0e-3)
    cah.apply(data)
    # the oracle's build_problem accumulates with the same flags (input flags, bad pointing)
    np.testing.assert_array_equal(data["invcov"].raw, pb.invcov)
    np.testing.assert_array_equal(data["cov"].data, pb.cov)
    np.testing.assert_array_equal(data["rcond"].raw, pb.rcond)
    assert data["hits"].data.dtype == np.int64 and data["hits"].raw.sum() > 0
    np.testing.assert_array_equal(covariance_rcond(data["invcov"], 1.0e-3).raw, pb.rcond)

    ops.BuildHitMap(pixel_dist="pixel_dist", view="scanning", hits="hits2").apply(data)
    np.testing.assert_array_equal(data["hits2"].raw, data["hits"].raw)
    ops.BuildInverseCovariance(pixel_dist="pixel_dist", view="scanning",
                               inverse_covariance="invcov2").apply(data)
    np.testing.assert_array_equal(data["invcov2"].raw, pb.invcov)

    # BinMap = pointing (already there: skipped) + BuildNoiseWeighted + covariance_apply
    binner = ops.BinMap(name="bin", pixel_dist="pixel_dist", covariance="cov", binned="binned",
                        pixel_pointing=pix, stokes_weights=wts, noise_model="noise_model",
                        full_pointing=True)
    binner.apply(data)
    idx = np.arange(n_det, dtype=np.int32)
    z_ref = np.zeros((pb.n_local_submap, pb.n_pix_submap, 3))
    O.build_noise_weighted(pb.global2local, z_ref, idx, pb.pixels, idx, pb.weights, idx,
                           obs["signal"], idx, obs["det_flags"], pb.det_scale, 1, pb.intervals,
                           obs["shared_flags"], 1, False)
    O.cov_apply_diag(pb.n_local_submap, pb.n_pix_submap, 3, pb.cov.reshape(-1), z_ref.reshape(-1))
    np.testing.assert_array_equal(data["binned"].data, z_ref)
    assert "bin_zmap" not in data
    # one detector at a time (full_pointing=False) sums the same map
    binner2 = ops.BinMap(name="bin1", pixel_dist="pixel_dist", covariance="cov", binned="binned1",
                         pixel_pointing=pix, stokes_weights=wts, noise_model="noise_model",
                         full_pointing=False, noiseweighted="zkeep")
    binner2.apply(data)
    assert_close_norm(data["binned1"].data, z_ref, rtol=1e-13, what="BinMap, single detectors")
    assert "zkeep" in data


def test_scan_mask_host_logic(monkeypatch):
    """ops/scan_map/scan_map.py:216-357."""
    from toast_b200.pixels import PixelData

    fake_device.install_operator_kernels(monkeypatch)
    obs, data = _data("c2", 4, 12000)
    ob = data.obs[0]
    dp = ops.PointingDetectorSimple(view="scanning", shared_flags="flags", shared_flag_mask=1)
    pix = ops.PixelsHealpix(detector_pointing=dp, nside=obs["nside"], nest=obs["nest"],
                            create_dist="pixel_dist")
    pix.apply(data)
    dist = data["pixel_dist"]
    mask = PixelData(dist, np.uint8, n_value=1)
    rng = np.random.default_rng(8)
    mask.data[:] = rng.integers(0, 4, mask.data.shape).astype(np.uint8)
    data["mask"] = mask
    before = ob.detdata["flags"].data.copy()
    ops.ScanMask(det_flags="flags", det_flags_value=4, pixels="pixels", mask_key="mask",
                 mask_bits=2, view="scanning").apply(data)
    expect = before.copy()
    p = ob.detdata["pixels"].data
    for iv in obs["intervals"]:
        a, b = int(iv["first"]), int(iv["last"])
        sm, lp = dist.global_pixel_to_submap(p[:, a:b])
        ok = sm >= 0
        hit = np.zeros(sm.shape, dtype=bool)
        hit[ok] = (mask.data[sm[ok], lp[ok], 0] & 2) != 0
        expect[:, a:b] |= np.where(hit, 4, 0).astype(np.uint8)
    np.testing.assert_array_equal(ob.detdata["flags"].data, expect)
    assert (expect != before).any()


def test_pipeline_staging_follows_requires_and_provides(monkeypatch):
    """pipeline.py:208-303 with use_accel=True: everything the operators require is created on
    the device and filled before the first operator runs, outputs are created by `ensure(...,
    accel=True)`, what the pipeline provides is copied back at the end, what it staged itself is
    dropped, and a buffer somebody else had staged stays where it was.  The device table is a
    recording stand-in (the kernels work on the host arrays)."""
    import toast_b200._libtoast as KP

    fake_device.install_operator_kernels(monkeypatch)
    table, log = {}, []
    key = lambda arr: arr.__array_interface__["data"][0]

    def create(arr, name):
        assert key(arr) not in table, f"{name}: created twice"
        table[key(arr)] = name
        log.append(("create", name))

    def need(arr, name, what):
        assert key(arr) in table, f"{what} of '{name}', which is not on the device"
        log.append((what, name))

    monkeypatch.setattr(KP, "accel_present", lambda arr, name: key(arr) in table)
    monkeypatch.setattr(KP, "accel_create", create)
    monkeypatch.setattr(KP, "accel_update_device", lambda a, n: need(a, n, "update_device"))
    monkeypatch.setattr(KP, "accel_update_host", lambda a, n: need(a, n, "update_host"))
    monkeypatch.setattr(KP, "accel_reset", lambda a, n: need(a, n, "reset"))

    def delete(arr, name):
        need(arr, name, "delete")
        del table[key(arr)]

    monkeypatch.setattr(KP, "accel_delete", delete)

    n_det = 4
    obs, data = _data("c2", n_det, 6000)
    ob = data.obs[0]
    pb = O.build_problem(obs, O)
    ob.detdata["signal"].accel_create("signal")          # staged by "somebody else"
    ob.detdata["signal"].accel_update_device("signal")
    del log[:]
    dp = ops.PointingDetectorSimple(view="scanning", shared_flags="flags", shared_flag_mask=1)
    pix = ops.PixelsHealpix(detector_pointing=dp, nside=obs["nside"], nest=obs["nest"],
                            create_dist="pixel_dist")
    wts = ops.StokesWeights(detector_pointing=dp, mode="IQU")
    ops.Pipeline(operators=[pix, wts]).apply(data, use_accel=True)
    build = ops.BuildNoiseWeighted(pixel_dist="pixel_dist", zmap="zmap", view="scanning",
                                   det_flags="flags", det_flag_mask=1, shared_flags="flags",
                                   shared_flag_mask=1)
    ops.Pipeline(operators=[build]).apply(data, use_accel=True)

    # results as without the device table
    np.testing.assert_array_equal(ob.detdata["pixels"].data, pb.pixels)
    np.testing.assert_array_equal(ob.detdata["weights"].data, pb.weights)
    idx = np.arange(n_det, dtype=np.int32)
    z_ref = np.zeros((pb.n_local_submap, pb.n_pix_submap, 3))
    O.build_noise_weighted(pb.global2local, z_ref, idx, pb.pixels, idx, pb.weights, idx,
                           obs["signal"], idx, obs["det_flags"], pb.det_scale, 1, pb.intervals,
                           obs["shared_flags"], 1, False)
    np.testing.assert_array_equal(data["zmap"].data, z_ref)

    names = lambda what: [n for w, n in log if w == what]
    # inputs were staged before use, outputs came back, the pipelines cleaned up after themselves
    assert "boresight_radec" in names("create") and "boresight_radec" in names("update_device")
    for out in ("pixels", "weights", "zmap"):
        assert out in names("create") and out in names("update_host"), out
    assert "signal" not in names("create") and "signal" not in names("delete")
    assert sorted(table.values()) == ["signal"]          # only the pre-staged buffer is left
    first_exec = min(i for i, (w, n) in enumerate(log) if (w, n) == ("create", "pixels"))
    assert log.index(("update_device", "boresight_radec")) < first_exec
