"""GPU tests of the host-side operator mirror (toast_b200.ops / templates / MapMaker) against
the oracle: the same operator chain the reference's tests drive
(tests/ops_pointing_healpix.py, ops_mapmaker_utils.py, ops_scan_map.py, template_offset.py,
ops_mapmaker_binning.py, ops_mapmaker.py), with the reference trait names."""

import numpy as np
import pytest

import helpers as H
from helpers import O, S, assert_close_norm
from toast_b200 import ops
from toast_b200.data import Data, observation_from_synthetic
from toast_b200.templates import Offset

pytestmark = pytest.mark.gpu


def _data(name="c1", n_det=4, n_samp=6000, nside=64, **kw):
    obs = S.make_observation(name, n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03, **kw)
    data = Data()
    data.obs.append(observation_from_synthetic(obs))
    return obs, data


def _pointing_ops(obs, view="scanning"):
    dp = ops.PointingDetectorSimple(view=view, shared_flags="flags", shared_flag_mask=1)
    pix = ops.PixelsHealpix(detector_pointing=dp, nside=obs["nside"], nest=obs["nest"],
                            create_dist="pixel_dist")
    wts = ops.StokesWeights(detector_pointing=dp, mode="IQU")
    return dp, pix, wts


@pytest.mark.parametrize("use_accel", [False, True])
@pytest.mark.parametrize("name,n_det,n_samp", [("c1", 4, 6000), ("c2", 6, 24000)])
def test_operator_chain_matches_oracle(name, n_det, n_samp, use_accel):
    ck = H.checker()
    obs, data = _data(name, n_det, n_samp)
    ob = data.obs[0]
    pb = O.build_problem(obs, ck)
    dp, pix, wts = _pointing_ops(obs)
    pipe = ops.Pipeline(operators=[pix, wts])
    pipe.apply(data, use_accel=use_accel)
    np.testing.assert_array_equal(ob.detdata["pixels"].data, pb.pixels)
    assert_close_norm(ob.detdata["weights"].data, pb.weights, what="weights")
    dist = data["pixel_dist"]
    np.testing.assert_array_equal(dist.local_submaps, pb.local_submaps)
    np.testing.assert_array_equal(dist.global_submap_to_local, pb.global2local)

    # a second apply must be a no-op (`exists => skip`, pixels_healpix.py:215-243)
    ob.detdata["pixels"].data[0, :10] = -5
    pix.apply(data)
    assert np.all(ob.detdata["pixels"].data[0, :10] == -5)
    ob.detdata["pixels"].data[:] = pb.pixels

    # BuildNoiseWeighted + covariance_apply == BinMap (tests/ops_mapmaker_binning.py:27-129)
    build = ops.BuildNoiseWeighted(pixel_dist="pixel_dist", zmap="zmap", view="scanning",
                                   det_flags="flags", det_flag_mask=1, shared_flags="flags",
                                   shared_flag_mask=1)
    ops.Pipeline(operators=[build]).apply(data, use_accel=use_accel)
    idx = np.arange(n_det, dtype=np.int32)
    z_ref = np.zeros((pb.n_local_submap, pb.n_pix_submap, 3))
    ck.build_noise_weighted(pb.global2local, z_ref, idx, pb.pixels, idx, pb.weights, idx,
                            obs["signal"], idx, obs["det_flags"], pb.det_scale, 1, pb.intervals,
                            obs["shared_flags"], 1, False)
    assert_close_norm(data["zmap"].data, z_ref, what="zmap")

    # ScanMap subtract then add back == original (tests/ops_scan_map.py:99-172)
    before = ob.detdata["signal"].data.copy()
    scan = ops.ScanMap(pixels="pixels", weights="weights", map_key="zmap", view="scanning",
                       subtract=True)
    ops.Pipeline(operators=[scan]).apply(data, use_accel=use_accel)
    d_ref = obs["signal"].copy()
    H.scan_fn(ck)(pb.global2local, pb.n_pix_submap, z_ref, d_ref, idx, pb.pixels, idx, pb.weights,
                  idx, pb.intervals, 1.0, False, True, False, False)
    assert_close_norm(ob.detdata["signal"].data, d_ref, what="scanned")
    scan.subtract = False
    scan.apply(data)
    assert_close_norm(ob.detdata["signal"].data, before, rtol=1e-12, what="scan round trip")
    ob.detdata["signal"].data[:] = before

    # NoiseWeight
    nw = ops.NoiseWeight(noise_model="noise_model", view="scanning")
    nw.apply(data)
    d_ref = obs["signal"].copy()
    ck.noise_weight(d_ref, idx, pb.intervals, pb.det_scale, False)
    np.testing.assert_array_equal(ob.detdata["signal"].data, d_ref)


def test_template_matrix_offset():
    """tests/template_offset.py:26-92: project(add(1)) counts the samples of every step."""
    obs, data = _data("c2", 4, 12000)
    ob = data.obs[0]
    tmpl = Offset(name="baselines", step_time=obs["step_time"], times="times",
                  noise_model="noise_model", det_flags="flags", det_flag_mask=1)
    tmat = ops.TemplateMatrix(templates=[tmpl], amplitudes="amps", view="scanning",
                              det_data="signal", det_flags="flags", det_flag_mask=1)
    tmat.transpose = True
    tmat.apply(data)  # creates zero amplitudes and projects the signal
    amps = data["amps"]["baselines"]
    nav, det_start, n_amp = O.offset_layout(4, obs["intervals"], obs["step_length"])
    assert amps.n_local == n_amp
    ref = np.zeros(n_amp)
    aflags = amps.local_flags.copy()
    for d in range(4):
        O.template_offset_project_signal(d, obs["signal"], d, obs["det_flags"], 1,
                                         obs["step_length"], int(det_start[d]), nav, ref, aflags,
                                         obs["intervals"])
    assert_close_norm(amps.local, ref, what="projected amplitudes")
    # offset variance / flags against the oracle restatement of offset.py:283-344
    sf = (obs["det_flags"] & 1).astype(np.uint8)
    var_ref, fl_ref = O.offset_variance(4, obs["n_samp"], obs["intervals"], obs["step_length"],
                                        nav, obs["detweight"], sf, 1)
    np.testing.assert_array_equal(aflags, fl_ref)
    assert_close_norm(tmpl._offsetvar, var_ref, what="offset variance")
    # add_to_signal of unit amplitudes then projection without flags == step sizes
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
    # dot ignores flagged amplitudes (templates/amplitudes.py:523-571)
    amps.local_flags[::3] = 1
    assert amps.dot(amps) == float(np.sum(amps.local_flags == 0))


@pytest.mark.parametrize("name,n_det,n_samp,nside", [("c1", 4, 6000, 64), ("c2", 6, 24000, 64)])
@pytest.mark.parametrize("regen", [False, True])
def test_mapmaker_end_to_end(name, n_det, n_samp, nside, regen):
    """ops.MapMaker with templates.Offset (tests/ops_mapmaker.py:53-187) against the oracle's
    restatement of the same stages driven by the reference kernels."""
    ck = H.checker()
    covapply = ck.cov_apply_diag
    obs, data = _data(name, n_det, n_samp, nside)
    pb = O.build_problem(obs, ck, rcond_threshold=1.0e-3)
    dp, pix, wts = _pointing_ops(obs)
    binning = ops.BinMap(pixel_dist="pixel_dist", covariance="cov", pixel_pointing=pix,
                         stokes_weights=wts, noise_model="noise_model", full_pointing=True)
    tmpl = Offset(name="baselines", step_time=obs["step_time"], times="times",
                  noise_model="noise_model")
    tmat = ops.TemplateMatrix(templates=[tmpl], amplitudes="amplitudes")
    mapper = ops.MapMaker(name="mm", det_data="signal", binning=binning, template_matrix=tmat,
                          solve_rcond_threshold=1.0e-3, map_rcond_threshold=1.0e-3, iter_max=12,
                          convergence=1.0e-30, regenerate_pointing=regen)
    signal0 = obs["signal"].copy()
    mapper.apply(data)

    # hit map: bit-exact
    hits_ref = np.zeros(pb.n_local_submap * pb.n_pix_submap, dtype=np.int64)
    sf0 = ((obs["det_flags"] & 1) != 0) | ((obs["shared_flags"] & 1) != 0)[None, :]
    for d in range(n_det):
        for iv in pb.intervals:
            a, b = int(iv["first"]), int(iv["last"])
            sm, lp = O.global_to_local(pb.pixels[d, a:b], pb.n_pix_submap, pb.global2local)
            lp[sf0[d, a:b]] = -1
            O.cov_accum_diag_hits(pb.n_local_submap, pb.n_pix_submap, 3, sm, lp, hits_ref)
    np.testing.assert_array_equal(data["mm_hits"].raw, hits_ref)
    assert_close_norm(data["mm_cov"].data, pb.cov, what="covariance")
    np.testing.assert_array_equal(data["mm_rcond"].raw > 0, pb.rcond > 0)

    # raw binned map
    binned_ref = O.bin_map(pb, ck, signal0, covapply)
    assert_close_norm(data["mm_binmap"].data, binned_ref, what="binned map")

    # PCG residual history
    rhs_ref = O.solver_rhs(pb, ck, signal0, covapply)
    amps_ref, hist_ref = O.solve(pb, ck, rhs_ref, convergence=1e-30, n_iter_max=12,
                                 covapply=covapply)
    H.assert_history_matches(mapper.history, hist_ref, H.pcg_envelope(pb, rhs_ref, 12), what=name)

    # destriped map: bin(signal - F a) with the SAME amplitudes must equal the oracle's binning
    # of that cleaned timestream to 1e-10 (the amplitudes themselves agree to the CG noise floor)
    amps = data["amplitudes"]["baselines"].local
    clean = signal0.copy()
    O.template_add(pb, O, -amps, clean)
    destriped_ref = O.bin_map(pb, ck, clean, covapply)
    assert_close_norm(data["mm_map"].data, destriped_ref, what="destriped map")
    # and the solved amplitudes reproduce the oracle's to the level CG noise allows
    dev = np.abs(amps - amps_ref).max() / np.abs(amps_ref).max()
    assert dev < max(1e-10, 1e3 * H.pcg_envelope(pb, rhs_ref, 12)[-1] ** 0.5), dev
    # the cleaned timestream is left in det_data (mapmaker.py:531-574)
    assert_close_norm(data.obs[0].detdata["signal"].data, clean, what="cleaned TOD")


def test_scan_mask():
    """ops/scan_map/scan_map.py:216-357 (used by the map-maker for the rcond / pixel masks)."""
    from toast_b200.pixels import PixelData

    obs, data = _data("c2", 4, 12000)
    ob = data.obs[0]
    dp, pix, wts = _pointing_ops(obs)
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
