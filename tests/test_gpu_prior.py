"""GPU parity tests of the Offset noise prior: the convolution / banded-solve
kernels against the scipy calls the reference makes (through oracle/offset_prior.py), and the
device-resident PCG with the prior against the oracle's restatement of solve()."""

import numpy as np
import pytest
import torch

import helpers as H
from helpers import O, S, assert_close_norm
from oracle import offset_prior as OP
from test_offset_prior import build_product, make_case
from toast_b200.solver import DeviceObservation, Destriper
from toast_b200.templates import offset_prior as PP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precond_width", [20, 4, 1])
@pytest.mark.parametrize("where", ["host", "device"])
def test_prior_kernels_match_scipy(precond_width, where):
    case = make_case(precond_width, n_amp_views=(120, 37, 700, 5), n_det=4)
    prior = build_product(case, cut=(2,)).finish()
    n = case["n_amp"]
    per = case["per"]
    rng = np.random.default_rng(8)
    a_in = rng.standard_normal(n)
    flags = (rng.random(n) < 0.1).astype(np.uint8)
    out0 = rng.standard_normal(n)

    ref_add = out0.copy()
    OP.add_prior(case["prior"], a_in, flags, ref_add)
    ref_add[2 * per:3 * per] = 0.0            # the cut detector (offset.py:946-947)
    ref_pre = np.zeros(n)
    OP.apply_precond(case["prior"], a_in, flags, ref_pre)
    ref_pre[2 * per:3 * per] = 0.0            # (offset.py:1003-1005)

    if where == "host":
        out = out0.copy()
        prior.add(a_in, flags, out)
        pre = np.full(n, 7.0)
        prior.precond(a_in, flags, pre)
    else:
        a_d, f_d = torch.from_numpy(a_in).cuda(), torch.from_numpy(flags).cuda()
        o_d = torch.from_numpy(out0).cuda()
        prior.add(a_d, f_d, o_d)
        p_d = torch.full((n,), 7.0, dtype=torch.float64, device="cuda")
        prior.precond(a_d, f_d, p_d)
        out, pre = o_d.cpu().numpy(), p_d.cpu().numpy()
    assert_close_norm(out, ref_add, rtol=1e-12, what="add_prior")
    assert_close_norm(pre, ref_pre, rtol=1e-12, what="apply_precond")
    assert np.all(out[flags != 0] == 0.0) and np.all(pre[flags != 0] == 0.0)
    with pytest.raises(RuntimeError):
        prior.add(a_in, flags, a_in)          # in place
    with pytest.raises(RuntimeError):
        prior.add(a_in[:-1], flags[:-1], out[:-1])


def test_prior_descriptor_validation():
    from toast_b200 import lib as L
    import ctypes as ct

    lib = L.load()
    d = L.tb_offset_prior_desc()
    seg = np.array([0, 4], dtype=np.int64)
    ln = np.array([6, 6], dtype=np.int64)      # overlapping segments
    fs = np.array([0, 0], dtype=np.int64)
    fl = np.array([3, 3], dtype=np.int64)
    taps = np.ones(3)
    d.n_amp, d.n_seg = 12, 2
    d.seg_start, d.seg_len = seg.ctypes.data, ln.ctypes.data
    d.filt_start, d.filt_len = fs.ctypes.data, fl.ctypes.data
    d.filters, d.n_filter_values = taps.ctypes.data, 3
    d.precond_mode = L.TB_PRECOND_TOEPLITZ
    d.prec_start, d.prec_width = fs.ctypes.data, fl.ctypes.data
    d.precond, d.n_precond_values = taps.ctypes.data, 3
    assert not lib.tb_offset_prior_create(ct.byref(d))
    assert "disjoint" in L.last_error()
    seg[1] = 6
    d.precond_mode = 7
    assert not lib.tb_offset_prior_create(ct.byref(d))
    d.precond_mode = L.TB_PRECOND_TOEPLITZ
    h = lib.tb_offset_prior_create(ct.byref(d))
    assert h
    lib.tb_offset_prior_destroy(h)


def _problem(name, n_det, n_samp, nside):
    obs = S.make_observation(name, n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03)
    # no input flags and a low rcond threshold: no flagged baselines -- with flagged baselines
    # the reference's banded preconditioner cannot be built at all (1 / offsetvar = inf fails
    # cholesky_banded's finiteness check, offset.py:520-531)
    pb = O.build_problem(obs, O, rcond_threshold=1e-8, use_flags=False)
    assert int(pb.amp_flags.sum()) == 0
    return obs, pb


@pytest.mark.parametrize("name,n_det,n_samp,nside", [("c1", 4, 6000, 16), ("c4", 4, 40000, 32)])
@pytest.mark.parametrize("precond_width", [20, 1])
def test_pcg_with_noise_prior(name, n_det, n_samp, nside, precond_width):
    obs, pb = _problem(name, n_det, n_samp, nside)
    rate = obs["rate"]
    step_time = pb.step_length / rate
    obstime = (pb.n_samp - 1) / rate
    psdfreq, psds = OP.analytic_psd(obs["sigma"], rate, fknee=0.05, fmin=1e-4, alpha=1.5,
                                    n_freq=300)
    oprior = OP.build_prior(psdfreq, psds, pb.det_scale, pb.offset_var, pb.n_amp_views, obstime,
                            step_time, rate, precond_width=precond_width)
    b = PP.OffsetPriorBuilder(pb.n_amp, precond_width)
    freq = PP.prior_frequencies(obstime, step_time, rate)
    for d in range(pb.n_det):
        b.add_detector(int(pb.det_start[d]), pb.n_amp_views, psdfreq, psds[d], pb.det_scale[d],
                       pb.offset_var, freq, step_time)
    prior = b.finish()

    dobs = DeviceObservation(
        focalplane=obs["focalplane"], boresight=obs["boresight"], intervals=obs["intervals"],
        det_scale=pb.det_scale, step_length=pb.step_length, nside=pb.nside, nest=pb.nest,
        n_pix_submap=pb.n_pix_submap, n_submap=pb.n_submap, global2local=pb.global2local,
        epsilon=obs["epsilon"], gamma=obs["gamma"], cal=obs["cal"],
        shared_flags=None, shared_flag_mask=0,   # (use_flags=False: no shared flags)
        solver_flags=pb.solver_flags, solver_flag_mask=pb.det_flag_mask)
    dobs.expand_pointing(np.zeros(pb.n_submap, dtype=np.uint8))
    ds = Destriper([dobs], pb.n_local_submap, pb.n_pix_submap, pb.cov, pb.offset_var,
                   pb.amp_flags, prior=prior)

    # LHS with the prior term
    rng = np.random.default_rng(2)
    a = rng.standard_normal(pb.n_amp)
    ref = O.solver_lhs(pb, O, a)
    OP.add_prior(oprior, a, pb.amp_flags, ref)
    q = torch.zeros(pb.n_amp, dtype=torch.float64, device="cuda")
    ds.lhs(torch.from_numpy(a).cuda(), q)
    assert_close_norm(q.cpu().numpy(), ref, what="LHS + prior")

    rhs_ref = O.solver_rhs(pb, O, obs["signal"])
    amps_ref, _ = O.solve(pb, O, rhs_ref, n_iter_max=3, prior=oprior)
    amps, _ = ds.solve(torch.from_numpy(rhs_ref).cuda(), n_iter_max=3)
    assert_close_norm(amps.cpu().numpy(), amps_ref, what="amplitudes after 3 iterations")
    _, hist_ref = O.solve(pb, O, rhs_ref, n_iter_max=10, prior=oprior)
    _, hist = ds.solve(torch.from_numpy(rhs_ref).cuda(), n_iter_max=10)
    H.assert_history_matches(hist, hist_ref, H.pcg_envelope(pb, rhs_ref, 10, prior=oprior),
                             what=f"{name} with prior")


@pytest.mark.parametrize("precond_width,det_flags,name,n_samp",
                         [(20, None, "c1", 6000), (1, "flags", "c2", 24000)])
@pytest.mark.parametrize("use_accel", [False, True])
def test_offset_template_with_noise_prior(precond_width, det_flags, name, n_samp, use_accel):
    """templates.Offset(use_noise_prior=True).add_prior / apply_precond -- which the reference
    can only run on the host (offset.py:888-891, 964-967) -- against the oracle restatement,
    with host arrays and with amplitudes registered in the accel table.  With the prior the
    baselines span the observation and the view only flags samples (offset.py:136-141).  The
    Toeplitz form runs on the ground scan with detector flags and turnarounds (flagged
    baselines); the banded form cannot be built with flagged baselines (the reference's
    cholesky_banded rejects 1 / offsetvar = inf), so it runs on the gap-free satellite scan."""
    from toast_b200.data import Data, NoiseModel, observation_from_synthetic
    from toast_b200.templates import Offset

    obs = S.make_observation(name, n_det=4, n_samp=n_samp, nside=64, eps_max=0.03)
    data = Data()
    ob = observation_from_synthetic(obs)
    data.obs.append(ob)
    dets = ob.local_detectors
    rate = obs["rate"]
    psdfreq, psds = OP.analytic_psd(obs["sigma"], rate, fknee=0.05, fmin=1e-4, alpha=1.5,
                                    n_freq=300)
    ob["noise_model"] = NoiseModel({d: float(w) for d, w in zip(dets, obs["detweight"])},
                                   {d: psdfreq for d in dets},
                                   {d: psds[i] for i, d in enumerate(dets)})
    tmpl = Offset(name="baselines", step_time=obs["step_time"], times="times",
                  noise_model="noise_model", det_flags=det_flags, det_flag_mask=1,
                  view="scanning", use_noise_prior=True, precond_width=precond_width)
    tmpl.initialize(data)
    nav, det_start, n_amp = O.offset_layout(4, O.make_intervals([(0, n_samp)]),
                                            obs["step_length"])
    np.testing.assert_array_equal(tmpl._obs_views[0], nav)
    assert tmpl._n_local == n_amp
    t = ob.shared["times"]
    oprior = OP.build_prior(psdfreq, psds, obs["detweight"], tmpl._offsetvar, nav,
                            float(t[-1] - t[0]), obs["step_time"], tmpl._obs_rate[0],
                            precond_width=precond_width)
    a_in, out, pre = tmpl.zeros(), tmpl.zeros(), tmpl.zeros()
    if det_flags is not None:
        a_in.local_flags[::5] = 1   # flagged baselines: outputs zeroed, inputs still convolved
    rng = np.random.default_rng(12)
    a_in.local[:] = rng.standard_normal(n_amp)
    out.local[:] = rng.standard_normal(n_amp)
    ref_add = out.local.copy()
    OP.add_prior(oprior, a_in.local, a_in.local_flags, ref_add)
    ref_pre = np.zeros(n_amp)
    OP.apply_precond(oprior, a_in.local, a_in.local_flags, ref_pre)
    if use_accel:
        for amp in (a_in, out, pre):
            amp.accel_create()
            amp.accel_update_device()
    tmpl.add_prior(a_in, out, use_accel=use_accel)
    tmpl.apply_precond(a_in, pre, use_accel=use_accel)
    if use_accel:
        for amp in (out, pre):
            amp.accel_update_host()
        for amp in (a_in, out, pre):
            amp.accel_delete()
    assert_close_norm(out.local, ref_add, rtol=1e-12, what="Offset.add_prior")
    assert_close_norm(pre.local, ref_pre, rtol=1e-12, what="Offset.apply_precond")


@pytest.mark.parametrize("chunk", [0, 8, 64, 256])
def test_partitioned_banded_solve(chunk):
    """The six-launch partitioned form of the banded preconditioner (option prior_chunk, default
    1024; 0 = one thread per segment) against the oracle."""
    from toast_b200 import lib as L

    lib = L.load()
    case = make_case(20, n_amp_views=(120, 37, 700, 5), n_det=4)
    try:
        L.check(lib.tb_set_option(b"prior_chunk", chunk))
        prior = build_product(case, cut=(2,)).finish()
    finally:
        lib.tb_set_option(b"prior_chunk", 1024)   # the default
    n, per = case["n_amp"], case["per"]
    rng = np.random.default_rng(8)
    a_in = rng.standard_normal(n)
    flags = (rng.random(n) < 0.1).astype(np.uint8)
    ref = np.zeros(n)
    OP.apply_precond(case["prior"], a_in, flags, ref)
    ref[2 * per:3 * per] = 0.0
    pre = np.full(n, 7.0)
    prior.precond(a_in, flags, pre)
    assert_close_norm(pre, ref, rtol=1e-12, what=f"partitioned precond (chunk {chunk})")
    assert np.all(pre[flags != 0] == 0.0)


def test_partitioned_banded_solve_on_a_12_hour_view():
    """The C4 shape: one view of 43 200 baselines per detector (43 chunks of 1024)."""
    case = make_case(20, n_amp_views=(43200,), n_det=2)
    prior = build_product(case).finish()
    n = case["n_amp"]
    rng = np.random.default_rng(12)
    a_in = rng.standard_normal(n)
    flags = (rng.random(n) < 0.05).astype(np.uint8)
    ref = np.zeros(n)
    OP.apply_precond(case["prior"], a_in, flags, ref)
    a_d, f_d = torch.from_numpy(a_in).cuda(), torch.from_numpy(flags).cuda()
    p_d = torch.full((n,), 7.0, dtype=torch.float64, device="cuda")
    prior.precond(a_d, f_d, p_d)
    assert_close_norm(p_d.cpu().numpy(), ref, rtol=1e-12, what="partitioned precond, 43200")
    ref_add = np.zeros(n)
    OP.add_prior(case["prior"], a_in, flags, ref_add)
    o_d = torch.zeros(n, dtype=torch.float64, device="cuda")
    prior.add(a_d, f_d, o_d)
    assert_close_norm(o_d.cpu().numpy(), ref_add, rtol=1e-12, what="add_prior, 43200")
