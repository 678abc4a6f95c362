"""CPU tests of the Offset noise prior:
  * the oracle's PSD helpers against the reference's own method bodies (golden fixture made by
    tests/golden/make_golden_prior.py from /root/reference),
  * the product's host-side assembly (toast_b200/templates/offset_prior.py) against the oracle,
  * the product's per-element kernel cores (csrc/tb_prior.cuh compiled for the host) against the
    scipy calls the reference makes (offset.py:884-1010)."""

import ctypes as ct

import numpy as np
import pytest

import helpers as H
from helpers import S
from oracle import offset_prior as OP
from toast_b200.templates import offset_prior as PP


@pytest.fixture(scope="module")
def golden():
    return np.load(f"{H.GOLDEN}/offset_prior.npz")


def test_oracle_helpers_match_reference_methods(golden):
    g = golden
    step = float(g["step_time"])
    for d in range(len(g["sigma"])):
        np.testing.assert_array_equal(OP.remove_white_noise(g["psdfreq"], g["psds"][d].copy()),
                                      g["corrpsd"][d])
        np.testing.assert_array_equal(
            OP.get_offset_psd(g["psdfreq"], g["psds"][d].copy(), g["freq"], step),
            g["offset_psd"][d])
    np.testing.assert_array_equal(
        OP.interpolate_psd(g["interp_x"], np.log(g["freq"]), np.log(g["offset_psd"][0])),
        g["interp_y"])
    np.testing.assert_array_equal(OP.truncate(g["filter_raw"].copy()), g["filter_truncated"])
    np.testing.assert_array_equal(OP.truncate(g["filter_raw"].copy(), lim=1e-2),
                                  g["filter_truncated_1e2"])
    np.testing.assert_array_equal(
        OP.prior_frequencies(float(g["obstime"]), step, float(g["rate"])), g["freq"])
    assert OP.prior_frequencies(1.0, 2.0, 10.0) is None  # a single baseline: prior disabled


def test_product_psd_helpers_match_reference_methods(golden):
    g = golden
    step = float(g["step_time"])
    for d in range(len(g["sigma"])):
        np.testing.assert_array_equal(PP.offset_psd(g["psdfreq"], g["psds"][d], g["freq"], step),
                                      g["offset_psd"][d])
    np.testing.assert_array_equal(PP._symmetric_cut(g["filter_raw"].copy()),
                                  g["filter_truncated"])
    np.testing.assert_array_equal(PP._symmetric_cut(g["filter_raw"].copy(), lim=1e-2),
                                  g["filter_truncated_1e2"])
    np.testing.assert_array_equal(
        PP.prior_frequencies(float(g["obstime"]), step, float(g["rate"])), g["freq"])


def make_case(precond_width, n_amp_views=(120, 37, 143, 5), n_det=3, seed=0):
    """Three detectors, four views (one shorter than the band), random amplitude variance."""
    rate, step_time, obstime = 10.0, 2.0, 600.0
    rng = np.random.default_rng(seed)
    sigma = 1.0 + 0.1 * rng.random(n_det)
    psdfreq, psds = OP.analytic_psd(sigma, rate, fknee=0.05, fmin=1e-4, alpha=1.5, n_freq=300)
    nav = np.array(n_amp_views, dtype=np.int64)
    per = int(nav.sum())
    var = 1.0 / rng.uniform(10, 20, size=n_det * per)
    detnoise = 1.0 / sigma**2
    prior = OP.build_prior(psdfreq, psds, detnoise, var, nav, obstime, step_time, rate,
                           precond_width=precond_width)
    return dict(rate=rate, step_time=step_time, obstime=obstime, psdfreq=psdfreq, psds=psds,
                nav=nav, per=per, var=var, detnoise=detnoise, n_det=n_det, n_amp=n_det * per,
                prior=prior, precond_width=precond_width)


def build_product(case, cut=()):
    b = PP.OffsetPriorBuilder(case["n_amp"], case["precond_width"])
    freq = PP.prior_frequencies(case["obstime"], case["step_time"], case["rate"])
    for d in range(case["n_det"]):
        if d in cut:
            b.add_cut_detector(d * case["per"], case["nav"])
        else:
            b.add_detector(d * case["per"], case["nav"], case["psdfreq"], case["psds"][d],
                           case["detnoise"][d], case["var"], freq, case["step_time"])
    return b


@pytest.mark.parametrize("precond_width", [20, 4, 1])
def test_product_assembly_matches_oracle(precond_width):
    case = make_case(precond_width)
    b = build_product(case)
    k = 0
    for d in range(case["n_det"]):
        for v in range(len(case["nav"])):
            np.testing.assert_array_equal(b.filters[k], case["prior"].filters[d][v])
            np.testing.assert_array_equal(
                b.precond[k], np.asarray(case["prior"].precond[d][v][0]).reshape(-1))
            assert b.seg_start[k] == d * case["per"] + int(case["nav"][:v].sum())
            k += 1
    assert b._nf == sum(f.size for f in b.filters) and b._np == sum(p.size for p in b.precond)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return ct.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("precond_width", [20, 4, 1])
def test_kernel_cores_match_scipy(precond_width):
    """tb_prior.cuh (conv_same_at, banded_cho_solve) looped over the segments exactly as the
    kernels do, against scipy.signal.convolve / cho_solve_banded through the oracle."""
    hm = H.host_math_lib()
    case = make_case(precond_width)
    b = build_product(case)
    n = case["n_amp"]
    rng = np.random.default_rng(3)
    a_in = rng.standard_normal(n)
    flags = (rng.random(n) < 0.1).astype(np.uint8)
    seg_start, seg_len = _i64(b.seg_start), _i64(b.seg_len)
    f_start, f_len = _i64(b.filt_start), _i64(b.filt_len)
    p_start, p_width = _i64(b.prec_start), _i64(b.prec_width)
    taps, pre = np.concatenate(b.filters), np.concatenate(b.precond)

    # add_prior accumulates onto what is already there
    out = rng.standard_normal(n)
    ref = out.copy()
    OP.add_prior(case["prior"], a_in, flags, ref)
    hm.tbp_conv_segments(ct.c_int64(len(seg_start)), _ptr(seg_start), _ptr(seg_len),
                         _ptr(f_start), _ptr(f_len), _ptr(taps), _ptr(a_in), _ptr(flags),
                         _ptr(out), ct.c_int(0))
    H.assert_close_norm(out, ref, rtol=1e-13, what="add_prior")
    assert np.all(out[flags != 0] == 0.0)

    ref = np.full(n, np.nan)
    OP.apply_precond(case["prior"], a_in, flags, ref)
    out = np.full(n, np.nan)
    if precond_width == 1:
        hm.tbp_conv_segments(ct.c_int64(len(seg_start)), _ptr(seg_start), _ptr(seg_len),
                             _ptr(p_start), _ptr(p_width), _ptr(pre), _ptr(a_in), _ptr(flags),
                             _ptr(out), ct.c_int(1))
    else:
        hm.tbp_banded_segments(ct.c_int64(len(seg_start)), _ptr(seg_start), _ptr(seg_len),
                               _ptr(p_start), _ptr(p_width), _ptr(pre), _ptr(a_in), _ptr(flags),
                               _ptr(out))
    H.assert_close_norm(out, ref, rtol=1e-12, what="apply_precond")
    assert np.all(out[flags != 0] == 0.0)


def test_filter_longer_than_the_segment_and_cut_detectors():
    """"same" convolution keeps the length of the AMPLITUDES even when the filter is longer
    (5-baseline view, 35-tap filter); a cut detector comes out as zeros."""
    hm = H.host_math_lib()
    case = make_case(20, n_amp_views=(5, 3, 1, 64))
    b = build_product(case, cut=(1,))
    n = case["n_amp"]
    rng = np.random.default_rng(4)
    a_in = rng.standard_normal(n)
    flags = np.zeros(n, dtype=np.uint8)
    ref = np.zeros(n)
    OP.add_prior(case["prior"], a_in, flags, ref)
    per = case["per"]
    ref[per:2 * per] = 0.0
    out = np.zeros(n)
    seg_start, seg_len = _i64(b.seg_start), _i64(b.seg_len)
    f_start, f_len = _i64(b.filt_start), _i64(b.filt_len)
    taps = np.concatenate(b.filters)
    assert min(seg_len) == 1 and max(f_len) > 5
    hm.tbp_conv_segments(ct.c_int64(len(seg_start)), _ptr(seg_start), _ptr(seg_len),
                         _ptr(f_start), _ptr(f_len), _ptr(taps), _ptr(a_in), _ptr(flags),
                         _ptr(out), ct.c_int(0))
    H.assert_close_norm(out, ref, rtol=1e-13, what="short segments")
    assert np.all(out[per:2 * per] == 0.0)


# ---- the reference's own Offset methods, executed from its source --------------------------------
# tests/golden/make_golden_offset_init.py runs Offset._initialize / _add_prior / _apply_precond of
# /root/reference with duck-typed stand-ins; the oracle restatements reproduce them bit for bit.
OFFSET_CASES = [("plain_c2", "c2", False, 20, True), ("plain_c1", "c1", False, 20, True),
                ("prior_banded_c1", "c1", True, 20, False),
                ("prior_banded4_c4", "c4", True, 4, False),
                ("prior_toeplitz_c2", "c2", True, 1, True)]


def reference_offset_case(tag, name, prior, use_det_flags):
    """Inputs of one fixture case, rebuilt exactly as the generator built them, and the oracle's
    layout / variance for them."""
    from helpers import O, S

    g = np.load(f"{H.GOLDEN}/offset_template.npz")
    n_det, n_samp = [int(x) for x in g[f"{tag}_args"]]
    obs = S.make_observation(name, n_det=n_det, n_samp=n_samp, nside=64, eps_max=0.03)
    # with a noise prior the baselines span the observation, the view only flags samples
    # (offset.py:136-141)
    bounds = O.make_intervals([(0, n_samp)]) if prior else obs["intervals"]
    nav, det_start, n_amp = O.offset_layout(n_det, bounds, obs["step_length"])
    in_view = np.zeros(n_samp, dtype=bool)
    for v in obs["intervals"]:
        in_view[v["first"]:v["last"]] = True
    sf = np.zeros((n_det, n_samp), dtype=np.uint8)
    if use_det_flags:
        sf |= obs["det_flags"] & 1
    sf |= (~in_view).astype(np.uint8)[None, :]
    var, fl = O.offset_variance(n_det, n_samp, bounds, obs["step_length"], nav, obs["detweight"],
                                sf, 1)
    return g, obs, bounds, nav, det_start, n_amp, sf, var, fl


@pytest.mark.parametrize("tag,name,prior,precond_width,use_det_flags", OFFSET_CASES)
def test_oracle_offset_glue_is_bit_identical_to_reference_methods(tag, name, prior,
                                                                  precond_width, use_det_flags):
    g, obs, bounds, nav, det_start, n_amp, sf, var, fl = reference_offset_case(
        tag, name, prior, use_det_flags)
    np.testing.assert_array_equal(nav, g[f"{tag}_n_amp_views"])
    np.testing.assert_array_equal(det_start, g[f"{tag}_det_start"])
    np.testing.assert_array_equal(fl, g[f"{tag}_amp_flags"])
    np.testing.assert_array_equal(var, g[f"{tag}_offset_var"])
    if not prior:
        return
    # the rate the template derives from the timestamps (1 / median dt: 9.999999999999858 Hz)
    n_samp, rate = obs["n_samp"], float(g[f"{tag}_rate"])
    assert abs(rate - obs["rate"]) < 1e-9
    assert rate == 1.0 / np.median(np.diff(np.arange(n_samp, dtype=np.float64) / obs["rate"]))
    psdfreq, psds = OP.analytic_psd(obs["sigma"], obs["rate"], fknee=0.05, fmin=1e-4, alpha=1.5,
                                    n_freq=300)
    t = np.arange(n_samp, dtype=np.float64) / obs["rate"]
    obstime = float(t[-1] - t[0])
    np.testing.assert_array_equal(
        OP.prior_frequencies(obstime, float(obs["step_time"]), rate), g[f"{tag}_freq"])
    pr = OP.build_prior(psdfreq, psds, obs["detweight"], var, nav, obstime,
                        float(obs["step_time"]), rate, precond_width=precond_width)
    for i in range(obs["n_det"]):
        for v in range(len(nav)):
            np.testing.assert_array_equal(pr.filters[i][v], g[f"{tag}_filter_{i}_{v}"])
            np.testing.assert_array_equal(np.asarray(pr.precond[i][v][0]),
                                          g[f"{tag}_precond_{i}_{v}"])
    out = g[f"{tag}_amps_out0"].copy()
    OP.add_prior(pr, g[f"{tag}_amps_in"], g[f"{tag}_flags_in"], out)
    np.testing.assert_array_equal(out, g[f"{tag}_add_prior"])
    pre = np.zeros_like(out)
    OP.apply_precond(pr, g[f"{tag}_amps_in"], g[f"{tag}_flags_in"], pre)
    np.testing.assert_array_equal(pre, g[f"{tag}_apply_precond"])

    # the product's assembly and kernel cores on the same case
    b = PP.OffsetPriorBuilder(n_amp, precond_width)
    freq = PP.prior_frequencies(obstime, float(obs["step_time"]), rate)
    for d in range(obs["n_det"]):
        b.add_detector(int(det_start[d]), nav, psdfreq, psds[d], obs["detweight"][d], var, freq,
                       float(obs["step_time"]))
    k = 0
    for i in range(obs["n_det"]):
        for v in range(len(nav)):
            np.testing.assert_array_equal(b.filters[k], g[f"{tag}_filter_{i}_{v}"])
            np.testing.assert_array_equal(b.precond[k], g[f"{tag}_precond_{i}_{v}"].reshape(-1))
            k += 1
    hm = H.host_math_lib()
    seg_start, seg_len = _i64(b.seg_start), _i64(b.seg_len)
    f_start, f_len = _i64(b.filt_start), _i64(b.filt_len)
    p_start, p_width = _i64(b.prec_start), _i64(b.prec_width)
    taps, pre_v = np.concatenate(b.filters), np.concatenate(b.precond)
    a_in = np.ascontiguousarray(g[f"{tag}_amps_in"])
    flags = np.ascontiguousarray(g[f"{tag}_flags_in"])
    out = g[f"{tag}_amps_out0"].copy()
    hm.tbp_conv_segments(ct.c_int64(len(seg_start)), _ptr(seg_start), _ptr(seg_len),
                         _ptr(f_start), _ptr(f_len), _ptr(taps), _ptr(a_in), _ptr(flags),
                         _ptr(out), ct.c_int(0))
    H.assert_close_norm(out, g[f"{tag}_add_prior"], rtol=1e-13, what="add_prior core")
    res = np.zeros_like(out)
    if precond_width == 1:
        hm.tbp_conv_segments(ct.c_int64(len(seg_start)), _ptr(seg_start), _ptr(seg_len),
                             _ptr(p_start), _ptr(p_width), _ptr(pre_v), _ptr(a_in), _ptr(flags),
                             _ptr(res), ct.c_int(1))
    else:
        hm.tbp_banded_segments(ct.c_int64(len(seg_start)), _ptr(seg_start), _ptr(seg_len),
                               _ptr(p_start), _ptr(p_width), _ptr(pre_v), _ptr(a_in),
                               _ptr(flags), _ptr(res))
    H.assert_close_norm(res, g[f"{tag}_apply_precond"], rtol=1e-12, what="apply_precond core")


def _numpy_project_batch(data_index, det_data, flag_index, flag_data, flag_mask, step_length,
                         amp_offsets, n_amp_views, amplitudes, amplitude_flags, intervals):
    """numpy stand-in of the CUDA kernel behind kernels.template_offset_project_signal_batch
    (template_offset.cpp:243-327), so that the template's HOST logic can run without a GPU."""
    view_off = np.concatenate([[0], np.cumsum(n_amp_views)[:-1]])
    for k, row in enumerate(data_index):
        for v, iv in enumerate(intervals):
            first, last = int(iv["first"]), int(iv["last"])
            s = np.arange(first, last)
            amp = int(amp_offsets[k]) + int(view_off[v]) + (s - first) // int(step_length)
            good = np.ones(len(s), dtype=bool)
            if flag_data is not None:
                good = (flag_data[int(flag_index[k]), first:last] & flag_mask) == 0
            good &= amplitude_flags[amp] == 0
            np.add.at(amplitudes, amp[good], det_data[int(row), first:last][good])


@pytest.mark.parametrize("tag,name,prior,precond_width,use_det_flags", OFFSET_CASES)
def test_template_mirror_host_logic_matches_reference_initialize(tag, name, prior,
                                                                 precond_width, use_det_flags,
                                                                 monkeypatch):
    """templates.Offset._initialize of the product (layout, amplitude flags, variance, and --
    with use_noise_prior -- baselines spanning the observation plus the filter / preconditioner
    assembly) against the reference's own _initialize, the CUDA projection kernel replaced by a
    numpy stand-in."""
    from toast_b200.data import Data, NoiseModel, observation_from_synthetic
    from toast_b200.templates import Offset
    from toast_b200.templates import offset as offset_module

    monkeypatch.setattr(offset_module.KC, "template_offset_project_signal_batch",
                        _numpy_project_batch)
    g, obs, bounds, nav, det_start, n_amp, sf, var, fl = reference_offset_case(
        tag, name, prior, use_det_flags)
    data = Data()
    ob = observation_from_synthetic(obs)
    data.obs.append(ob)
    dets = ob.local_detectors
    psdfreq, psds = OP.analytic_psd(obs["sigma"], obs["rate"], fknee=0.05, fmin=1e-4, alpha=1.5,
                                    n_freq=300)
    ob["noise_model"] = NoiseModel({d: float(w) for d, w in zip(dets, obs["detweight"])},
                                   {d: psdfreq for d in dets},
                                   {d: psds[i] for i, d in enumerate(dets)})
    tmpl = Offset(name="baselines", step_time=float(obs["step_time"]), times="times",
                  noise_model="noise_model", det_flags="flags" if use_det_flags else None,
                  det_flag_mask=1, view="scanning", use_noise_prior=prior,
                  precond_width=precond_width)
    tmpl._defer_prior = True   # OffsetPrior.finish() uploads to the device
    tmpl.initialize(data)
    np.testing.assert_array_equal(tmpl._obs_views[0], g[f"{tag}_n_amp_views"])
    np.testing.assert_array_equal([tmpl._det_start[d] for d in dets], g[f"{tag}_det_start"])
    assert tmpl._obs_rate[0] == float(g[f"{tag}_rate"])
    np.testing.assert_array_equal(tmpl._amp_flags.astype(np.uint8), g[f"{tag}_amp_flags"])
    np.testing.assert_array_equal(tmpl._offsetvar, g[f"{tag}_offset_var"])
    if not prior:
        return
    b = tmpl._prior_builder(data)
    k = 0
    for i in range(obs["n_det"]):
        for v in range(len(nav)):
            np.testing.assert_array_equal(b.filters[k], g[f"{tag}_filter_{i}_{v}"])
            np.testing.assert_array_equal(b.precond[k], g[f"{tag}_precond_{i}_{v}"].reshape(-1))
            assert b.seg_start[k] == det_start[i] + int(nav[:v].sum()) and b.seg_len[k] == nav[v]
            k += 1


@pytest.mark.parametrize("n,w,m", [(800, 20, 64), (800, 20, 19), (431, 4, 50), (60, 20, 32),
                                   (37, 20, 19), (5000, 20, 256), (300, 1, 16),
                                   (43200, 20, 256), (43200, 20, 1024), (5000, 20, 1024),
                                   (1024, 20, 1024), (1025, 20, 1024)])
def test_partitioned_banded_solve_matches_the_sequential_one(n, w, m):
    """The chunk-parallel form of cho_solve_banded (tb_prior.cuh: fwd/bwd_chunk, fwd/bwd_response,
    chunk_correct; device wiring is the next step) against scipy, for chunk sizes down to the
    minimum w - 1, a last chunk shorter than the band, and a single chunk."""
    import scipy.linalg

    hm = H.host_math_lib()
    rng = np.random.default_rng(n + w)
    # an SPD banded matrix like the preconditioner's: diag(1 / var) + Toeplitz(filter)
    lags = np.exp(-np.arange(w) / 3.0) * np.where(np.arange(w) % 2 == 0, 1.0, -0.7)
    ab = np.zeros((w, n))
    ab[:] = lags[:, None]
    ab[0] += rng.uniform(2.0, 20.0, size=n)
    for k in range(1, w):
        ab[k, n - k:] = 0.0
    cb = scipy.linalg.cholesky_banded(ab.copy(), lower=True)
    b = rng.standard_normal(n)
    ref = scipy.linalg.cho_solve_banded((cb, True), b)
    cbc = np.ascontiguousarray(cb)
    x = np.zeros(n)
    hm.tbp_banded_partitioned(_ptr(cbc), ct.c_int64(w), ct.c_int64(n), ct.c_int64(m), _ptr(b),
                              _ptr(x))
    H.assert_close_norm(x, ref, rtol=1e-12, what=f"partitioned solve n={n} w={w} m={m}")


@pytest.mark.parametrize("chunk", [8, 19, 64, 1000, 1024])
def test_partitioned_segments_match_the_reference_preconditioner(chunk):
    """The per-thread functions of the partitioned solve (pb_* in tb_prior.cuh), run over all
    segments in the order of the six launches, against Offset._apply_precond on the banded
    cases: different widths, views shorter than the band, a cut detector, flagged amplitudes."""
    hm = H.host_math_lib()
    # (the third case straddles the default chunk length: 1024 / 1025 / several chunks of 1024)
    for pw, views in ((20, (120, 37, 700, 5)), (4, (50, 3, 1, 64)), (20, (1024, 1025, 3000))):
        case = make_case(pw, n_amp_views=views, n_det=3)
        b = build_product(case, cut=(1,))
        n = case["n_amp"]
        per = case["per"]
        rng = np.random.default_rng(pw + chunk)
        a_in = rng.standard_normal(n)
        flags = (rng.random(n) < 0.1).astype(np.uint8)
        ref = np.zeros(n)
        OP.apply_precond(case["prior"], a_in, flags, ref)
        ref[per:2 * per] = 0.0
        seg_start, seg_len = _i64(b.seg_start), _i64(b.seg_len)
        p_start, p_width = _i64(b.prec_start), _i64(b.prec_width)
        pre = np.concatenate(b.precond)
        out = np.full(n, np.nan)
        hm.tbp_banded_partitioned_segments(ct.c_int64(len(seg_start)), _ptr(seg_start),
                                           _ptr(seg_len), _ptr(p_start), _ptr(p_width),
                                           _ptr(pre), ct.c_int64(chunk), _ptr(a_in), _ptr(flags),
                                           _ptr(out))
        H.assert_close_norm(out, ref, rtol=1e-12, what=f"partitioned precond pw={pw}")
        assert np.all(out[flags != 0] == 0.0) and np.all(out[per:2 * per] == 0.0)


def test_single_baseline_observation_disables_the_prior(monkeypatch):
    """offset.py:208-214: an observation shorter than one baseline gets no noise prior -- its
    amplitudes come out of add_prior / apply_precond as zeros (offset.py:946-947, 1003-1005),
    which the product expresses as cut segments."""
    from toast_b200.data import Data, NoiseModel, observation_from_synthetic
    from toast_b200.templates import Offset
    from toast_b200.templates import offset as offset_module

    monkeypatch.setattr(offset_module.KC, "template_offset_project_signal_batch",
                        _numpy_project_batch)
    obs = S.make_observation("c1", n_det=2, n_samp=80, nside=64)   # 8 s at 10 Hz
    data = Data()
    ob = observation_from_synthetic(obs)
    data.obs.append(ob)
    dets = ob.local_detectors
    psdfreq, psds = OP.analytic_psd(obs["sigma"], obs["rate"], n_freq=100)
    ob["noise_model"] = NoiseModel({d: float(w) for d, w in zip(dets, obs["detweight"])},
                                   {d: psdfreq for d in dets},
                                   {d: psds[i] for i, d in enumerate(dets)})
    tmpl = Offset(name="baselines", step_time=10.0, times="times", noise_model="noise_model",
                  det_flags=None, view="scanning", use_noise_prior=True)
    tmpl._defer_prior = True
    tmpl.initialize(data)
    assert tmpl._n_local == 2 and list(tmpl._obs_views[0]) == [1]
    assert PP.prior_frequencies(7.9, 10.0, 10.0) is None
    b = tmpl._prior_builder(data)
    assert b.filt_start == [-1, -1] and b.prec_start == [-1, -1]
    assert b.seg_start == [0, 1] and b.seg_len == [1, 1]
    # the cut segments zero their amplitudes (kernel cores on the host)
    hm = H.host_math_lib()
    out = np.array([3.0, 4.0])
    a_in = np.array([1.0, 2.0])
    flags = np.zeros(2, dtype=np.uint8)
    ss, sl, fs, fl = (_i64(v) for v in (b.seg_start, b.seg_len, b.filt_start, b.filt_len))
    taps = np.zeros(1)
    hm.tbp_conv_segments(ct.c_int64(2), _ptr(ss), _ptr(sl), _ptr(fs), _ptr(fl), _ptr(taps),
                         _ptr(a_in), _ptr(flags), _ptr(out), ct.c_int(0))
    assert np.all(out == 0.0)
