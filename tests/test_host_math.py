"""CPU tests of the PRODUCT's per-sample arithmetic (toast_b200/csrc/tb_math.cuh compiled for
the host by tests/csrc/host_math.cpp): the correctly-rounded atan2, the two-tier pixel path and
the trig-free Stokes weights, against glibc and the oracle."""

import ctypes as ct
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from helpers import O, S

CSRC = os.path.join(H.ROOT, "tests", "csrc")


@pytest.fixture(scope="module")
def hm():
    return H.host_math_lib()


def _p(a):
    return ct.c_void_p(a.ctypes.data)


def test_atan2_cr_is_correctly_rounded_and_matches_glibc(hm):
    import math

    rng = np.random.default_rng(1)
    n = 400000
    y = rng.standard_normal(n)
    x = rng.standard_normal(n)
    y[:100] = 0.0
    x[100:200] = 0.0
    y[200:300] *= 1e-12
    x[300:400] *= 1e-12
    out = np.zeros(n)
    hm.tbm_atan2_cr(ct.c_int64(n), _p(y), _p(x), _p(out))
    libm = np.array([math.atan2(a, b) for a, b in zip(y, x)])  # glibc, as the reference uses
    ulp = np.spacing(np.abs(libm))
    assert np.all(np.abs(out - libm) <= ulp)            # never more than 1 ulp apart
    assert np.mean(out != libm) < 3e-3                  # glibc 2.39 is not correctly rounded
    # spot-check correct rounding with mpmath on the disagreements
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    for i in np.flatnonzero(out != libm)[:40]:
        t = mp.atan2(mp.mpf(float(y[i])), mp.mpf(float(x[i])))
        assert abs(mp.mpf(float(out[i])) - t) <= abs(mp.mpf(float(libm[i])) - t)
    # axes
    assert out[0] in (0.0, math.pi)


@pytest.mark.parametrize("nest", [1, 0])
@pytest.mark.parametrize("nside", [1, 64, 2048, 1 << 16])
def test_two_tier_pixel_path_matches_oracle(hm, nside, nest):
    rng = np.random.default_rng(5)
    n = 200000
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    s = np.sqrt(0.5)
    q[:9] = [[0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [s, 0, 0, s], [0, s, 0, s],
             [0, 0, s, s], [s, s, 0, 0], [0.5, 0.5, 0.5, 0.5]]
    n_submap, nps = S.n_submap_for(nside, 16)
    idx = np.zeros(1, dtype=np.int32)
    ref = np.zeros((1, n), dtype=np.int64)
    O.pixels_healpix(idx, q.reshape(1, n, 4), np.zeros(1, dtype=np.uint8), 0, idx, ref,
                     S.make_intervals([(0, n)]), np.zeros(n_submap, dtype=np.uint8), nps, nside,
                     nest)
    for guard, all_exact in ((1.0, False), (0.0, False), (1e20, True)):
        pix = np.zeros(n, dtype=np.int64)
        n_exact = hm.tbm_quat2pix(ct.c_int64(n), _p(q), ct.c_int64(nside), ct.c_int(nest),
                                  ct.c_double(guard), _p(pix))
        np.testing.assert_array_equal(pix, ref[0])
        assert (n_exact == n) if all_exact else (n_exact < 50)


def test_guard_band_flags_every_sample_a_one_ulp_phi_change_can_move(hm):
    """If nudging phi by +-2 ulp changes the pixel, the sample must have been flagged ambiguous
    (that is what routes it to the exact path on the GPU)."""
    rng = np.random.default_rng(7)
    n = 300000
    nside = 2048
    # adversarial samples: phi chosen so that nside*(1/2 + tt) - 3/4 nside z sits within
    # rounding of an integer (random directions essentially never do: P ~ 1e-10 per sample)
    z = rng.uniform(-0.66, 0.66, n)
    j = rng.integers(0, 4 * nside, n).astype(np.float64)
    tt = np.mod((j + 0.75 * nside * z - 0.5 * nside) / nside, 4.0)
    phi = tt * (np.pi / 2)
    phi = np.where(phi > np.pi, phi - 2 * np.pi, phi)
    # half of them generic
    z[: n // 2] = rng.uniform(-1, 1, n // 2)
    phi[: n // 2] = rng.uniform(-np.pi, np.pi, n // 2)
    for nest in (1, 0):
        base = np.zeros(n, dtype=np.int64)
        amb = np.zeros(n, dtype=np.uint8)
        hm.tbm_zphi2pix(ct.c_int64(n), _p(z), _p(phi), ct.c_int64(nside), ct.c_int(nest), _p(base),
                        _p(amb))
        moved = np.zeros(n, dtype=bool)
        for k in (-2, -1, 1, 2):
            ph2 = phi.copy()
            for _ in range(abs(k)):
                ph2 = np.nextafter(ph2, np.inf if k > 0 else -np.inf)
            p2 = np.zeros(n, dtype=np.int64)
            a2 = np.zeros(n, dtype=np.uint8)
            hm.tbm_zphi2pix(ct.c_int64(n), _p(z), _p(ph2), ct.c_int64(nside), ct.c_int(nest),
                            _p(p2), _p(a2))
            moved |= p2 != base
        assert moved.sum() > 1000
        assert np.all(amb[moved] == 1)
        assert amb[: n // 2].mean() < 1e-3


@pytest.mark.parametrize("hwp", [False, True])
def test_trig_free_stokes_weights_match_libm_chain(hm, hwp):
    obs = S.make_observation("c4", n_det=4, n_samp=100000, with_signal=False, eps_max=0.05)
    idx = np.arange(4, dtype=np.int32)
    iv = S.make_intervals([(0, 100000)])
    q = np.zeros((4, 100000, 4))
    O.pointing_detector(obs["focalplane"], obs["boresight"], idx, q, iv, np.zeros(1, np.uint8), 0)
    ang = (np.arange(100000) * 0.013) % 6.2 if hwp else None
    for d in range(4):
        w_ref = np.zeros((1, 100000, 3))
        O.stokes_weights_IQU(idx[:1], q[d:d + 1], idx[:1], w_ref, ang if hwp else np.zeros(1), iv,
                             obs["epsilon"][d:d + 1], np.array([0.2]), np.array([1.3]), True)
        w = np.zeros((100000, 3))
        hm.tbm_stokes_iqu(ct.c_int64(100000), _p(np.ascontiguousarray(q[d])), ct.c_double(1.3),
                          ct.c_double(obs["epsilon"][d]), ct.c_double(-1.0), ct.c_double(0.2),
                          _p(ang) if hwp else None, _p(w))
        assert np.max(np.abs(w - w_ref[0])) < 2e-14  # 1e-10 bar, four orders of margin


@pytest.mark.parametrize("name,n_det,n_samp,nside,threshold",
                         [("c1", 4, 6000, 64, 1e-3), ("c2", 6, 24000, 64, 1e-3),
                          ("c4", 8, 40000, 128, 1e-8), ("c5", 16, 30000, 256, 1e-8)])
def test_device_eigen_inverse_matches_the_lapack_restatement(hm, name, n_det, n_samp, nside,
                                                             threshold):
    """k_cov_invert3's per-pixel code (tbm::cov_invert3: cyclic Jacobi) on the host against the
    oracle's restatement of toast_map_cov.cpp:246-396 (eigh): the same pixels kept, rcond and
    the inverse to 1e-12 at the round-1 threshold 1e-3 -- two orders inside the 1e-10 bar the
    GPU tests hold the kernel to -- and to the conditioning bound eps / rcond at the production
    threshold 1e-8 (condition numbers up to 1e8: the inverse is only defined that far)."""
    obs = S.make_observation(name, n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03)
    pb = O.build_problem(obs, O, rcond_threshold=threshold)
    npix = pb.invcov.size // 6
    cov_ref, rc_ref = pb.invcov.copy(), np.zeros(npix)
    O.cov_eigendecompose_diag(1, npix, 3, cov_ref, rc_ref, threshold, True)
    cov, rc = pb.invcov.copy(), np.zeros(npix)
    hm.tbm_cov_invert3(ct.c_int64(npix), _p(cov), _p(rc), ct.c_double(threshold))
    kept = rc_ref > 0
    assert kept.sum() >= 10
    np.testing.assert_array_equal(rc > 0, kept)
    assert np.all(cov.reshape(npix, 6)[~kept] == 0.0)
    bar = max(1e-12, np.finfo(np.float64).eps / threshold)
    assert np.max(np.abs(rc[kept] - rc_ref[kept]) / rc_ref[kept]) <= bar
    H.assert_close_norm(cov, cov_ref, rtol=bar, what="inverse covariance (device code on host)")
