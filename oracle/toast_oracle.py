"""oracle/toast_oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for the plain-C restatement in ``oracle/toast_oracle.c`` plus numpy
restatements of the reference's *Python* glue on the hot path (Offset amplitude layout,
PixelDistribution, the PCG ``solve()`` loop, ``SolverLHS``/``SolverRHS``).  Used only by
``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``.
The product package ``toast_b200`` never imports this module.

Parity status: PINNED against the reference's compiled kernels (``oracle/_ref``) by
``tests/test_oracle.py`` and the committed fixtures in ``tests/golden``; the PCG loop ``solve``
is pinned bit for bit against the reference's own ``solve()`` executed from its source
(``tests/golden/make_golden_solve.py``), ``cov_eigendecompose_diag`` against the reference's
LAPACK-based implementation (``tests/golden/make_golden_cov.py``).

The keyword names and argument order of the kernel wrappers follow the reference's
``_libtoast`` signatures so a test can call either this module,
``oracle/_ref`` or the CUDA library with the same argument tuple.
"""

import ctypes as ct
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "liboracle.so")
_SRC_PATH = os.path.join(HERE, "toast_oracle.c")

# intervals.py:26-45 / intervals.hpp:10-15
interval_dtype = np.dtype(
    {
        "names": ["start", "stop", "first", "last"],
        "formats": ["d", "d", "q", "q"],
        "offsets": [0, 8, 16, 24],
    }
)


def make_intervals(ranges):
    """Build an Interval array from [(first, last_exclusive), ...]."""
    iv = np.zeros(len(ranges), dtype=interval_dtype)
    for i, (a, b) in enumerate(ranges):
        iv[i] = (float(a), float(b), int(a), int(b))
    return iv


def build(force=False):
    """Compile the C restatement (gcc, no FMA contraction)."""
    if (
        not force
        and os.path.exists(_LIB_PATH)
        and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC_PATH)
    ):
        return _LIB_PATH
    cmd = [
        "gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
        "-o", _LIB_PATH, _SRC_PATH, "-lm",
    ]
    subprocess.check_call(cmd)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ct.CDLL(_LIB_PATH)
    return _lib


def load_ref():
    """Import the compiled REFERENCE kernels (oracle/_ref), or return None."""
    d = os.path.join(HERE, "_ref")
    if not os.path.isdir(d):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import _toast_oracle as m
    except ImportError:
        return None
    if not getattr(m, "_tb_assigned", False):
        # OmpManager::get_device throws until a device is assigned (accelerator.cpp:308-317)
        m.accel_assign_device(1, 0, 1.0, True)
        m._tb_assigned = True
    return m


def _p(a, dtype=None):
    if a is None:
        return None
    if dtype is not None:
        assert a.dtype == np.dtype(dtype), (a.dtype, dtype)
    assert a.flags["C_CONTIGUOUS"]
    return ct.c_void_p(a.ctypes.data)


def _opt_flags(flags, n_samp):
    """Optional arrays are signalled by shape (len != n_samp => absent), as in the reference."""
    if flags is None or flags.shape[-1] != n_samp:
        return None
    return flags


i64 = ct.c_int64
u8 = ct.c_uint8
i32 = ct.c_int32
f64 = ct.c_double
cint = ct.c_int


def pointing_detector(focalplane, boresight, quat_index, quats, intervals, shared_flags,
                      shared_flag_mask, use_accel=False):
    n_det = len(quat_index)
    n_samp = boresight.shape[0]
    fl = _opt_flags(shared_flags, n_samp)
    lib().tbo_pointing_detector(
        _p(focalplane, "f8"), _p(boresight, "f8"), _p(quat_index, "i4"), _p(quats, "f8"),
        _p(intervals), i64(len(intervals)), _p(fl, "u1"), u8(shared_flag_mask), i64(n_det),
        i64(n_samp))


def pixels_healpix(quat_index, quats, shared_flags, shared_flag_mask, pixel_index, pixels,
                   intervals, hit_submaps, n_pix_submap, nside, nest, use_accel=False):
    n_det = len(quat_index)
    n_samp = quats.shape[1]
    fl = _opt_flags(shared_flags, n_samp)
    lib().tbo_pixels_healpix(
        _p(quat_index, "i4"), _p(quats, "f8"), _p(fl, "u1"), u8(shared_flag_mask),
        _p(pixel_index, "i4"), _p(pixels, "i8"), _p(intervals), i64(len(intervals)),
        _p(hit_submaps, "u1"), i64(n_pix_submap), i64(nside), cint(1 if nest else 0),
        i64(n_det), i64(n_samp))


def stokes_weights_IQU(quat_index, quats, weight_index, weights, hwp, intervals, epsilon,
                       gamma, cal, IAU, use_accel=False):
    n_det = len(quat_index)
    n_samp = quats.shape[1]
    h = hwp if (hwp is not None and hwp.shape[0] == n_samp) else None
    lib().tbo_stokes_weights_IQU(
        _p(quat_index, "i4"), _p(quats, "f8"), _p(weight_index, "i4"), _p(weights, "f8"),
        _p(h, "f8"), _p(intervals), i64(len(intervals)), _p(epsilon, "f8"), _p(gamma, "f8"),
        _p(cal, "f8"), cint(1 if IAU else 0), i64(n_det), i64(n_samp))


def stokes_weights_I(weight_index, weights, intervals, cal, use_accel=False):
    n_det = len(weight_index)
    n_samp = weights.shape[1]
    lib().tbo_stokes_weights_I(
        _p(weight_index, "i4"), _p(weights, "f8"), _p(intervals), i64(len(intervals)),
        _p(cal, "f8"), i64(n_det), i64(n_samp))


def noise_weight(det_data, data_index, intervals, detector_weights, use_accel=False):
    n_det = len(data_index)
    n_samp = det_data.shape[1]
    lib().tbo_noise_weight(
        _p(det_data, "f8"), _p(data_index, "i4"), _p(intervals), i64(len(intervals)),
        _p(detector_weights, "f8"), i64(n_det), i64(n_samp))


def build_noise_weighted(global2local, zmap, pixel_index, pixels, weight_index, weights,
                         data_index, det_data, flag_index, det_flags, det_scale,
                         det_flag_mask, intervals, shared_flags, shared_flag_mask,
                         use_accel=False):
    n_det = len(pixel_index)
    n_samp = pixels.shape[1]
    nnz = weights.shape[2] if weights.ndim == 3 else 1
    n_pix_submap = zmap.shape[1]
    df = det_flags if (det_flags is not None and det_flags.ndim == 2
                       and det_flags.shape[1] == n_samp) else None
    sf = _opt_flags(shared_flags, n_samp)
    lib().tbo_build_noise_weighted(
        _p(global2local, "i8"), _p(zmap, "f8"), _p(pixel_index, "i4"), _p(pixels, "i8"),
        _p(weight_index, "i4"), _p(weights, "f8"), i64(nnz), _p(data_index, "i4"),
        _p(det_data, "f8"), _p(flag_index, "i4"), _p(df, "u1"), _p(det_scale, "f8"),
        u8(det_flag_mask), _p(intervals), i64(len(intervals)), _p(sf, "u1"),
        u8(shared_flag_mask), i64(n_pix_submap), i64(n_det), i64(n_samp))


_SCAN = {
    np.dtype("f8"): "tbo_scan_map_f64",
    np.dtype("f4"): "tbo_scan_map_f32",
    np.dtype("i8"): "tbo_scan_map_i64",
    np.dtype("i4"): "tbo_scan_map_i32",
}


def scan_map(global2local, n_pix_submap, mapdata, det_data, data_index, pixels, pixel_index,
             weights, weight_index, intervals, data_scale, should_zero, should_subtract,
             should_scale, use_accel=False):
    n_det = len(data_index)
    n_samp = pixels.shape[1]
    nnz = mapdata.shape[2]
    fn = getattr(lib(), _SCAN[mapdata.dtype])
    fn(_p(global2local, "i8"), i64(n_pix_submap), _p(mapdata), i64(nnz), _p(det_data, "f8"),
       _p(data_index, "i4"), _p(pixels, "i8"), _p(pixel_index, "i4"), _p(weights, "f8"),
       _p(weight_index, "i4"), _p(intervals), i64(len(intervals)), f64(data_scale),
       cint(bool(should_zero)), cint(bool(should_subtract)), cint(bool(should_scale)),
       i64(n_det), i64(n_samp))


def template_offset_add_to_signal(step_length, amp_offset, n_amp_views, amplitudes,
                                  amplitude_flags, data_index, det_data, intervals,
                                  use_accel=False):
    n_samp = det_data.shape[1]
    lib().tbo_offset_add_to_signal(
        i64(step_length), i64(amp_offset), _p(n_amp_views, "i8"), _p(amplitudes, "f8"),
        _p(amplitude_flags, "u1"), i32(data_index), _p(det_data, "f8"), _p(intervals),
        i64(len(intervals)), i64(n_samp))


def template_offset_project_signal(data_index, det_data, flag_index, flag_data, flag_mask,
                                   step_length, amp_offset, n_amp_views, amplitudes,
                                   amplitude_flags, intervals, use_accel=False):
    n_samp = det_data.shape[1]
    fd = flag_data if (flag_index >= 0 and flag_data.ndim == 2
                       and flag_data.shape[1] == n_samp) else None
    lib().tbo_offset_project_signal(
        i32(data_index), _p(det_data, "f8"), i32(flag_index), _p(fd, "u1"), u8(flag_mask),
        i64(step_length), i64(amp_offset), _p(n_amp_views, "i8"), _p(amplitudes, "f8"),
        _p(amplitude_flags, "u1"), _p(intervals), i64(len(intervals)), i64(n_samp))


def template_offset_apply_diag_precond(offset_var, amplitudes_in, amplitude_flags,
                                       amplitudes_out, use_accel=False):
    lib().tbo_offset_apply_diag_precond(
        _p(offset_var, "f8"), _p(amplitudes_in, "f8"), _p(amplitude_flags, "u1"),
        _p(amplitudes_out, "f8"), i64(len(amplitudes_in)))


def cov_accum_diag_hits(nsub, subsize, nnz, indx_submap, indx_pix, hits):
    lib().tbo_cov_accum_diag_hits(
        i64(nsub), i64(subsize), i64(len(indx_submap)), _p(indx_submap, "i8"),
        _p(indx_pix, "i8"), _p(hits, "i8"))


def cov_accum_diag_invnpp(nsub, subsize, nnz, indx_submap, indx_pix, weights, scale, invnpp):
    lib().tbo_cov_accum_diag_invnpp(
        i64(nsub), i64(subsize), i64(nnz), i64(len(indx_submap)), _p(indx_submap, "i8"),
        _p(indx_pix, "i8"), _p(weights, "f8"), f64(scale), _p(invnpp, "f8"))


def cov_apply_diag(nsub, subsize, nnz, mat, vec):
    lib().tbo_cov_apply_diag(i64(nsub), i64(subsize), i64(nnz), _p(mat, "f8"), _p(vec, "f8"))


def healpix_ang2pix(nside, nest, theta, phi):
    pix = np.zeros(len(theta), dtype=np.int64)
    lib().tbo_healpix_ang2pix(i64(nside), cint(bool(nest)), i64(len(theta)), _p(theta, "f8"),
                              _p(phi, "f8"), _p(pix, "i8"))
    return pix


def healpix_vec2pix(nside, nest, vec):
    pix = np.zeros(len(vec), dtype=np.int64)
    lib().tbo_healpix_vec2pix(i64(nside), cint(bool(nest)), i64(len(vec)), _p(vec, "f8"),
                              _p(pix, "i8"))
    return pix


def healpix_ring2nest(nside, ringpix):
    out = np.zeros(len(ringpix), dtype=np.int64)
    lib().tbo_healpix_ring2nest(i64(nside), i64(len(ringpix)), _p(ringpix, "i8"), _p(out, "i8"))
    return out


def healpix_nest2ring(nside, nestpix):
    out = np.zeros(len(nestpix), dtype=np.int64)
    lib().tbo_healpix_nest2ring(i64(nside), i64(len(nestpix)), _p(nestpix, "i8"), _p(out, "i8"))
    return out


def num_threads():
    return int(lib().tbo_num_threads())


# ---------------------------------------------------------------------------------------
# numpy restatements of the reference's Python glue
# ---------------------------------------------------------------------------------------

def cov_eigendecompose_diag(nsub, subsize, nnz, data, cond, threshold, invert=True):
    """libtoast/src/toast_map_cov.cpp:246-396 restated with numpy.linalg.eigh.

    ``data`` is [nsub*subsize, nnz(nnz+1)/2] upper triangle row-major, modified in place.
    PINNED: the reference's own implementation runs in the build container with its LAPACK calls
    forwarded to the OpenBLAS scipy bundles (oracle/ref_shim/lapack_shim.cpp); its outputs are
    the fixture tests/golden/cov_invert.npz (tests/test_oracle.py: rcond to 1e-12, the same
    pixels kept, inverses within 1e-13 / rcond).
    """
    npix = nsub * subsize
    block = nnz * (nnz + 1) // 2
    d = data.reshape(npix, block)
    if nnz == 1:
        if invert:
            nz = d[:, 0] != 0
            d[nz, 0] = 1.0 / d[nz, 0]
        if cond is not None:
            cond.reshape(-1)[:] = 1.0
        return
    iu = np.triu_indices(nnz)
    full = np.zeros((npix, nnz, nnz))
    full[:, iu[0], iu[1]] = d
    full[:, iu[1], iu[0]] = d
    evals, evecs = np.linalg.eigh(full)
    emin = evals.min(axis=1)
    emax = evals.max(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        rc = np.where(emax > 0.0, emin / emax, 0.0)
        inv = np.einsum("pik,pk,pjk->pij", evecs, 1.0 / evals, evecs)
    ok = rc >= threshold
    if invert:
        out = np.where(ok[:, None], inv[:, iu[0], iu[1]], 0.0)
        d[:, :] = out
    if cond is not None:
        cond.reshape(-1)[:] = np.where(ok, rc, 0.0)


def pixel_distribution(hit_submaps):
    """pixels.py:59-241: glob2loc[n_submap] int64, -1 where the submap is not local."""
    local = np.flatnonzero(hit_submaps).astype(np.int64)
    g2l = np.full(len(hit_submaps), -1, dtype=np.int64)
    g2l[local] = np.arange(len(local), dtype=np.int64)
    return local, g2l


def offset_layout(n_det, intervals, step_length):
    """templates/offset/offset.py:166-176,245-253: amplitudes are detector-major; per view
    n_amp = ceil(len / step).  Returns (n_amp_views[int64], det_start[int64], n_local)."""
    nav = []
    for iv in intervals:
        ln = int(iv["last"] - iv["first"])
        n = ln // step_length
        if n * step_length < ln:
            n += 1
        nav.append(n)
    nav = np.array(nav, dtype=np.int64)
    per_det = int(nav.sum())
    det_start = np.arange(n_det, dtype=np.int64) * per_det
    return nav, det_start, per_det * n_det


def offset_variance(n_det, n_samp, intervals, step_length, n_amp_views, detnoise, solver_flags,
                    det_flag_mask, good_fraction=0.5):
    """templates/offset/offset.py:283-344: offset_var = 1/(detnoise*n_good), amp_flags where the
    good fraction <= good_fraction or detnoise <= 0.  ``solver_flags`` is [n_det, n_samp] u8 or
    None.  Samples outside every interval never reach an amplitude (the bounds view IS the view)."""
    per_det = int(n_amp_views.sum())
    n_amp = per_det * n_det
    var = np.zeros(n_amp)
    aflags = np.zeros(n_amp, dtype=np.uint8)
    off = 0
    for d in range(n_det):
        for ivw, iv in enumerate(intervals):
            first, last = int(iv["first"]), int(iv["last"])
            na = int(n_amp_views[ivw])
            if detnoise[d] <= 0:
                aflags[off:off + na] = 1
                off += na
                continue
            if solver_flags is not None:
                bad = (solver_flags[d, first:last] & det_flag_mask) != 0
            else:
                bad = np.zeros(last - first, dtype=bool)
            nbad = np.add.reduceat(bad.astype(np.int64), np.arange(0, last - first, step_length))
            amplen = np.full(na, step_length, dtype=np.int64)
            amplen[-1] = (last - first) - (na - 1) * step_length
            ngood = amplen - nbad
            keep = (ngood / amplen) > good_fraction
            with np.errstate(divide="ignore"):
                var[off:off + na] = np.where(keep, 1.0 / (detnoise[d] * ngood), 0.0)
            aflags[off:off + na] = np.where(keep, 0, 1)
            off += na
    return var, aflags


class Problem:
    """Bundle of arrays describing one observation of the destriping problem, in the
    reference's buffer layouts.  Plain attribute bag used by the restated
    solver below and by the tests."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def amp_dot(a, b, aflags):
    """templates/amplitudes.py:523-571: dot over unflagged local amplitudes."""
    return np.float64(np.dot(np.where(aflags == 0, a, 0), np.where(aflags == 0, b, 0)))


def expand_pointing(pb, K):
    """PointingDetectorSimple -> PixelsHealpix -> StokesWeights with kernel namespace K
    (this module or oracle/_ref).  Returns (pixels, weights)."""
    n_det, n_samp = pb.n_det, pb.n_samp
    idx = np.arange(n_det, dtype=np.int32)
    quats = np.zeros((n_det, n_samp, 4))
    K.pointing_detector(pb.focalplane, pb.boresight, idx, quats, pb.intervals,
                        pb.shared_flags, pb.shared_flag_mask, False)
    pixels = np.zeros((n_det, n_samp), dtype=np.int64)
    hits = np.zeros(pb.n_submap, dtype=np.uint8)
    K.pixels_healpix(idx, quats, pb.shared_flags, pb.shared_flag_mask, idx, pixels,
                     pb.intervals, hits, pb.n_pix_submap, pb.nside, pb.nest, False)
    weights = np.zeros((n_det, n_samp, 3))
    K.stokes_weights_IQU(idx, quats, idx, weights, pb.hwp, pb.intervals, pb.epsilon, pb.gamma,
                         pb.cal, pb.IAU, False)
    return pixels, weights, hits


def template_add(pb, K, amps, det_data):
    """TemplateMatrix(transpose=False): mapmaker_templates.py:330-341 -> offset.py:727-810."""
    for d in range(pb.n_det):
        K.template_offset_add_to_signal(pb.step_length, int(pb.det_start[d]), pb.n_amp_views,
                                        amps, pb.amp_flags, d, det_data, pb.intervals, False)


def template_project(pb, K, det_data, amps_out):
    """TemplateMatrix(transpose=True): mapmaker_templates.py:287-298 -> offset.py:813-881."""
    for d in range(pb.n_det):
        if pb.solver_flags is not None:
            K.template_offset_project_signal(d, det_data, d, pb.solver_flags, pb.det_flag_mask,
                                             pb.step_length, int(pb.det_start[d]),
                                             pb.n_amp_views, amps_out, pb.amp_flags,
                                             pb.intervals, False)
        else:
            K.template_offset_project_signal(d, det_data, -1, np.zeros(1, dtype=np.uint8),
                                             pb.det_flag_mask, pb.step_length,
                                             int(pb.det_start[d]), pb.n_amp_views, amps_out,
                                             pb.amp_flags, pb.intervals, False)


def bin_map(pb, K, det_data, covapply, reverse=False):
    """BinMap: mapmaker_binning.py:179-294 = BuildNoiseWeighted + covariance_apply.
    ``reverse``: hand the detectors to the kernel in reverse order -- the same reference
    arithmetic with another (equally valid) summation order; the difference between the two
    results measures how far the reference's own output is defined (tests: order envelope)."""
    idx = np.arange(pb.n_det, dtype=np.int32)
    det_scale = pb.det_scale
    if reverse:
        # (det_scale is indexed by loop position, the data arrays through idx)
        idx = np.ascontiguousarray(idx[::-1])
        det_scale = np.ascontiguousarray(det_scale[::-1])
    zmap = np.zeros((pb.n_local_submap, pb.n_pix_submap, 3))
    if pb.solver_flags is not None:
        K.build_noise_weighted(pb.global2local, zmap, idx, pb.pixels, idx, pb.weights, idx,
                               det_data, idx, pb.solver_flags, det_scale, pb.det_flag_mask,
                               pb.intervals, pb.shared_flags, pb.shared_flag_mask, False)
    else:
        K.build_noise_weighted(pb.global2local, zmap, idx, pb.pixels, idx, pb.weights, idx,
                               det_data, idx, np.zeros((1, 1), dtype=np.uint8), det_scale,
                               pb.det_flag_mask, pb.intervals, pb.shared_flags,
                               pb.shared_flag_mask, False)
    covapply(pb.n_local_submap, pb.n_pix_submap, 3, pb.cov.reshape(-1), zmap.reshape(-1))
    return zmap


def _scan(K):
    return getattr(K, "ops_scan_map_float64", None) or K.scan_map


def solver_lhs(pb, K, amps_in, covapply=None, reverse=False):
    """SolverLHS._exec with full_pointing=True: mapmaker_solve.py:342-506."""
    covapply = covapply or cov_apply_diag
    idx = np.arange(pb.n_det, dtype=np.int32)
    det_temp = np.zeros((pb.n_det, pb.n_samp))
    template_add(pb, K, amps_in, det_temp)
    binned = bin_map(pb, K, det_temp, covapply, reverse=reverse)
    out = np.zeros_like(amps_in)
    det_temp[:] = 0
    template_add(pb, K, amps_in, det_temp)
    _scan(K)(pb.global2local, pb.n_pix_submap, binned, det_temp, idx, pb.pixels, idx,
             pb.weights, idx, pb.intervals, 1.0, False, True, False, False)
    K.noise_weight(det_temp, idx, pb.intervals, pb.det_scale, False)
    template_project(pb, K, det_temp, out)
    return out


def solver_rhs(pb, K, signal, covapply=None, reverse=False):
    """SolverRHS._exec with full_pointing=True: mapmaker_solve.py:107-229."""
    covapply = covapply or cov_apply_diag
    idx = np.arange(pb.n_det, dtype=np.int32)
    binned = bin_map(pb, K, signal, covapply, reverse=reverse)
    det_temp = signal.copy()
    _scan(K)(pb.global2local, pb.n_pix_submap, binned, det_temp, idx, pb.pixels, idx,
             pb.weights, idx, pb.intervals, 1.0, False, True, False, False)
    K.noise_weight(det_temp, idx, pb.intervals, pb.det_scale, False)
    rhs = np.zeros(pb.n_amp)
    template_project(pb, K, det_temp, rhs)
    return rhs


def solve(pb, K, rhs, convergence=1.0e-12, n_iter_max=100, n_iter_min=3, covapply=None,
          prior=None, trace=None):
    """The PCG loop of mapmaker_solve.py:524-755, zero starting guess.  Returns
    (amplitudes, [relative residual per iteration]).  ``prior``: an
    ``oracle.offset_prior.OraclePrior`` -- the LHS then includes the noise prior
    (mapmaker_solve.py:395-412) and the preconditioner is Offset._apply_precond's banded /
    Toeplitz form (offset.py:962-1010).  ``trace``: optional list that receives, per iteration, a
    dict with copies of the state BEFORE the iteration (x, r, d, delta) and what the iteration
    produced (q = A d, alpha, sqsum = r.r after the update) -- the restart-parity tests load the
    state into the device solver, run ONE iteration there and compare."""
    fl = pb.amp_flags
    _lhs, _diag = solver_lhs, K.template_offset_apply_diag_precond
    if prior is not None:
        from . import offset_prior as _OP

        def solver_lhs_p(pb, K, amps, covapply=None):
            out = np.zeros_like(amps)
            _OP.add_prior(prior, amps, fl, out)
            return out + _lhs(pb, K, amps, covapply)

        class _KP:
            @staticmethod
            def template_offset_apply_diag_precond(var, a_in, flags, a_out, use_accel):
                _OP.apply_precond(prior, a_in, flags, a_out)

        return _solve(pb, _KP, rhs, convergence, n_iter_max, n_iter_min, covapply,
                      lambda pb_, K_, a, c=None: solver_lhs_p(pb_, K, a, c), trace)
    return _solve(pb, K, rhs, convergence, n_iter_max, n_iter_min, covapply, solver_lhs, trace)


def _solve(pb, K, rhs, convergence, n_iter_max, n_iter_min, covapply, solver_lhs, trace=None):
    fl = pb.amp_flags
    result = np.zeros_like(rhs)
    lhs_out = solver_lhs(pb, K, result, covapply)
    residual = rhs - lhs_out
    precond = np.zeros_like(rhs)
    K.template_offset_apply_diag_precond(pb.offset_var, residual, fl, precond, False)
    proposal = precond.copy()
    sqsum = amp_dot(rhs, rhs, fl)
    sqsum_init = sqsum
    sqsum_best = sqsum
    last_best = sqsum
    delta = amp_dot(proposal, residual, fl)
    history = []
    for it in range(n_iter_max):
        if not np.isfinite(sqsum):
            raise RuntimeError("Residual is not finite")
        lhs_out = solver_lhs(pb, K, proposal, covapply)
        alpha = delta / amp_dot(proposal, lhs_out, fl)
        if trace is not None:
            trace.append(dict(x=result.copy(), r=residual.copy(), d=proposal.copy(),
                              delta=float(delta), q=lhs_out.copy(), alpha=float(alpha)))
        result += proposal * alpha
        residual -= lhs_out * alpha
        sqsum = amp_dot(residual, residual, fl)
        relative = sqsum / sqsum_init
        history.append(relative)
        if trace is not None:
            trace[-1].update(sqsum=float(sqsum), sqsum_init=float(sqsum_init))
        if relative < convergence or sqsum < 1e-30:
            break
        sqsum_best = min(sqsum, sqsum_best)
        if it % 10 == 0 and it >= n_iter_min:
            if last_best < sqsum_best * 2:
                break
            last_best = sqsum_best
        K.template_offset_apply_diag_precond(pb.offset_var, residual, fl, precond, False)
        delta_last = delta
        delta = amp_dot(precond, residual, fl)
        beta = delta / delta_last
        proposal *= beta
        proposal += precond
    return result, history


# ---------------------------------------------------------------------------------------
# Problem assembly (restates the setup stages of SolveAmplitudes: mapmaker_templates.py:
# 632-990 -- solver flags, pixel distribution, CovarianceAndHits, rcond mask, Offset layout)
# ---------------------------------------------------------------------------------------

def global_to_local(pix, n_pix_submap, g2l):
    """_libtoast/pixels.cpp:10-41."""
    pix = np.asarray(pix, dtype=np.int64)
    good = pix >= 0
    gsm = np.where(good, pix // n_pix_submap, 0)
    sm = np.where(good, g2l[gsm], -1).astype(np.int64)
    lp = np.where(good, pix - gsm * n_pix_submap, -1).astype(np.int64)
    return sm, lp


def build_problem(obs, K=None, shared_flag_mask=1, det_flag_mask=1, rcond_threshold=1.0e-3,
                  IAU=False, hwp=None, use_flags=True, external=None):
    """Expand pointing and assemble every buffer one PCG iteration touches.

    ``obs`` is a dict from ``toast_b200.synthetic.make_observation``; ``K`` is the kernel
    namespace (this module, or the compiled reference from ``load_ref()``).

    ``external``: dict(hit_submaps, cov [n_local, n_pix_submap, 6], rcond [n_local *
    n_pix_submap]) computed from a LARGER detector set that contains this observation's
    detectors.  The pixel distribution, the pixel covariance and the rcond mask are then taken
    from it instead of being accumulated from these detectors alone, so that a few-detector
    subset of a production-size problem keeps the production covariance (the full-size parity
    tests: the compiled reference runs 4 detectors, the covariance comes from the whole shard)."""
    K = K or sys.modules[__name__]
    from toast_b200.synthetic import n_submap_for

    pb = Problem()
    pb.n_det, pb.n_samp = obs["n_det"], obs["n_samp"]
    pb.nside, pb.nest = obs["nside"], obs["nest"]
    pb.n_submap, pb.n_pix_submap = n_submap_for(pb.nside, obs["nside_submap"])
    pb.focalplane, pb.boresight = obs["focalplane"], obs["boresight"]
    pb.intervals = obs["intervals"]
    pb.epsilon, pb.gamma, pb.cal = obs["epsilon"], obs["gamma"], obs["cal"]
    pb.IAU = IAU
    pb.hwp = hwp if hwp is not None else np.zeros(1)
    pb.shared_flags = obs["shared_flags"] if use_flags else np.zeros(1, dtype=np.uint8)
    pb.shared_flag_mask = shared_flag_mask
    pb.det_flag_mask = det_flag_mask
    pb.det_scale = np.ascontiguousarray(obs["detweight"])
    pb.step_length = obs["step_length"]

    pb.pixels, pb.weights, pb.hit_submaps = expand_pointing(pb, K)
    if external is not None:
        ext_hits = np.asarray(external["hit_submaps"], dtype=np.uint8)
        assert np.all(ext_hits[pb.hit_submaps != 0] != 0), "external hit map must cover this one"
        pb.own_hit_submaps = pb.hit_submaps
        pb.hit_submaps = ext_hits
    pb.local_submaps, pb.global2local = pixel_distribution(pb.hit_submaps)
    pb.n_local_submap = len(pb.local_submaps)

    in_view = np.zeros(pb.n_samp, dtype=bool)
    for iv in pb.intervals:
        in_view[iv["first"]:iv["last"]] = True

    # solver flags: bit0 = input flags / bad pointing / outside view (SURVEY 8b vi)
    sf = np.zeros((pb.n_det, pb.n_samp), dtype=np.uint8)
    if use_flags:
        sf |= ((obs["det_flags"] & det_flag_mask) != 0).astype(np.uint8)
        sf |= ((obs["shared_flags"] & shared_flag_mask) != 0).astype(np.uint8)[None, :]
    sf |= (~in_view).astype(np.uint8)[None, :]
    sf |= (pb.pixels < 0).astype(np.uint8)

    # inverse pixel covariance (mapmaker_utils.py:440-515) -> covariance
    block = 6
    invcov = np.zeros(pb.n_local_submap * pb.n_pix_submap * block if external is None else 0)
    for d in range(pb.n_det if external is None else 0):
        for iv in pb.intervals:
            a, b = int(iv["first"]), int(iv["last"])
            sm, lp = global_to_local(pb.pixels[d, a:b], pb.n_pix_submap, pb.global2local)
            lp[sf[d, a:b] != 0] = -1
            cov_accum_diag_invnpp(pb.n_local_submap, pb.n_pix_submap, 3, sm, lp,
                                  np.ascontiguousarray(pb.weights[d, a:b]).reshape(-1),
                                  float(pb.det_scale[d]), invcov)
    if external is None:
        pb.invcov = invcov.copy()
        rcond = np.zeros(pb.n_local_submap * pb.n_pix_submap)
        cov_eigendecompose_diag(pb.n_local_submap, pb.n_pix_submap, 3, invcov, rcond,
                                rcond_threshold, True)
    else:
        pb.invcov = None
        invcov = np.ascontiguousarray(external["cov"], dtype=np.float64).reshape(-1)
        rcond = np.ascontiguousarray(external["rcond"], dtype=np.float64).reshape(-1)
        assert invcov.size == pb.n_local_submap * pb.n_pix_submap * block
        assert rcond.size == pb.n_local_submap * pb.n_pix_submap
    pb.cov = invcov.reshape(pb.n_local_submap, pb.n_pix_submap, block)
    pb.rcond = rcond

    # rcond mask -> solver flags bit0 as well (mapmaker_templates.py:895-939 uses a
    # separate bit; one bit suffices since the solver mask is 255)
    badpix = (rcond.reshape(pb.n_local_submap, pb.n_pix_submap) == 0.0)
    for d in range(pb.n_det):
        sm, lp = global_to_local(pb.pixels[d], pb.n_pix_submap, pb.global2local)
        ok = sm >= 0
        bad = np.zeros(pb.n_samp, dtype=bool)
        bad[ok] = badpix[sm[ok], lp[ok]]
        sf[d, bad] |= 1
    pb.solver_flags = sf
    # the solver bins with solver flags only (shared mask 0: mapmaker_templates.py:959-962)
    pb.shared_flag_mask_solver = 0

    pb.n_amp_views, pb.det_start, pb.n_amp = offset_layout(pb.n_det, pb.intervals,
                                                           pb.step_length)
    pb.offset_var, pb.amp_flags = offset_variance(
        pb.n_det, pb.n_samp, pb.intervals, pb.step_length, pb.n_amp_views, pb.det_scale,
        pb.solver_flags, pb.det_flag_mask)
    return pb
