"""TEST INFRASTRUCTURE ONLY -- never imported by toast_b200.

CPU stand-ins for the device side of ``ops.MapMaker`` so that its HOST LOGIC (solver flags, pixel
distribution, rcond mask, amplitude flags / variance, the order of the stages, the products and
where they land) runs in the CPU suite: the per-sample compute is done by the oracle on the numpy
views of CPU tensors, exactly as tests/test_distributed_gloo.py stands the oracle in for the
kernels.  On the GPU the same operator code runs against the real kernels (tests/test_gpu_ops.py).

``install(monkeypatch)`` swaps

* the five kernel wrappers MapMaker calls (``toast_b200.kernels``: cov_accum, cov_invert,
  ops_scan_map_float64, template_offset_project_signal_batch, template_offset_add_to_signal_batch),
* ``toast_b200.solver.DeviceObservation`` / ``Destriper`` (pointing expansion, RHS, PCG, binning
  by the oracle's restatement of the same stages), and
* the handful of ``torch.cuda`` stream calls the operator makes (no-ops on the CPU).
"""

import contextlib

import numpy as np
import torch

from oracle import toast_oracle as O


def _np(t):
    return t.numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


class _KC:
    """The kernel wrappers of toast_b200.kernels that MapMaker uses, on CPU tensors."""

    @staticmethod
    def cov_accum(global2local, n_local_submap, n_pix_submap, nnz, hits, invcov, pixel_index,
                  pixels, weight_index, weights, flag_index, det_flags, det_scale, det_flag_mask,
                  intervals, shared_flags, shared_flag_mask, use_accel=False, stream=None):
        g2l = np.asarray(global2local, dtype=np.int64)
        px = _np(pixels)
        wt = None if weights is None else _np(weights)
        fl = None if det_flags is None else _np(det_flags)
        for k, pi in enumerate(pixel_index):
            wi = None if weight_index is None else weight_index[k]
            for iv in intervals:
                a, b = int(iv["first"]), int(iv["last"])
                sm, lp = O.global_to_local(px[pi, a:b], n_pix_submap, g2l)
                bad = np.zeros(b - a, dtype=bool)
                if fl is not None:
                    bad |= (fl[flag_index[k], a:b] & det_flag_mask) != 0
                if shared_flags is not None:
                    bad |= (_np(shared_flags)[a:b] & shared_flag_mask) != 0
                lp[bad] = -1
                if hits is not None:
                    O.cov_accum_diag_hits(n_local_submap, n_pix_submap, nnz, sm, lp, _np(hits))
                if invcov is not None:
                    O.cov_accum_diag_invnpp(n_local_submap, n_pix_submap, nnz, sm, lp,
                                            np.ascontiguousarray(wt[wi, a:b]).reshape(-1),
                                            float(det_scale[k]), _np(invcov).reshape(-1))

    @staticmethod
    def cov_invert(npix, nnz, cov, rcond, threshold, use_accel=False, stream=None):
        O.cov_eigendecompose_diag(1, npix, nnz, _np(cov).reshape(-1), _np(rcond), threshold, True)

    @staticmethod
    def ops_scan_map_float64(global2local, n_pix_submap, mapdata, det_data, data_index, pixels,
                             pixel_index, weights, weight_index, intervals, data_scale,
                             should_zero, should_subtract, should_scale, use_accel=False,
                             stream=None):
        O.scan_map(np.asarray(global2local, dtype=np.int64), n_pix_submap, _np(mapdata),
                   _np(det_data), data_index, _np(pixels), pixel_index, _np(weights), weight_index,
                   intervals, data_scale, should_zero, should_subtract, should_scale, False)

    @staticmethod
    def template_offset_project_signal_batch(data_index, det_data, flag_index, flag_data,
                                             flag_mask, step_length, amp_offsets, n_amp_views,
                                             amplitudes, amplitude_flags, intervals,
                                             use_accel=False, stream=None):
        for k, di in enumerate(data_index):
            O.template_offset_project_signal(
                int(di), _np(det_data), int(flag_index[k]), _np(flag_data), flag_mask,
                int(step_length), int(amp_offsets[k]), np.asarray(n_amp_views, dtype=np.int64),
                _np(amplitudes), _np(amplitude_flags), intervals, False)

    @staticmethod
    def template_offset_add_to_signal_batch(step_length, amp_offsets, n_amp_views, amplitudes,
                                            amplitude_flags, data_index, det_data, intervals,
                                            use_accel=False, stream=None):
        for k, di in enumerate(data_index):
            O.template_offset_add_to_signal(
                int(step_length), int(amp_offsets[k]), np.asarray(n_amp_views, dtype=np.int64),
                _np(amplitudes), _np(amplitude_flags), int(di), _np(det_data), intervals, False)


class FakeDeviceObservation:
    """The attributes ops.MapMaker reads from solver.DeviceObservation, as CPU tensors."""

    def __init__(self, *, focalplane, boresight, intervals, det_scale, step_length, nside, nest,
                 n_pix_submap, n_submap, global2local, epsilon, gamma, cal, IAU=False,
                 shared_flags=None, shared_flag_mask=0, solver_flags=None, solver_flag_mask=255,
                 hwp=None, amp_offsets=None, device="cpu", **unused):
        self.n_det, self.n_samp = int(focalplane.shape[0]), int(boresight.shape[0])
        self.focalplane, self.boresight = np.ascontiguousarray(focalplane), _np(boresight)
        self.intervals, self.step_length = np.ascontiguousarray(intervals), int(step_length)
        self.det_scale = np.ascontiguousarray(det_scale, dtype=np.float64)
        self.nside, self.nest, self.IAU = int(nside), bool(nest), bool(IAU)
        self.n_pix_submap, self.n_submap = int(n_pix_submap), int(n_submap)
        self.global2local = np.ascontiguousarray(global2local, dtype=np.int64)
        self.epsilon, self.gamma, self.cal = (np.ascontiguousarray(x) for x in (epsilon, gamma, cal))
        self.shared_flags = None if shared_flags is None else _np(shared_flags)
        self.shared_flag_mask, self.solver_flag_mask = int(shared_flag_mask), int(solver_flag_mask)
        self.solver_flags = solver_flags           # torch uint8 [n_det, n_samp], edited in place
        self.hwp = hwp
        nav = []
        for iv in self.intervals:
            ln = int(iv["last"] - iv["first"])
            nav.append(ln // self.step_length + (1 if ln % self.step_length else 0))
        self.n_amp_views = np.array(nav, dtype=np.int64)
        self.n_amp_det = int(self.n_amp_views.sum())
        self.amp_offsets = np.ascontiguousarray(amp_offsets, dtype=np.int64)
        self.n_amp = self.n_amp_det * self.n_det
        self.pixels = self.weights = None

    def expand_pointing(self, hit_submaps=None):
        pb = O.Problem(n_det=self.n_det, n_samp=self.n_samp, nside=self.nside, nest=self.nest,
                       n_submap=self.n_submap, n_pix_submap=self.n_pix_submap,
                       focalplane=self.focalplane, boresight=self.boresight,
                       intervals=self.intervals, epsilon=self.epsilon, gamma=self.gamma,
                       cal=self.cal, IAU=self.IAU,
                       hwp=np.zeros(1) if self.hwp is None else _np(self.hwp),
                       shared_flags=np.zeros(1, dtype=np.uint8) if self.shared_flags is None
                       else self.shared_flags, shared_flag_mask=self.shared_flag_mask)
        pixels, weights, hits = O.expand_pointing(pb, O)
        self.pixels, self.weights = torch.from_numpy(pixels), torch.from_numpy(weights)
        if hit_submaps is not None:
            hit_submaps |= hits

    def set_global2local(self, g2l):
        self.global2local = np.ascontiguousarray(g2l, dtype=np.int64)


def _world():
    import torch.distributed as dist

    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _covapply(nsub, subsize, nnz, mat, vec):
    """The reduction of the path and the step after it (pixels.py:710-779 + covariance.py:
    262-306): sum the noise-weighted map over the ranks (gloo, in place), then apply C."""
    if _world() > 1:
        import torch.distributed as dist

        dist.all_reduce(torch.from_numpy(vec))
    O.cov_apply_diag(nsub, subsize, nnz, mat, vec)


@contextlib.contextmanager
def _global_dots():
    """Amplitude dot products summed over the ranks (templates/amplitudes.py:560-571) inside the
    oracle's PCG."""
    if _world() == 1:
        yield
        return
    import torch.distributed as dist

    local = O.amp_dot

    def dot(a, b, aflags):
        v = torch.tensor([float(local(a, b, aflags))], dtype=torch.float64)
        dist.all_reduce(v)
        return np.float64(v[0].item())

    O.amp_dot = dot
    try:
        yield
    finally:
        O.amp_dot = local


class FakeDestriper:
    """solver.Destriper's interface towards ops.MapMaker, computed with the oracle's kernels in
    the order of SolverRHS / SolverLHS / solve() / BinMap (mapmaker_solve.py:107-229, 342-506,
    524-755; mapmaker_binning.py:179-294) over ONE OR SEVERAL observations: every observation
    bins into the same map, the map is reduced and multiplied by the covariance once, every
    observation then scans it.  Under torch.distributed (gloo) the map and the dot products are
    summed over the ranks where the device solver does it."""

    def __init__(self, observations, n_local_submap, n_pix_submap, cov, offset_var, amp_flags,
                 regen=False, device="cpu", prior=None, **unused):
        self.prior = prior        # an oracle.offset_prior.OraclePrior (install(oracle_prior=True))
        assert prior is None or len(observations) == 1
        self.obs = list(observations)
        self.n_local_submap, self.n_pix_submap = int(n_local_submap), int(n_pix_submap)
        self.cov = cov if isinstance(cov, torch.Tensor) else torch.from_numpy(np.asarray(cov))
        self.offset_var = torch.as_tensor(_np(offset_var)).to(torch.float64)
        self.amp_flags = torch.as_tensor(_np(amp_flags)).to(torch.uint8)
        self.n_amp = int(self.offset_var.numel())

    def _pb(self, d):
        return O.Problem(
            n_det=d.n_det, n_samp=d.n_samp, step_length=d.step_length, det_start=d.amp_offsets,
            n_amp_views=d.n_amp_views, n_amp=self.n_amp, amp_flags=_np(self.amp_flags),
            offset_var=_np(self.offset_var), intervals=d.intervals,
            solver_flags=_np(d.solver_flags), det_flag_mask=d.solver_flag_mask,
            det_scale=d.det_scale, global2local=d.global2local, pixels=_np(d.pixels),
            weights=_np(d.weights), shared_flags=np.zeros(1, dtype=np.uint8), shared_flag_mask=0,
            n_local_submap=self.n_local_submap, n_pix_submap=self.n_pix_submap,
            cov=_np(self.cov).reshape(self.n_local_submap, self.n_pix_submap, 6))

    def _bin(self, pbs, timestreams):
        z = np.zeros((self.n_local_submap, self.n_pix_submap, 3))
        for pb, tod in zip(pbs, timestreams):
            idx = np.arange(pb.n_det, dtype=np.int32)
            O.build_noise_weighted(pb.global2local, z, idx, pb.pixels, idx, pb.weights, idx, tod,
                                   idx, pb.solver_flags, pb.det_scale, pb.det_flag_mask,
                                   pb.intervals, pb.shared_flags, pb.shared_flag_mask, False)
        _covapply(self.n_local_submap, self.n_pix_submap, 3, pbs[0].cov.reshape(-1),
                  z.reshape(-1))
        return z

    @staticmethod
    def _scan_weight_project(pb, binned, tod, out):
        idx = np.arange(pb.n_det, dtype=np.int32)
        O.scan_map(pb.global2local, pb.n_pix_submap, binned, tod, idx, pb.pixels, idx,
                   pb.weights, idx, pb.intervals, 1.0, False, True, False, False)
        O.noise_weight(tod, idx, pb.intervals, pb.det_scale, False)
        O.template_project(pb, O, tod, out)

    def _lhs(self, pbs, amps):
        def from_template():
            tods = [np.zeros((pb.n_det, pb.n_samp)) for pb in pbs]
            for pb, tod in zip(pbs, tods):
                O.template_add(pb, O, amps, tod)
            return tods

        binned = self._bin(pbs, from_template())
        out = np.zeros_like(amps)
        for pb, tod in zip(pbs, from_template()):
            self._scan_weight_project(pb, binned, tod, out)
        return out

    def rhs(self, signals):
        pbs = [self._pb(d) for d in self.obs]
        binned = self._bin(pbs, [_np(s) for s in signals])
        out = np.zeros(self.n_amp)
        for pb, sig in zip(pbs, signals):
            self._scan_weight_project(pb, binned, _np(sig).copy(), out)
        return torch.from_numpy(out)

    def solve(self, rhs, convergence=1.0e-12, n_iter_max=100, n_iter_min=3, x0=None):
        pbs = [self._pb(d) for d in self.obs]
        if self.prior is not None:
            # (one observation: _lhs is then O.solver_lhs, which O.solve wraps with the prior)
            amps, hist = O.solve(pbs[0], O, _np(rhs), convergence=convergence,
                                 n_iter_max=n_iter_max, n_iter_min=n_iter_min,
                                 covapply=_covapply, prior=self.prior)
            return torch.from_numpy(amps), hist
        with _global_dots():
            amps, hist = O._solve(pbs[0], O, _np(rhs), convergence, n_iter_max, n_iter_min,
                                  _covapply, lambda pb, K, a, c=None: self._lhs(pbs, a))
        return torch.from_numpy(amps), hist

    def bin_signal(self, signals):
        pbs = [self._pb(d) for d in self.obs]
        return torch.from_numpy(self._bin(pbs, [_np(s) for s in signals]))


class _Stream:
    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass


def _oracle_build_prior(self, data):
    """templates.Offset._build_prior with the oracle's restatement of offset.py:356-560 (one
    observation): the noise prior the stand-in solver applies."""
    from oracle import offset_prior as OP

    ob = data.obs[0]
    dets = self._all_dets
    noise = ob[self.noise_model]
    t = ob.shared[self.times]
    self._prior = OP.build_prior(
        noise.freq(dets[0]), np.array([noise.psd(d) for d in dets]),
        np.array([noise.detector_weight(d) for d in dets]), self._offsetvar, self._obs_views[0],
        float(t[-1] - t[0]), self.step_time, self._obs_rate[0], precond_width=self.precond_width)


def install(monkeypatch, oracle_prior=False):
    """Route ops.MapMaker(device="cpu") through the stand-ins above."""
    import toast_b200.ops.mapmaker as MM
    import toast_b200.solver as SV

    if oracle_prior:
        import toast_b200.templates.offset as TO

        monkeypatch.setattr(TO.Offset, "_build_prior", _oracle_build_prior)

    for name in ("cov_accum", "cov_invert", "ops_scan_map_float64",
                 "template_offset_project_signal_batch", "template_offset_add_to_signal_batch"):
        monkeypatch.setattr(MM.KC, name, getattr(_KC, name))
    monkeypatch.setattr(SV, "DeviceObservation", FakeDeviceObservation)
    monkeypatch.setattr(SV, "Destriper", FakeDestriper)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None, priority=0: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)


def install_operator_kernels(monkeypatch):
    """Route the per-operator kernels (the ``_libtoast`` names the operator mirror calls with
    ``use_accel=False``: host buffers in, host buffers out) to the oracle, which has the
    reference's positional signatures, so that the operators' own logic -- views, detector rows,
    flags, ``ensure()`` / skip, the pixel distribution, Pipeline staging -- runs in the CPU suite."""
    import toast_b200._libtoast as KP
    import toast_b200.kernels as KCm

    for name in ("pointing_detector", "pixels_healpix", "stokes_weights_IQU", "stokes_weights_I",
                 "noise_weight", "build_noise_weighted", "template_offset_add_to_signal",
                 "template_offset_project_signal", "template_offset_apply_diag_precond",
                 "cov_apply_diag"):
        monkeypatch.setattr(KP, name, getattr(O, name))
    for dt in ("float64", "float32", "int64", "int32"):
        monkeypatch.setattr(KP, f"ops_scan_map_{dt}", O.scan_map)
    monkeypatch.setattr(KP, "accel_present", lambda arr, name: False)
    for name in ("template_offset_project_signal_batch", "cov_accum", "cov_invert"):
        monkeypatch.setattr(KCm, name, getattr(_KC, name))
