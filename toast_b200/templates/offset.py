"""templates.Offset (``templates/offset/offset.py:28-1030``): step-wise constant baselines per
detector and view, optionally with the noise prior (``use_noise_prior``; ``offset_prior.py``).

Amplitude layout (``offset.py:166-176, 245-253``): detector-major; per observation and view
``ceil(view_len / step_length)`` amplitudes with ``step_length = rint(step_time * rate)``
(``:723-724``).  An amplitude is flagged when its good-sample fraction is <= ``good_fraction``
or the detector noise weight is <= 0, and ``offset_var = 1 / (detweight * n_good)``
(``:283-344``) is the diagonal preconditioner.  Per-sample work is done by the CUDA kernels.
"""

import numpy as np

from .. import _libtoast as K
from .. import kernels as KC
from .amplitudes import Amplitudes
from .offset_prior import OffsetPriorBuilder, prior_frequencies


def median_spacing(t, device=None):
    """``np.median(np.diff(times))`` -- the sample spacing the reference derives the rate from
    (``offset.py:167`` -> ``utils.py:655-685`` ``rate_from_times``).  With ``device`` given and a long vector the difference and the sort
    run there with torch (12 hours at 50 Hz cost a tenth of a second in numpy's partition); the
    value is the same: IEEE differences, the middle element of the sorted differences or the
    mean of the two middle ones."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    n = len(t) - 1
    if device is None or n < (1 << 18):
        return float(np.median(np.diff(t)))
    import torch

    tt = torch.from_numpy(t).to(device)
    s = torch.sort(tt[1:] - tt[:-1]).values
    if n % 2 == 1:
        return float(s[n // 2])
    return (float(s[n // 2 - 1]) + float(s[n // 2])) / 2.0


class Template:
    """templates/template.py:24-263 (the parts Offset uses)."""

    _defaults = dict(name=None, data=None, view=None, det_data="signal", det_mask=1,
                     det_flags=None, det_flag_mask=1)

    def __init__(self, **kwargs):
        for klass in reversed(type(self).__mro__):
            for k, v in getattr(klass, "_defaults", {}).items():
                setattr(self, k, v)
        data = kwargs.pop("data", None)
        for k, v in kwargs.items():
            if not hasattr(self, k):
                raise AttributeError(f"{type(self).__name__} has no trait '{k}'")
            setattr(self, k, v)
        if self.name is None:
            self.name = type(self).__name__
        if data is not None:
            self.initialize(data)

    def initialize(self, data, detectors=None):
        self.data = data
        self._initialize(data, detectors)

    def detectors(self):
        return self._all_dets

    def zeros(self):
        return self._zeros()

    def add_to_signal(self, detector, amplitudes, use_accel=False):
        self._add_to_signal(detector, amplitudes, use_accel=use_accel)

    def project_signal(self, detector, amplitudes, use_accel=False):
        self._project_signal(detector, amplitudes, use_accel=use_accel)

    def add_prior(self, amplitudes_in, amplitudes_out, use_accel=False):
        self._add_prior(amplitudes_in, amplitudes_out, use_accel=use_accel)

    def apply_precond(self, amplitudes_in, amplitudes_out, use_accel=False):
        self._apply_precond(amplitudes_in, amplitudes_out, use_accel=use_accel)


class Offset(Template):
    _defaults = dict(step_time=10000.0, times="times", noise_model=None, good_fraction=0.5,
                     use_noise_prior=False, precond_width=20)

    def _step_length(self, stime, rate):
        return int(np.rint(stime * rate))

    def _initialize(self, new_data, detectors=None):
        if self.use_noise_prior and self.noise_model is None:
            raise RuntimeError("cannot use the noise prior without a noise model")  # :136-138
        self._prior = None
        # offset.py:136-141: with a noise prior the baselines span the whole observation and the
        # view only flags samples; without it the view defines the baseline boundaries
        self._bounds_view = None if self.use_noise_prior else self.view
        self._obs_flags = {}
        self._obs_views, self._obs_view_flags = {}, {}
        self._obs_rate, self._obs_dets = {}, {}
        all_dets = {}
        for iob, ob in enumerate(new_data.obs):
            t = ob.shared[self.times]
            rate = 1.0 / median_spacing(t, getattr(self, "_device", None)) if len(t) > 1 else 1.0
            self._obs_rate[iob] = rate
            step = self._step_length(self.step_time, rate)
            views = []
            for vw in ob.intervals[self._bounds_view]:
                ln = int(vw["last"] - vw["first"])
                n = ln // step
                if n * step < ln:
                    n += 1
                views.append(n)
            self._obs_views[iob] = np.array(views, dtype=np.int64)
            vf = np.ones(ob.n_local_samples, dtype=np.uint8)
            for vw in ob.intervals[self.view]:
                vf[vw["first"]:vw["last"]] = 0
            self._obs_view_flags[iob] = vf
            self._obs_dets[iob] = set()
            for d in ob.select_local_detectors(selection=detectors, flagmask=self.det_mask):
                if d not in ob.detdata[self.det_data].detectors:
                    continue
                self._obs_dets[iob].add(d)
                all_dets.setdefault(d, None)
        self._all_dets = list(all_dets)
        self._det_start = {}
        offset = 0
        for det in self._all_dets:
            self._det_start[det] = offset
            for iob, ob in enumerate(new_data.obs):
                if det in self._obs_dets[iob]:
                    offset += int(np.sum(self._obs_views[iob]))
        self._n_local = offset
        self._n_global = offset
        if new_data.comm.comm_world is not None:
            buf = np.array([float(offset)])
            new_data.comm.allreduce_(buf)
            self._n_global = int(buf[0])

        # amplitude flags and variance: n_good per step = F^T (good-sample indicator), computed
        # with the projection kernel instead of the reference's Python loop (offset.py:283-344)
        self._amp_flags = np.zeros(self._n_local, dtype=bool)
        self._offsetvar = np.zeros(self._n_local)
        if self._n_local == 0:
            return
        if getattr(self, "_defer_variance", False):
            # MapMaker computes flags and variance itself, on the device, from the full solver
            # flags (and then builds the prior); only the per-observation flag arrays that
            # project_signal needs under a noise prior are prepared here
            if self.use_noise_prior:
                for iob, ob in enumerate(new_data.obs):
                    self._obs_flags[iob] = self._combined_flags(ob, iob)
            return
        n_good = np.zeros(self._n_local)
        amplen = np.zeros(self._n_local)
        detnoise = np.ones(self._n_local)
        zero_flags = np.zeros(self._n_local, dtype=np.uint8)
        for iob, ob in enumerate(new_data.obs):
            dets = [d for d in self._all_dets if d in self._obs_dets[iob]]
            if len(dets) == 0:
                continue
            step = self._step_length(self.step_time, self._obs_rate[iob])
            nav = self._obs_views[iob]
            per_det = int(nav.sum())
            offs = np.array([self._obs_amp_offset(d, iob) for d in dets], dtype=np.int64)
            ones = np.ones((len(dets), ob.n_local_samples))
            didx = np.arange(len(dets), dtype=np.int32)
            bounds = ob.intervals[self._bounds_view]
            if self.use_noise_prior:
                # the view is not what the kernel iterates over: its complement is a flag
                # (offset.py:318-325), kept per observation for project_signal as well
                self._obs_flags[iob] = self._combined_flags(ob, iob)
                fd = self._obs_flags[iob]
                fidx = np.array([self._flag_row(ob, d) for d in dets], dtype=np.int32)
                KC.template_offset_project_signal_batch(didx, ones, fidx, fd, 1, step, offs, nav,
                                                        n_good, zero_flags, bounds)
            elif self.det_flags is not None:
                fl = np.ascontiguousarray(
                    ob.detdata[self.det_flags].data[ob.detdata[self.det_flags].indices(dets)])
                KC.template_offset_project_signal_batch(didx, ones, didx, fl, self.det_flag_mask,
                                                        step, offs, nav, n_good, zero_flags,
                                                        bounds)
            else:
                KC.template_offset_project_signal_batch(didx, ones, None, None, 0, step, offs,
                                                        nav, n_good, zero_flags, bounds)
            lens = np.concatenate([
                np.minimum(step, int(vw["last"] - vw["first"]) - step * np.arange(na))
                for vw, na in zip(bounds, nav)]) if per_det else np.zeros(0)
            for d, o in zip(dets, offs):
                amplen[o:o + per_det] = lens
                if self.noise_model is not None:
                    detnoise[o:o + per_det] = ob[self.noise_model].detector_weight(d)
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = np.where(amplen > 0, n_good / amplen, 0.0)
            keep = (frac > self.good_fraction) & (detnoise > 0)
            self._offsetvar[:] = np.where(keep, 1.0 / (detnoise * n_good), 0.0)
        self._amp_flags[:] = ~keep
        # (MapMaker recomputes the variance from the full solver flags first and builds the
        # prior itself: _defer_prior)
        if self.use_noise_prior and not getattr(self, "_defer_prior", False):
            self._build_prior(new_data)

    def _flag_row(self, ob, det):
        """Row of ``det`` in the combined flag array of its observation."""
        if self.det_flags is not None:
            return int(ob.detdata[self.det_flags].indices([det])[0])
        return 0

    def _combined_flags(self, ob, iob):
        """(det_flags & mask != 0) | outside-the-view as one uint8 array in the layout of the
        flag detdata (a single shared row without detector flags): what the reference builds
        per call at offset.py:829-832."""
        vf = self._obs_view_flags[iob]
        if self.det_flags is None:
            return np.ascontiguousarray(vf[None, :])
        fd = ob.detdata[self.det_flags].data
        return np.ascontiguousarray(((fd & self.det_flag_mask) != 0).astype(np.uint8)
                                    | vf[None, :])

    def _build_prior(self, data):
        """offset.py:203-222 + 356-560: per (detector, observation, view) the real-space noise
        filter and the preconditioner, uploaded once; applied by tb_offset_prior_add / _precond."""
        b = self._prior_builder(data)
        if b is not None:
            self._prior = b.finish()

    def _prior_builder(self, data):
        """Host-side assembly only (no device needed)."""
        b = OffsetPriorBuilder(self._n_local, self.precond_width)
        any_prior = False
        for det in self._all_dets:
            for iob, ob in enumerate(data.obs):
                if det not in self._obs_dets[iob]:
                    continue
                t = ob.shared[self.times]
                freq = prior_frequencies(float(t[-1] - t[0]), self.step_time, self._obs_rate[iob])
                start = self._obs_amp_offset(det, iob)
                if freq is None:
                    # a single baseline in this observation: the reference disables the prior
                    # for it (:208-214) and its amplitudes come out as zeros (:946-947)
                    b.add_cut_detector(start, self._obs_views[iob])
                    continue
                noise = ob[self.noise_model]
                b.add_detector(start, self._obs_views[iob], noise.freq(det), noise.psd(det),
                               noise.detector_weight(det), self._offsetvar, freq, self.step_time)
                any_prior = True
        return b if (any_prior or self._n_local > 0) else None

    def _obs_amp_offset(self, det, iob):
        off = self._det_start[det]
        for j in range(iob):
            if det in self._obs_dets[j]:
                off += int(np.sum(self._obs_views[j]))
        return off

    def _zeros(self):
        z = Amplitudes(self.data.comm, self._n_global, self._n_local)
        z.local_flags[:] = np.where(self._amp_flags, 1, 0)
        return z

    def _add_to_signal(self, detector, amplitudes, use_accel=False):
        if detector not in self._all_dets:
            return
        for iob, ob in enumerate(self.data.obs):
            if detector not in self._obs_dets[iob]:
                continue
            step = self._step_length(self.step_time, self._obs_rate[iob])
            K.template_offset_add_to_signal(
                step, self._obs_amp_offset(detector, iob), self._obs_views[iob], amplitudes.local,
                amplitudes.local_flags, int(ob.detdata[self.det_data].indices([detector])[0]),
                ob.detdata[self.det_data].data, ob.intervals[self._bounds_view], use_accel)

    def _project_signal(self, detector, amplitudes, use_accel=False):
        if detector not in self._all_dets:
            return
        for iob, ob in enumerate(self.data.obs):
            if detector not in self._obs_dets[iob]:
                continue
            step = self._step_length(self.step_time, self._obs_rate[iob])
            fmask = self.det_flag_mask
            if self.use_noise_prior:
                # baselines span the observation: samples outside the view are cut by the
                # combined flags built at initialisation (registered on first accel use)
                fidx, fdata, fmask = self._flag_row(ob, detector), self._obs_flags[iob], 1
                if use_accel and not K.accel_present(fdata, "offset_flags"):
                    K.accel_create(fdata, "offset_flags")
                    K.accel_update_device(fdata, "offset_flags")
            elif self.det_flags is not None:
                # the reference ORs the view flags into a per-call host copy of the whole flag
                # array (offset.py:829-832); the view is already what the kernel iterates over,
                # so the registered flag buffer can be used directly (SURVEY 8b vi)
                fidx = int(ob.detdata[self.det_flags].indices([detector])[0])
                fdata = ob.detdata[self.det_flags].data
            else:
                fidx = -1
                fdata = np.zeros((1, 1), dtype=np.uint8)
            K.template_offset_project_signal(
                int(ob.detdata[self.det_data].indices([detector])[0]),
                ob.detdata[self.det_data].data, fidx, fdata, fmask, step,
                self._obs_amp_offset(detector, iob), self._obs_views[iob], amplitudes.local,
                amplitudes.local_flags, ob.intervals[self._bounds_view], use_accel)

    def prior(self):
        """The device-resident noise prior (None without ``use_noise_prior``), for Destriper."""
        return self._prior

    def _add_prior(self, amplitudes_in, amplitudes_out, use_accel=False):
        if not self.use_noise_prior or self._prior is None:
            return  # offset.py:884-887: nothing without a noise prior
        # (the reference raises NotImplementedError for use_accel here, :888-891)
        self._prior.add(amplitudes_in.local, amplitudes_in.local_flags, amplitudes_out.local,
                        use_accel=use_accel)

    def _apply_precond(self, amplitudes_in, amplitudes_out, use_accel=False):
        if self._n_local == 0:
            return
        if self.use_noise_prior and self._prior is not None:
            self._prior.precond(amplitudes_in.local, amplitudes_in.local_flags,
                                amplitudes_out.local, use_accel=use_accel)
            return
        K.template_offset_apply_diag_precond(self._offsetvar, amplitudes_in.local,
                                             amplitudes_in.local_flags, amplitudes_out.local,
                                             use_accel)
