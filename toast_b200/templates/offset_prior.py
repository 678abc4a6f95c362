"""Noise prior of the Offset template: host-side assembly, device-side application.

Reference: ``templates/offset/offset.py``.  The ASSEMBLY (``:203-222`` frequency grid, ``:356-560``
filters and preconditioners, ``:589-712`` PSD helpers) is host Python there and stays host numpy /
scipy here -- it runs once per solve.  The APPLICATION runs every PCG iteration: the reference
loops over (detector, observation, view) segments in Python calling ``scipy.signal.convolve`` /
``scipy.linalg.cho_solve_banded`` and has no accelerator form (``:888-891``, ``:964-967``); here all
segments go through one CUDA launch each (``tb_offset_prior_add`` / ``tb_offset_prior_precond``,
``csrc/tb_prior.cu``).
"""

import ctypes as ct

import numpy as np
import scipy.linalg
import scipy.optimize

from .. import lib as L

_LOW_F = 1.0e-10


def _loglog_interp(x, logfreq, logval):
    """offset.py:589-606: log-log interpolation, symmetric in f, |f| < 1e-10 pinned."""
    ax = np.maximum(np.abs(x), _LOW_F)
    return np.exp(np.interp(np.log(ax), logfreq, logval))


def _symmetric_cut(kernel, lim=1.0e-4):
    """offset.py:608-616: keep the lags up to the last one above lim x the zero-lag value (odd
    half-width), centred."""
    half = kernel.size // 2
    above = np.flatnonzero(np.abs(kernel[:half]) > np.abs(kernel[0]) * lim)
    cut = int(above[-1])
    cut += 1 - cut % 2
    centred = np.roll(kernel, half)
    return centred[half - cut:half + cut + 1]


def _correlated_part(freq, psd):
    """offset.py:618-653: the PSD minus its white plateau (a line fitted in log-log to the top
    20 % of the spectrum, evaluated at the last frequency), floored at 1e-10 x max - plateau."""
    n = len(psd)
    first = int(0.8 * n)
    if n - first < 10:
        first = 0 if n < 10 else n - 10
    lx, ly = np.log(freq[first:]), np.log(psd[first:])

    def line(x, a, b, c):
        return a * (x - b) + c

    # same call and starting point as the reference: the three-parameter line is degenerate, so
    # the converged plateau depends on them at the 1e-8 level
    par, _ = scipy.optimize.curve_fit(line, lx, ly, p0=[0.0, lx[-1], ly[-1]])
    plateau = np.exp(line(lx, *par))[-1]
    out = psd - plateau
    floor = 1.0e-10 * np.amax(psd) - plateau
    out[out < floor] = floor
    return out


def offset_psd(psd_freq, psd, freq, step_time, m_max=5):
    """offset.py:655-712: P_a(f) = (1/T) sum_m P(f + m/T) sinc^2(pi T (f + m/T)), |m| < m_max,
    of the correlated part of the detector PSD."""
    lf, lp = np.log(psd_freq), np.log(_correlated_part(psd_freq, np.array(psd, dtype=np.float64)))
    fbase = 1.0 / step_time

    def window(f, m):
        x = np.pi * step_time * (f + m * fbase)
        small = np.abs(x) < 1.0e-30
        xs = np.where(small, 1.0, x)
        return np.where(small, 1.0, (np.sin(xs) / xs) ** 2)

    total = _loglog_interp(freq, lf, lp) * window(freq, 0)
    for m in range(1, m_max):
        total += _loglog_interp(freq + m * fbase, lf, lp) * window(freq, m)
        total += _loglog_interp(freq - m * fbase, lf, lp) * window(freq, -m)
    return total * fbase


def prior_frequencies(obstime, step_time, rate):
    """offset.py:203-222; None = a single baseline in the observation, prior disabled."""
    if obstime / step_time < 1.0:
        return None
    lo = np.floor(np.log10(1.0 / obstime)) - 1
    hi = min(np.ceil(np.log10(1.0 / step_time)) + 2, np.log10(rate))
    return np.logspace(lo, hi, 1000)


class OffsetPriorBuilder:
    """Collects the segments of every (detector, observation) and their filters / factors."""

    def __init__(self, n_amp, precond_width=20):
        self.n_amp = int(n_amp)
        self.precond_width = int(precond_width)
        self.seg_start, self.seg_len = [], []
        self.filt_start, self.filt_len, self.filters = [], [], []
        self.prec_start, self.prec_width, self.precond = [], [], []
        self._nf = 0
        self._np = 0

    def add_cut_detector(self, amp_start, n_amp_views):
        """A detector without a noise filter (offset.py:946-947, 1003-1005): zeros."""
        off = int(amp_start)
        for na in n_amp_views:
            self.seg_start.append(off)
            self.seg_len.append(int(na))
            self.filt_start.append(-1)
            self.filt_len.append(0)
            self.prec_start.append(-1)
            self.prec_width.append(0)
            off += int(na)

    def add_detector(self, amp_start, n_amp_views, psd_freq, psd, detnoise, offset_var, freq,
                     step_time):
        """offset.py:360-560 for one detector of one observation: per view the truncated
        real-space 1/PSD filter and the banded Cholesky factor (or Toeplitz kernel) of
        diag(1/offset_var) + filter."""
        opsd = offset_psd(psd_freq, psd, freq, step_time)
        lf, lpsd, lfilt = np.log(freq), np.log(opsd), np.log(1.0 / opsd)
        cache = {}
        off = int(amp_start)
        for na in n_amp_views:
            na = int(na)
            flen = 2
            while flen < 2 * na:
                flen *= 2
            if flen not in cache:
                ff = np.fft.rfftfreq(flen, step_time)
                filt = _symmetric_cut(np.fft.irfft(_loglog_interp(ff, lf, lfilt)))
                toep = None
                if self.precond_width == 1:
                    toep = _symmetric_cut(np.fft.irfft(_loglog_interp(ff, lf, lpsd)))
                    if detnoise != 0:
                        toep[toep.size // 2] += 1.0 / detnoise
                cache[flen] = (filt, toep)
            filt, toep = cache[flen]
            self.seg_start.append(off)
            self.seg_len.append(na)
            self.filt_start.append(self._nf)
            self.filt_len.append(filt.size)
            self.filters.append(filt)
            self._nf += filt.size
            if self.precond_width == 1:
                pre = toep
                width = toep.size
            else:
                pre = self._banded_factor(filt, offset_var[off:off + na], detnoise, na)
                width = pre.shape[0]
            self.prec_start.append(self._np)
            self.prec_width.append(width)
            self.precond.append(np.ascontiguousarray(pre).reshape(-1))
            self._np += pre.size
            off += na

    def _banded_factor(self, filt, var, detnoise, na):
        """offset.py:500-553: lower banded Cholesky factor of M = diag(1/var) + Toeplitz(filter),
        the band doubled until the factorisation succeeds."""
        centre = filt.size // 2
        width = self.precond_width
        while True:
            wband = min(width, centre)
            rows = max(wband, min(width, na))
            ab = np.zeros((rows, na))
            if detnoise != 0:
                ab[0] = 1.0 / var
            ab[:wband] += filt[centre:centre + wband, None]
            try:
                return scipy.linalg.cholesky_banded(ab, overwrite_ab=True, lower=True,
                                                    check_finite=True)
            except scipy.linalg.LinAlgError:
                if width < centre and width < na:
                    width *= 2
                else:
                    raise RuntimeError("Offset noise prior: banded Cholesky failed at the "
                                       "maximum band width")

    def finish(self):
        return OffsetPrior(self)


class OffsetPrior:
    """Device-resident noise prior: ``add`` is Offset._add_prior, ``precond`` is
    Offset._apply_precond (use_noise_prior=True).  Arguments are numpy arrays (staged through
    the call: ``use_accel=False``; or registered in the accel table: ``use_accel=True``) or CUDA
    tensors."""

    def __init__(self, b):
        i64 = lambda v: np.ascontiguousarray(v, dtype=np.int64)
        order = np.argsort(i64(b.seg_start), kind="stable")
        self._arrays = [i64(b.seg_start)[order], i64(b.seg_len)[order], i64(b.filt_start)[order],
                        i64(b.filt_len)[order], i64(b.prec_start)[order],
                        i64(b.prec_width)[order],
                        np.concatenate(b.filters) if b.filters else np.zeros(0),
                        np.concatenate(b.precond) if b.precond else np.zeros(0)]
        a = self._arrays
        d = L.tb_offset_prior_desc()
        d.n_amp, d.n_seg = b.n_amp, len(a[0])
        d.seg_start, d.seg_len = a[0].ctypes.data, a[1].ctypes.data
        d.filt_start, d.filt_len = a[2].ctypes.data, a[3].ctypes.data
        d.filters, d.n_filter_values = a[6].ctypes.data, a[6].size
        d.precond_mode = L.TB_PRECOND_TOEPLITZ if b.precond_width == 1 else L.TB_PRECOND_BANDED
        d.prec_start, d.prec_width = a[4].ctypes.data, a[5].ctypes.data
        d.precond, d.n_precond_values = a[7].ctypes.data, a[7].size
        self.n_amp = b.n_amp
        self.n_segments = len(a[0])
        self.lib = L.load()
        self.h = self.lib.tb_offset_prior_create(ct.byref(d))
        if not self.h:
            raise RuntimeError(L.last_error())

    def _call(self, fn, amps_in, flags, amps_out, use_accel, stream):
        for v in (amps_in, flags, amps_out):
            n = v.numel() if L._is_tensor(v) else v.size
            if n != self.n_amp:
                raise RuntimeError("amplitude vector length does not match the prior")
        L.check(fn(self.h, L.ptr(amps_in), L.ptr(flags), L.ptr(amps_out),
                   L.mem_of(amps_out, use_accel), stream))

    def add(self, amps_in, flags, amps_out, use_accel=False, stream=None):
        self._call(self.lib.tb_offset_prior_add, amps_in, flags, amps_out, use_accel, stream)

    def precond(self, amps_in, flags, amps_out, use_accel=False, stream=None):
        self._call(self.lib.tb_offset_prior_precond, amps_in, flags, amps_out, use_accel, stream)

    def __del__(self):
        try:
            self.lib.tb_offset_prior_destroy(self.h)
        except Exception:
            pass
