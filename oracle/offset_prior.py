"""oracle/offset_prior.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the Offset template's noise prior and banded / Toeplitz preconditioner
(``templates/offset/offset.py:203-222,356-712,884-1010`` of the reference).  The reference runs
these in Python on the host -- numpy + ``scipy.signal.convolve`` + ``scipy.linalg.cho_solve_banded``
-- and raises NotImplementedError on an accelerator (``offset.py:888-891,964-967``); this module
restates that algorithm with the same library routines in the same order, segment by segment
(each step cites the reference lines it follows), so that its results can be -- and are --
bit-identical to the reference's.

Parity status: PINNED.  ``tests/golden/make_golden_prior.py`` and
``tests/golden/make_golden_offset_init.py`` execute the reference's OWN method bodies from
``/root/reference`` (lifted out of offset.py with ``ast`` and bound to duck-typed stand-ins for
the template, the data, astropy units and the noise model): the PSD helpers, the whole
``_initialize`` (layout, amplitude flags / variance, filters, banded and Toeplitz preconditioners),
``_add_prior`` and ``_apply_precond``.  Every function of this module reproduces those outputs
bit for bit (tests/test_offset_prior.py; fixtures ``offset_prior.npz``, ``offset_template.npz``).

The product package ``toast_b200`` never imports this module.
"""

import numpy as np
import scipy.linalg
import scipy.optimize
import scipy.signal


# --------------------------------------------------------------------------------------------
# helpers: offset.py:589-712
# --------------------------------------------------------------------------------------------
def interpolate_psd(x, lfreq, lpsd):
    """offset.py:589-606: log-log interpolation, |x| < 1e-10 pinned to the threshold."""
    thresh = 1.0e-10
    lowf = np.abs(x) < thresh
    good = np.logical_not(lowf)
    logx = np.empty_like(x)
    logx[lowf] = np.log(thresh)
    logx[good] = np.log(np.abs(x[good]))
    return np.exp(np.interp(logx, lfreq, lpsd))


def truncate(noisefilter, lim=1e-4):
    """offset.py:608-616: cut the real-space filter where it falls below lim x its zero-lag
    value, odd half-width, rolled so that it is symmetric about its centre."""
    icenter = noisefilter.size // 2
    ind = np.abs(noisefilter[:icenter]) > np.abs(noisefilter[0]) * lim
    icut = np.argwhere(ind)[-1][0]
    if icut % 2 == 0:
        icut += 1
    noisefilter = np.roll(noisefilter, icenter)
    return noisefilter[icenter - icut:icenter + icut + 1]


def remove_white_noise(freq, psd):
    """offset.py:618-653: subtract the white plateau fitted to the top 20 % of the spectrum."""
    corrpsd = psd.copy()
    n_corrpsd = len(corrpsd)
    plat_off = int(0.8 * n_corrpsd)
    if n_corrpsd - plat_off < 10:
        if n_corrpsd < 10:
            plat_off = 0
        else:
            plat_off = n_corrpsd - 10
    cfreq = np.log(freq[plat_off:])
    cdata = np.log(corrpsd[plat_off:])

    def lin_func(x, a, b, c):
        return a * (x - b) + c

    params, _ = scipy.optimize.curve_fit(lin_func, cfreq, cdata, p0=[0.0, cfreq[-1], cdata[-1]])
    cdata = np.exp(lin_func(cfreq, params[0], params[1], params[2]))
    plat = cdata[-1]
    corrmax = np.amax(corrpsd)
    corrthresh = 1.0e-10 * corrmax - plat
    corrpsd -= plat
    corrpsd[corrpsd < corrthresh] = corrthresh
    return corrpsd


def get_offset_psd(psdfreq, psd, freq, step_time):
    """offset.py:655-712: PSD of the baseline offsets (Keihanen et al. 2010, eq. 22-24 with the
    reference's algebra correction), m = -4..4 aliases of the sinc^2 window."""
    psd = remove_white_noise(psdfreq, psd)
    logfreq = np.log(psdfreq)
    logpsd = np.log(psd)
    m_max = 5
    tbase = step_time
    fbase = 1.0 / tbase

    def g(f, m):
        x = np.pi * tbase * (f + m * fbase)
        bad = np.abs(x) < 1.0e-30
        good = np.logical_not(bad)
        result = np.empty_like(x)
        result[bad] = 1.0
        result[good] = (np.sin(x[good]) / x[good]) ** 2
        return result

    offset_psd = interpolate_psd(freq, logfreq, logpsd) * g(freq, 0)
    for m in range(1, m_max):
        offset_psd[:] += interpolate_psd(freq + m * fbase, logfreq, logpsd) * g(freq, m)
        offset_psd[:] += interpolate_psd(freq - m * fbase, logfreq, logpsd) * g(freq, -m)
    offset_psd *= fbase
    return offset_psd


def prior_frequencies(obstime, step_time, rate):
    """offset.py:203-222: log-spaced grid the offset PSD is tabulated on; None when the
    observation holds a single baseline (prior disabled)."""
    fbase = 1.0 / step_time
    if (obstime * fbase) < 1.0:
        return None
    powmin = np.floor(np.log10(1 / obstime)) - 1
    powmax = min(np.ceil(np.log10(1 / step_time)) + 2, np.log10(rate))
    return np.logspace(powmin, powmax, 1000)


# --------------------------------------------------------------------------------------------
# assembly: offset.py:356-560
# --------------------------------------------------------------------------------------------
class OraclePrior:
    """filters[d][v], precond[d][v] = (array, lower) in the reference's layout, for one
    observation with ``n_det`` detectors and the amplitude layout of ``oracle.offset_layout``."""

    def __init__(self, filters, precond, n_amp_views, precond_width):
        self.filters = filters
        self.precond = precond
        self.n_amp_views = np.asarray(n_amp_views, dtype=np.int64)
        self.precond_width = precond_width


def build_prior(psdfreq, psds, detnoise, offset_var, n_amp_views, obstime, step_time, rate,
                precond_width=20):
    """offset.py:356-560 for one observation.  ``psds[d]`` is the detector PSD on ``psdfreq``,
    ``detnoise[d]`` its detector weight, ``offset_var`` the diagonal amplitude variance
    (detector-major, views inside)."""
    freq = prior_frequencies(obstime, step_time, rate)
    if freq is None:
        return None
    n_det = len(psds)
    per_det = int(np.sum(n_amp_views))
    filters, precond = [], []
    offset = 0
    for d in range(n_det):
        offset_psd = get_offset_psd(psdfreq, np.asarray(psds[d], dtype=np.float64), freq,
                                    step_time)
        logfreq = np.log(freq)
        logpsd = np.log(offset_psd)
        logfilter = np.log(1.0 / offset_psd)
        fl, pl = [], []
        for ivw, n_amp_view in enumerate(n_amp_views):
            n_amp_view = int(n_amp_view)
            offsetvar_slice = offset_var[offset:offset + n_amp_view]
            filterlen = 2
            while filterlen < 2 * n_amp_view:
                filterlen *= 2
            filterfreq = np.fft.rfftfreq(filterlen, step_time)
            fourierfilter = interpolate_psd(filterfreq, logfreq, logfilter)
            noisefilter = truncate(np.fft.irfft(fourierfilter))
            fl.append(noisefilter)
            lower = None
            if precond_width == 1:
                # offset.py:485-499 (Toeplitz)
                preconditioner = truncate(np.fft.irfft(interpolate_psd(filterfreq, logfreq,
                                                                         logpsd)))
                icenter = preconditioner.size // 2
                if detnoise[d] != 0:
                    preconditioner[icenter] += 1.0 / detnoise[d]
            else:
                # offset.py:500-553 (banded Cholesky of M, not of M^-1)
                icenter = noisefilter.size // 2
                try_width = precond_width
                while True:
                    wband = min(try_width, icenter)
                    pw = max(wband, min(try_width, n_amp_view))
                    preconditioner = np.zeros([pw, n_amp_view], dtype=np.float64)
                    if detnoise[d] != 0:
                        preconditioner[0, :] = 1.0 / offsetvar_slice
                    preconditioner[:wband, :] += np.repeat(
                        noisefilter[icenter:icenter + wband, np.newaxis], n_amp_view, 1)
                    lower = True
                    try:
                        preconditioner = scipy.linalg.cholesky_banded(
                            preconditioner, overwrite_ab=True, lower=lower, check_finite=True)
                        break
                    except scipy.linalg.LinAlgError:
                        if try_width < icenter and try_width < n_amp_view:
                            try_width *= 2
                        else:
                            raise RuntimeError("cholesky_banded failed at the maximum width")
            pl.append((preconditioner, lower))
            offset += n_amp_view
        filters.append(fl)
        precond.append(pl)
    assert offset == per_det * n_det
    return OraclePrior(filters, precond, n_amp_views, precond_width)


# --------------------------------------------------------------------------------------------
# application: offset.py:884-960 (_add_prior) and :962-1010 (_apply_precond)
# --------------------------------------------------------------------------------------------
def add_prior(prior, amps_in, amp_flags, amps_out):
    offset = 0
    for d in range(len(prior.filters)):
        for ivw, n_amp_view in enumerate(prior.n_amp_views):
            sl = slice(offset, offset + int(n_amp_view))
            a_in = amps_in[sl]
            a_out = amps_out[sl]
            a_out[:] += scipy.signal.convolve(a_in, prior.filters[d][ivw], mode="same",
                                              method="auto")
            a_out[amp_flags[sl] != 0] = 0.0
            offset += int(n_amp_view)


def apply_precond(prior, amps_in, amp_flags, amps_out):
    offset = 0
    for d in range(len(prior.precond)):
        for ivw, n_amp_view in enumerate(prior.n_amp_views):
            sl = slice(offset, offset + int(n_amp_view))
            a_in = amps_in[sl]
            if prior.precond_width <= 1:
                out = scipy.signal.convolve(a_in, prior.precond[d][ivw][0], mode="same",
                                            method="auto")
            else:
                out = scipy.linalg.cho_solve_banded(prior.precond[d][ivw], a_in,
                                                    overwrite_b=False, check_finite=True)
            out[amp_flags[sl] != 0] = 0.0
            amps_out[sl] = out
            offset += int(n_amp_view)


def analytic_psd(sigma, rate, fknee=0.05, fmin=1.0e-5, alpha=1.0, n_freq=400):
    """A toast.noise_sim.AnalyticNoise-shaped spectrum for the tests:
    P(f) = NET^2 (f^alpha + fknee^alpha) / (f^alpha + fmin^alpha), NET^2 = sigma^2 / rate."""
    freq = np.logspace(np.log10(fmin), np.log10(rate / 2.0), n_freq)
    net2 = np.asarray(sigma, dtype=np.float64)[:, None] ** 2 / rate
    psd = net2 * (freq[None, :] ** alpha + fknee ** alpha) / (freq[None, :] ** alpha
                                                              + fmin ** alpha)
    return freq, psd
