"""Generate tests/golden/offset_prior.npz by EXECUTING the reference's own PSD helper methods of
the Offset template (``/root/reference/src/toast/templates/offset/offset.py``: ``_interpolate_psd``,
``_truncate``, ``_remove_white_noise``, ``_get_offset_psd``).

``import toast`` is impossible in this container (astropy, traitlets, ... are missing), but these
four methods only need numpy / scipy, ``self.step_time`` / ``self.det_data_units`` and a noise
model with ``freq(det)`` / ``psd(det)`` quantities.  Their function definitions are taken out of
the reference source with ``ast`` -- executed where they lie, nothing is copied into this
repository -- and bound to a stand-in object.  Run in the build container only:

    python tests/golden/make_golden_prior.py
"""

import ast
import os
import sys
import types

import numpy as np
import scipy
import scipy.optimize

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/toast/templates/offset/offset.py"
WANT = ["_interpolate_psd", "_truncate", "_remove_white_noise", "_get_offset_psd"]


class _Unit:
    """Stand-in for astropy units: every unit is 1."""

    def __pow__(self, p):
        return self

    def __mul__(self, o):
        return self

    __rmul__ = __mul__


class _Quantity:
    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)

    def to_value(self, unit):
        return self.a.copy()


class _Noise:
    def __init__(self, freq, psds):
        self._f, self._p = freq, psds

    def freq(self, det):
        return _Quantity(self._f)

    def psd(self, det):
        return _Quantity(self._p[det])


def load_reference_methods():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Offset")
    funcs = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    for f in funcs:
        f.decorator_list = []
    mod = ast.Module(body=funcs, type_ignores=[])
    ns = {"np": np, "scipy": scipy, "u": types.SimpleNamespace(Hz=_Unit(), second=_Unit())}
    exec(compile(mod, REF, "exec"), ns)
    return {name: ns[name] for name in WANT}


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import offset_prior as OP

    m = load_reference_methods()
    self = types.SimpleNamespace(det_data_units=_Unit())
    for name, fn in m.items():
        setattr(self, name, types.MethodType(fn, self))

    rate, step_time, obstime = 10.0, 2.0, 600.0
    sigma = np.array([1.0, 1.07, 1.03])
    psdfreq, psds = OP.analytic_psd(sigma, rate, fknee=0.05, fmin=1e-4, alpha=1.5, n_freq=300)
    noise = _Noise(psdfreq, psds)
    freq = OP.prior_frequencies(obstime, step_time, rate)

    out = dict(rate=rate, step_time=step_time, obstime=obstime, sigma=sigma, psdfreq=psdfreq,
               psds=psds, freq=freq)
    out["corrpsd"] = np.stack([self._remove_white_noise(psdfreq, psds[d].copy())
                               for d in range(3)])
    out["offset_psd"] = np.stack([self._get_offset_psd(noise, freq, step_time, d)
                                  for d in range(3)])
    x = np.concatenate([[0.0, 1e-12, -3e-3], np.linspace(-0.4, 0.6, 41)])
    out["interp_x"] = x
    out["interp_y"] = self._interpolate_psd(x, np.log(freq), np.log(out["offset_psd"][0]))
    # real-space filter of the 1/PSD spectrum for a 300-baseline view, before / after truncation
    filterlen = 1024
    ff = np.fft.rfftfreq(filterlen, step_time)
    raw = np.fft.irfft(self._interpolate_psd(ff, np.log(freq),
                                             np.log(1.0 / out["offset_psd"][1])))
    out["filter_raw"] = raw
    out["filter_truncated"] = self._truncate(raw.copy())
    out["filter_truncated_1e2"] = self._truncate(raw.copy(), lim=1e-2)
    np.savez_compressed(os.path.join(HERE, "offset_prior.npz"), **out)
    print("wrote offset_prior.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
