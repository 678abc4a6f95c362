"""Generate tests/golden/offset_template.npz by EXECUTING the reference's own
``templates.Offset`` methods -- ``_initialize`` (amplitude layout, amplitude flags and variance,
noise filters and preconditioners: offset.py:123-586), ``_add_prior`` (:884-960) and
``_apply_precond`` (:962-1010) -- from ``/root/reference/src/toast/templates/offset/offset.py``.

``import toast`` is impossible in this container; the method definitions are lifted out of the
reference source with ``ast`` (executed where they lie, nothing is copied into this repository)
and bound to duck-typed stand-ins for the template object, the data, the observations, astropy
units and the noise model.  Run in the build container only:

    python tests/golden/make_golden_offset_init.py
"""

import ast
import os
import re
import sys
import types
from collections import OrderedDict

import numpy as np
import scipy
import scipy.linalg
import scipy.optimize
import scipy.signal

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src/toast"
METHODS = ["_initialize", "_step_length", "_interpolate_psd", "_truncate", "_remove_white_noise",
           "_get_offset_psd", "_add_prior", "_apply_precond"]


# ---- stand-ins ----------------------------------------------------------------------------------
class Unit:
    def __pow__(self, p):
        return self

    def __mul__(self, o):
        return self

    __rmul__ = __mul__

    def __rtruediv__(self, o):
        return self


class Quantity:
    def __init__(self, v):
        self.v = v

    def to_value(self, unit):
        return self.v.copy() if isinstance(self.v, np.ndarray) else self.v


class Quiet:
    def __getattr__(self, name):
        return lambda *a, **k: None


class Logger:
    @staticmethod
    def get():
        return Quiet()


class AlignedF64:
    def __init__(self, n):
        self._a = np.zeros(n)

    @staticmethod
    def zeros(n):
        return AlignedF64(n)

    def array(self):
        return self._a


class Interval:
    def __init__(self, first, last):
        self.first, self.last = int(first), int(last)


class IntervalList(list):
    @property
    def data(self):
        return self


class Noise:
    def __init__(self, dets, weights, freq, psds):
        self._w = dict(zip(dets, weights))
        self._f = freq
        self._p = dict(zip(dets, psds))

    def detector_weight(self, det):
        return Quantity(float(self._w[det]))

    def freq(self, det):
        return Quantity(self._f)

    def psd(self, det):
        return Quantity(self._p[det])


class DetData:
    def __init__(self, dets, data):
        self.detectors = list(dets)
        self.data = data

    def __getitem__(self, key):
        det, slc = key
        return self.data[self.detectors.index(det), slc]


class Obs:
    def __init__(self, name, dets, n_samp, times, view_ranges, flags, noise):
        self.name = name
        self.local_detectors = list(dets)
        self.n_local_samples = n_samp
        self.shared = {"times": times}
        self.intervals = {None: IntervalList([Interval(0, n_samp)]),
                          "scanning": IntervalList([Interval(a, b) for a, b in view_ranges])}
        self.detdata = {"signal": DetData(dets, np.zeros((len(dets), 1))),
                        "flags": DetData(dets, flags)}
        self._meta = {"noise_model": noise}

    def select_local_detectors(self, selection=None, flagmask=0):
        return [d for d in self.local_detectors if selection is None or d in selection]

    def __getitem__(self, k):
        return self._meta[k]

    def __contains__(self, k):
        return k in self._meta


class Data:
    def __init__(self, obs):
        self.obs = obs
        self.comm = types.SimpleNamespace(comm_world=None, world_rank=0)


class Amps:
    def __init__(self, local, flags):
        self.local, self.local_flags = local, flags


def load_reference():
    tree = ast.parse(open(f"{REF}/templates/offset/offset.py").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Offset")
    funcs = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
    for f in funcs:
        f.decorator_list = []
    utree = ast.parse(open(f"{REF}/utils.py").read())
    funcs.append(next(n for n in utree.body
                      if isinstance(n, ast.FunctionDef) and n.name == "rate_from_times"))
    ns = {"np": np, "scipy": scipy, "os": os, "re": re, "OrderedDict": OrderedDict,
          "u": types.SimpleNamespace(Hz=Unit(), second=Unit()), "Logger": Logger,
          "AlignedF64": AlignedF64, "MPI": None, "set_matplotlib_backend": lambda: None}
    exec(compile(ast.Module(body=funcs, type_ignores=[]), REF, "exec"), ns)
    return ns


def make_template(ns, **traits):
    t = types.SimpleNamespace(
        name="baselines", view="scanning", times="times", det_data="signal", det_mask=1,
        det_flags=None, det_flag_mask=1, good_fraction=0.5, noise_model="noise_model",
        use_noise_prior=False, precond_width=20, pattern=None, debug_plots=None,
        det_data_units=Unit(), clear=lambda: None)
    t.__dict__.update(traits)
    for m in METHODS:
        setattr(t, m, types.MethodType(ns[m], t))
    return t


def synthetic_obs(name, n_det, n_samp, nside=64):
    from oracle import offset_prior as OP
    from toast_b200 import synthetic as S

    obs = S.make_observation(name, n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03)
    dets = [f"D{i:05d}" for i in range(n_det)]
    freq, psds = OP.analytic_psd(obs["sigma"], obs["rate"], fknee=0.05, fmin=1e-4, alpha=1.5,
                                 n_freq=300)
    times = np.arange(n_samp, dtype=np.float64) / obs["rate"]
    ranges = [(int(v["first"]), int(v["last"])) for v in obs["intervals"]]
    ob = Obs("obs0", dets, n_samp, times, ranges, obs["det_flags"],
             Noise(dets, obs["detweight"], freq, psds))
    return obs, ob, dets, freq, psds


def run_case(ns, tag, name, n_det, n_samp, out, **traits):
    obs, ob, dets, freq, psds = synthetic_obs(name, n_det, n_samp)
    t = make_template(ns, step_time=Quantity(float(obs["step_time"])), **traits)
    data = Data([ob])
    t._initialize(data)
    t.data = data
    out[f"{tag}_args"] = np.array([n_det, n_samp])
    out[f"{tag}_n_amp_views"] = t._obs_views[0]
    out[f"{tag}_det_start"] = np.array([t._det_start[d] for d in dets], dtype=np.int64)
    out[f"{tag}_amp_flags"] = np.asarray(t._amp_flags, dtype=np.uint8)
    out[f"{tag}_offset_var"] = np.asarray(t._offsetvar)
    out[f"{tag}_rate"] = np.array(t._obs_rate[0])
    if not t.use_noise_prior:
        return
    out[f"{tag}_freq"] = t._freq[0]
    for i, d in enumerate(dets):
        for v, f in enumerate(t._filters[0][d]):
            out[f"{tag}_filter_{i}_{v}"] = f
            out[f"{tag}_precond_{i}_{v}"] = np.asarray(t._precond[0][d][v][0])
    rng = np.random.default_rng(21)
    n = t._n_local
    a_in = rng.standard_normal(n)
    flags = np.where(t._amp_flags, 1, 0).astype(np.uint8)
    flags[::7] = 1 if t.precond_width == 1 else flags[::7]
    a_out = rng.standard_normal(n)
    out[f"{tag}_amps_in"], out[f"{tag}_flags_in"] = a_in, flags
    out[f"{tag}_amps_out0"] = a_out.copy()
    t._add_prior(Amps(a_in, flags), Amps(a_out, flags))
    out[f"{tag}_add_prior"] = a_out
    pre = Amps(np.zeros(n), flags)
    t._apply_precond(Amps(a_in, flags), pre)
    out[f"{tag}_apply_precond"] = pre.local
    print(tag, "n_amp", n, "views", t._obs_views[0], "flagged", int(np.sum(t._amp_flags)),
          "filter len", [len(f) for f in t._filters[0][dets[0]]])


def main():
    ns = load_reference()
    out = {}
    # plain template: the view defines the baseline boundaries, detector flags cut baselines
    run_case(ns, "plain_c2", "c2", 4, 12000, out, det_flags="flags")
    run_case(ns, "plain_c1", "c1", 4, 6000, out, det_flags="flags")
    # noise prior: baselines span the observation (offset.py:136-141), banded preconditioner
    run_case(ns, "prior_banded_c1", "c1", 3, 6000, out, use_noise_prior=True, precond_width=20)
    run_case(ns, "prior_banded4_c4", "c4", 2, 20000, out, use_noise_prior=True, precond_width=4)
    # Toeplitz preconditioner, with detector flags and a view that excludes the turnarounds
    run_case(ns, "prior_toeplitz_c2", "c2", 3, 12000, out, use_noise_prior=True,
             precond_width=1, det_flags="flags")
    np.savez_compressed(os.path.join(HERE, "offset_template.npz"), **out)
    print("wrote offset_template.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
