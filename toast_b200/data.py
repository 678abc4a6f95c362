"""Minimal host-side data model in the reference's layouts.

Mirrors the parts of ``toast.Data`` / ``Observation`` / ``DetectorData`` / ``IntervalList`` that
the hot-path operators touch (``observation_data.py:35-603,725-860``, ``intervals.py:48``):
detector data is ONE C-contiguous ``[n_det, n_samp, *sample_shape]`` buffer addressed through
int32 row indices, shared data is ``[n_samp, ...]``, a view is an ``Interval`` array whose
``last`` is exclusive, and ``ensure()`` reports whether an output already exists (operators then
skip the work, ``pixels_healpix.py:215-243``).  Instrument, schedule, I/O and MPI layers are out
of scope; ``Comm`` reads the torch.distributed world.
"""

import numpy as np

from .lib import interval_dtype


def make_intervals(ranges):
    iv = np.zeros(len(ranges), dtype=interval_dtype)
    for i, (a, b) in enumerate(ranges):
        iv[i] = (float(a), float(b), int(a), int(b))
    return iv


class Comm:
    """World communicator facts (toast.mpi.Comm, mpi.py:113-143) from torch.distributed."""

    def __init__(self, group=None):
        self.group = group
        self.world_rank, self.world_size = 0, 1
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self.world_rank = dist.get_rank(group)
                self.world_size = dist.get_world_size(group)
        except ImportError:
            pass
        self.comm_world = None if self.world_size == 1 else self

    # the two collectives the hot path needs, on host numpy arrays (pixels.py:710-779,
    # templates/amplitudes.py:560-571)
    def allreduce_(self, array, op="sum"):
        if self.world_size == 1:
            return array
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(array)
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX,
                        group=self.group)
        array[...] = t.cpu().numpy()
        return array


class DetectorData:
    """observation_data.py:35-603: rows of one contiguous buffer, one row per detector."""

    def __init__(self, detectors, shape, dtype, pinned=False):
        self.detectors = list(detectors)
        self._index = {d: i for i, d in enumerate(self.detectors)}
        full = (len(self.detectors),) + tuple(shape)
        self._pinned = None
        if pinned:
            # page-locked host buffer (what the reference's accelerator mapping needs for
            # asynchronous copies); numpy view of a pinned torch tensor, kept alive here
            import torch

            self._pinned = torch.zeros(full, dtype=getattr(torch, np.dtype(dtype).name),
                                       pin_memory=True)
            self.data = self._pinned.numpy()
        else:
            self.data = np.zeros(full, dtype=dtype)
        self.dtype = np.dtype(dtype)
        self.detector_shape = tuple(shape)

    def indices(self, dets):
        return np.array([self._index[d] for d in dets], dtype=np.int32)

    def __getitem__(self, key):
        if isinstance(key, tuple):
            det, rest = key[0], key[1:]
            if isinstance(det, str):
                det = self._index[det]
            return self.data[(det,) + rest]
        if isinstance(key, str):
            return self.data[self._index[key]]
        return self.data[key]

    def __setitem__(self, key, value):
        if isinstance(key, tuple) and isinstance(key[0], str):
            key = (self._index[key[0]],) + key[1:]
        elif isinstance(key, str):
            key = self._index[key]
        self.data[key] = value

    # accelerator mirror (observation_data.py:559-603 -> accel.py:374-484)
    def _k(self):
        from . import _libtoast

        return _libtoast

    def accel_exists(self):
        return self._k().accel_present(self.data, "detdata")

    def accel_create(self, name="detdata"):
        self._k().accel_create(self.data, name)

    def accel_update_device(self, name="detdata"):
        self._k().accel_update_device(self.data, name)

    def accel_update_host(self, name="detdata"):
        self._k().accel_update_host(self.data, name)

    def accel_reset(self, name="detdata"):
        self._k().accel_reset(self.data, name)

    def accel_delete(self, name="detdata"):
        self._k().accel_delete(self.data, name)


class DetDataManager(dict):
    """observation_data.py:725-860."""

    def __init__(self, n_samp):
        super().__init__()
        self.n_samp = n_samp

    def ensure(self, name, sample_shape=(), dtype=np.float64, detectors=None, accel=False):
        """Create ``name`` if needed.  Returns True if it already existed with every requested
        detector (callers then skip recomputation)."""
        shape = (self.n_samp,) + tuple(sample_shape)
        if name in self:
            dd = self[name]
            if dd.detector_shape != shape or dd.dtype != np.dtype(dtype):
                raise RuntimeError(f"detdata '{name}' exists with a different shape or dtype")
            if all(d in dd._index for d in detectors):
                return True
            raise RuntimeError(f"detdata '{name}' exists with a different detector set")
        self[name] = DetectorData(detectors, shape, dtype)
        if accel:
            # observation_data.py `ensure(..., accel=True)`: create on the device, zero-filled
            self[name].accel_create(name)
            self[name].accel_reset(name)
        return False


class NoiseModel:
    """Only what the hot path reads: the per-detector inverse-variance weight
    (noise.py ``detector_weight``) and -- for the Offset noise prior -- the PSD of each detector
    (``freq(det)`` in Hz, ``psd(det)`` in signal-units^2 s; plain arrays, no astropy units)."""

    def __init__(self, weights, freqs=None, psds=None):
        self._w = dict(weights)
        self._f = dict(freqs) if freqs is not None else None
        self._p = dict(psds) if psds is not None else None

    def detector_weight(self, det):
        return self._w[det]

    def freq(self, det):
        if self._f is None:
            raise RuntimeError("this noise model carries no PSDs")
        return self._f[det]

    def psd(self, det):
        if self._p is None:
            raise RuntimeError("this noise model carries no PSDs")
        return self._p[det]


class Observation:
    def __init__(self, name, detectors, n_samp):
        self.name = name
        self.local_detectors = list(detectors)
        self.n_local_samples = int(n_samp)
        self.shared = {}
        self.detdata = DetDataManager(n_samp)
        self._intervals = {None: make_intervals([(0, n_samp)])}
        self._meta = {}
        self.det_flags = {d: 0 for d in detectors}  # per-detector cut flags (det_mask)

    @property
    def intervals(self):
        return self._intervals

    def select_local_detectors(self, selection=None, flagmask=0):
        out = []
        for d in self.local_detectors:
            if selection is not None and d not in selection:
                continue
            if self.det_flags.get(d, 0) & flagmask:
                continue
            out.append(d)
        return out

    def __getitem__(self, key):
        return self._meta[key]

    def __setitem__(self, key, value):
        self._meta[key] = value

    def __contains__(self, key):
        return key in self._meta


class Data:
    """Container of observations plus named global objects (maps, amplitudes, ...)."""

    def __init__(self, comm=None):
        self.comm = comm if comm is not None else Comm()
        self.obs = []
        self._meta = {}

    def __getitem__(self, key):
        return self._meta[key]

    def __setitem__(self, key, value):
        self._meta[key] = value

    def __delitem__(self, key):
        del self._meta[key]

    def __contains__(self, key):
        return key in self._meta

    def all_local_detectors(self, selection=None, flagmask=0):
        seen = {}
        for ob in self.obs:
            for d in ob.select_local_detectors(selection, flagmask):
                seen[d] = None
        return list(seen)


def observation_from_synthetic(obs, name="obs0", det_prefix="D", pinned=False):
    """Wrap a ``toast_b200.synthetic.make_observation`` dict as an Observation with the
    reference's default keys (``defaults.py``): boresight_radec, flags, signal, noise_model."""
    n_det, n_samp = obs["n_det"], obs["n_samp"]
    dets = [f"{det_prefix}{i:05d}" for i in range(n_det)]
    ob = Observation(name, dets, n_samp)
    ob.shared["boresight_radec"] = obs["boresight"]
    ob.shared["flags"] = obs["shared_flags"]
    ob.shared["times"] = np.arange(n_samp, dtype=np.float64) / obs["rate"]
    ob.intervals["scanning"] = obs["intervals"]
    ob.detdata["flags"] = DetectorData(dets, (n_samp,), np.uint8, pinned=pinned)
    ob.detdata["flags"].data[:] = obs["det_flags"]
    if "signal" in obs:
        ob.detdata["signal"] = DetectorData(dets, (n_samp,), np.float64, pinned=pinned)
        ob.detdata["signal"].data[:] = obs["signal"]
    ob["noise_model"] = NoiseModel({d: float(w) for d, w in zip(dets, obs["detweight"])})
    ob["focalplane"] = {
        d: dict(quat=obs["focalplane"][i], pol_efficiency=None, epsilon=float(obs["epsilon"][i]),
                gamma=float(obs["gamma"][i]), cal=float(obs["cal"][i]))
        for i, d in enumerate(dets)
    }
    ob["rate"] = obs["rate"]
    return ob
