"""Amplitudes / AmplitudesMap (``templates/amplitudes.py:34-571, 804-975``): the local slice of
a distributed template-amplitude vector with uint8 flags, ``+= -= *=`` arithmetic, and a dot
product that ignores flagged entries and sums over processes."""

import numpy as np


class Amplitudes:
    def __init__(self, comm, n_global, n_local, local_indices=None, local_ranges=None,
                 dtype=np.float64):
        self._comm = comm
        self.n_global = int(n_global)
        self.n_local = int(n_local)
        self.local = np.zeros(self.n_local, dtype=dtype)
        self.local_flags = np.zeros(self.n_local, dtype=np.uint8)

    def reset(self):
        self.local[:] = 0

    def reset_flags(self):
        self.local_flags[:] = 0

    def duplicate(self):
        out = Amplitudes(self._comm, self.n_global, self.n_local)
        out.local[:] = self.local
        out.local_flags[:] = self.local_flags
        return out

    def __iadd__(self, other):
        self.local += other.local if isinstance(other, Amplitudes) else other
        return self

    def __isub__(self, other):
        self.local -= other.local if isinstance(other, Amplitudes) else other
        return self

    def __imul__(self, other):
        self.local *= other.local if isinstance(other, Amplitudes) else other
        return self

    def sync(self):
        """No-op for disjoint amplitudes such as Offset (amplitudes.py:375-379)."""
        return

    def dot(self, other):
        """amplitudes.py:523-571."""
        if other.n_global != self.n_global or other.n_local != self.n_local:
            raise RuntimeError("Amplitudes must have the same number of values")
        if self.n_global == 0:
            return 0.0
        local = 0.0
        if self.n_local > 0:
            local = np.dot(np.where(self.local_flags == 0, self.local, 0),
                           np.where(other.local_flags == 0, other.local, 0))
        if self._comm is None or self._comm.comm_world is None:
            return local
        buf = np.array([local], dtype=np.float64)
        self._comm.allreduce_(buf)
        return float(buf[0])

    # accelerator mirror
    def _k(self):
        from .. import _libtoast

        return _libtoast

    def accel_exists(self):
        return self._k().accel_present(self.local, "amplitudes")

    def accel_create(self, name="amplitudes"):
        self._k().accel_create(self.local, name)
        self._k().accel_create(self.local_flags, name + "_flags")

    def accel_update_device(self, name="amplitudes"):
        self._k().accel_update_device(self.local, name)
        self._k().accel_update_device(self.local_flags, name + "_flags")

    def accel_update_host(self, name="amplitudes"):
        self._k().accel_update_host(self.local, name)

    def accel_delete(self, name="amplitudes"):
        self._k().accel_delete(self.local, name)
        self._k().accel_delete(self.local_flags, name + "_flags")


class AmplitudesMap(dict):
    """Dictionary of Amplitudes keyed by template name (amplitudes.py:804-975)."""

    def reset(self):
        for v in self.values():
            v.reset()

    def duplicate(self):
        out = AmplitudesMap()
        for k, v in self.items():
            out[k] = v.duplicate()
        return out

    def dot(self, other):
        return sum(v.dot(other[k]) for k, v in self.items())

    def __iadd__(self, other):
        for k, v in self.items():
            v += other[k] if isinstance(other, AmplitudesMap) else other
        return self

    def __isub__(self, other):
        for k, v in self.items():
            v -= other[k] if isinstance(other, AmplitudesMap) else other
        return self

    def __imul__(self, other):
        for k, v in self.items():
            v *= other[k] if isinstance(other, AmplitudesMap) else other
        return self

    def accel_exists(self):
        return all(v.accel_exists() for v in self.values())

    def accel_create(self, name="amplitudes"):
        for k, v in self.items():
            v.accel_create(f"{name}_{k}")

    def accel_update_device(self, name="amplitudes"):
        for k, v in self.items():
            v.accel_update_device(f"{name}_{k}")

    def accel_update_host(self, name="amplitudes"):
        for k, v in self.items():
            v.accel_update_host(f"{name}_{k}")

    def accel_delete(self, name="amplitudes"):
        for k, v in self.items():
            v.accel_delete(f"{name}_{k}")
