"""PixelDistribution / PixelData (``pixels.py:59-241, 436-969``) for the hot path.

A map is ``[n_local_submap, n_pix_submap, n_value]`` over the locally hit submaps;
``global_submap_to_local`` is the int64 lookup (-1 = not local) the kernels take.  The map
reduction ``sync_allreduce`` is the collective of the path; on a GPU job it runs over NCCL on the
device-resident buffer (``toast_b200.solver.Destriper``), here it is the host-array form used by
the stand-alone operators.
"""

import numpy as np


class PixelDistribution:
    def __init__(self, n_pix, n_submap, local_submaps, comm=None):
        self.n_pix = int(n_pix)
        self.n_submap = int(n_submap)
        if self.n_pix % self.n_submap:
            self.n_pix_submap = self.n_pix // self.n_submap + 1
        else:
            self.n_pix_submap = self.n_pix // self.n_submap
        self.local_submaps = np.asarray(local_submaps, dtype=np.int64)
        self.n_local_submap = len(self.local_submaps)
        self.comm = comm
        self.nest = True
        g2l = np.full(self.n_submap, -1, dtype=np.int64)
        g2l[self.local_submaps] = np.arange(self.n_local_submap, dtype=np.int64)
        self.global_submap_to_local = g2l

    def global_pixel_to_submap(self, gl):
        """pixels.py:196-221 / _libtoast/pixels.cpp:10-41."""
        gl = np.asarray(gl, dtype=np.int64)
        good = gl >= 0
        gsm = np.where(good, gl // self.n_pix_submap, 0)
        sm = np.where(good, self.global_submap_to_local[gsm], -1).astype(np.int64)
        lp = np.where(good, gl - gsm * self.n_pix_submap, -1).astype(np.int64)
        return sm, lp


class PixelData:
    def __init__(self, dist, dtype=np.float64, n_value=1, units=None, zero=True):
        self.distribution = dist
        self.n_value = int(n_value)
        self.dtype = np.dtype(dtype)
        # zero=False: the caller overwrites every value (a device -> host copy of a finished map)
        alloc = np.zeros if zero else np.empty
        self.data = alloc((dist.n_local_submap, dist.n_pix_submap, self.n_value), dtype=dtype)
        self.raw = self.data.reshape(-1)
        self.units = units

    def reset(self):
        self.data[:] = 0

    def sync_allreduce(self, comm_bytes=10000000):
        """pixels.py:710-779: sum the map over all processes."""
        comm = self.distribution.comm
        if comm is None:
            return
        comm.allreduce_(self.data, op="sum")

    sync_alltoallv = sync_allreduce  # same result (tests/ops_mapmaker_utils.py:211-399)

    def _k(self):
        from . import _libtoast

        return _libtoast

    def accel_create(self, name="pixeldata"):
        self._k().accel_create(self.data, name)

    def accel_update_device(self, name="pixeldata"):
        self._k().accel_update_device(self.data, name)

    def accel_update_host(self, name="pixeldata"):
        self._k().accel_update_host(self.data, name)

    def accel_delete(self, name="pixeldata"):
        self._k().accel_delete(self.data, name)

    def accel_exists(self):
        return self._k().accel_present(self.data, "pixeldata")
