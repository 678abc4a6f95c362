"""PointingDetectorSimple, PixelsHealpix, StokesWeights
(``ops/pointing_detector/pointing_detector.py:117-290``, ``ops/pixels_healpix/pixels_healpix.py:
140-296``, ``ops/stokes_weights/stokes_weights.py:138-288``) on the CUDA kernels."""

import numpy as np

from .. import _libtoast as K
from ..pixels import PixelDistribution
from .operator import Operator


def _view_intervals(ob, view):
    return ob.intervals[view]


class PointingDetectorSimple(Operator):
    _defaults = dict(view=None, shared_flags="flags", shared_flag_mask=1, det_mask=1,
                     boresight="boresight_radec", quats="quats", focalplane_key="focalplane")

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        for ob in data.obs:
            dets = ob.select_local_detectors(detectors, flagmask=self.det_mask)
            if len(dets) == 0:
                continue
            exists = ob.detdata.ensure(self.quats, sample_shape=(4,), dtype=np.float64,
                                       detectors=ob.select_local_detectors(None, self.det_mask),
                                       accel=use_accel)
            if exists and getattr(self, "_done", {}).get((id(ob), tuple(dets))):
                continue  # pointing_detector.py:207-214
            fp = ob[self.focalplane_key]
            fp_quats = np.array([fp[d]["quat"] for d in dets], dtype=np.float64)
            flags = ob.shared[self.shared_flags] if self.shared_flags is not None else \
                np.zeros(1, dtype=np.uint8)
            K.pointing_detector(fp_quats, ob.shared[self.boresight],
                                ob.detdata[self.quats].indices(dets), ob.detdata[self.quats].data,
                                _view_intervals(ob, self.view), flags, self.shared_flag_mask,
                                use_accel)
            self.__dict__.setdefault("_done", {})[(id(ob), tuple(dets))] = True

    def _requires(self):
        req = {"shared": [self.boresight], "detdata": [], "intervals": []}
        if self.shared_flags is not None:
            req["shared"].append(self.shared_flags)
        return req

    def _provides(self):
        return {"detdata": [self.quats]}


class PointingDetectorFP(Operator):
    """``ops/pointing_detector_fp.py:14-139``: detector pointing in the focalplane frame -- every
    sample of a detector gets its focalplane quaternion (boresight constantly at the zenith).
    Host-only in the reference too (it has no kernel and no accelerator support); same traits,
    ``boresight`` / ``coord_in`` / ``coord_out`` accepted and ignored as there."""

    _defaults = dict(view=None, shared_flags="flags", shared_flag_mask=1, det_mask=1,
                     boresight=None, quats="quats", coord_in=None, coord_out=None,
                     focalplane_key="focalplane")

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        import warnings

        for trait in ("boresight", "coord_in", "coord_out"):
            if getattr(self, trait) is not None:
                warnings.warn(f"PointingDetectorFP will not use the provided {trait} = "
                              f"{getattr(self, trait)}")  # pointing_detector_fp.py:86-93
        for ob in data.obs:
            dets = ob.select_local_detectors(detectors, flagmask=self.det_mask)
            if len(dets) == 0:
                continue
            exists = ob.detdata.ensure(self.quats, sample_shape=(4,), dtype=np.float64,
                                       detectors=dets)
            if exists:
                continue  # pointing_detector_fp.py:104-111
            fp = ob[self.focalplane_key]
            qd = ob.detdata[self.quats]
            for det in dets:
                qd[det] = np.asarray(fp[det]["quat"], dtype=np.float64)

    def supports_accel(self):
        return False

    def _requires(self):
        return {"meta": [], "shared": [], "detdata": [], "intervals": []}

    def _provides(self):
        return {"meta": [], "shared": [], "detdata": [self.quats]}


class PixelsHealpix(Operator):
    _defaults = dict(detector_pointing=None, nside=64, nside_submap=16, nest=True, view=None,
                     pixels="pixels", create_dist=None, single_precision=False)

    def _geometry(self):
        # pixels_healpix.py:122-137
        nside_submap = min(self.nside_submap, self.nside)
        self._n_pix = 12 * self.nside**2
        self._n_pix_submap = 12 * nside_submap**2
        self._n_submap = (self.nside // nside_submap) ** 2

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        if self.detector_pointing is None:
            raise RuntimeError("The detector_pointing trait must be set")
        if self.single_precision:
            raise NotImplementedError("single_precision pixels are not implemented")
        self._geometry()
        if getattr(self, "_local_submaps", None) is None and self.create_dist is not None:
            self._local_submaps = np.zeros(self._n_submap, dtype=np.uint8)
        view = self.view if self.view is not None else self.detector_pointing.view
        self.detector_pointing.apply(data, detectors=detectors, use_accel=use_accel)
        dp = self.detector_pointing
        for ob in data.obs:
            dets = ob.select_local_detectors(detectors, flagmask=dp.det_mask)
            if len(dets) == 0:
                continue
            alld = ob.select_local_detectors(None, flagmask=dp.det_mask)
            exists = ob.detdata.ensure(self.pixels, sample_shape=(), dtype=np.int64,
                                       detectors=alld, accel=use_accel)
            done = self.__dict__.setdefault("_done", {})
            hit = self._local_submaps if self.create_dist is not None else \
                np.zeros(self._n_submap, dtype=np.uint8)
            if exists and done.get((id(ob), tuple(dets))):
                if self.create_dist is not None:  # pixels_healpix.py:218-236
                    for d in dets:
                        for iv in _view_intervals(ob, view):
                            p = ob.detdata[self.pixels][d, iv["first"]:iv["last"]]
                            hit[p[p >= 0] // self._n_pix_submap] = 1
                continue
            flags = ob.shared[dp.shared_flags] if dp.shared_flags is not None else \
                np.zeros(1, dtype=np.uint8)
            tmp = np.zeros(self._n_submap, dtype=np.uint8)
            K.pixels_healpix(ob.detdata[dp.quats].indices(dets), ob.detdata[dp.quats].data, flags,
                             dp.shared_flag_mask, ob.detdata[self.pixels].indices(dets),
                             ob.detdata[self.pixels].data, _view_intervals(ob, view), tmp,
                             self._n_pix_submap, self.nside, self.nest, use_accel)
            hit[:] |= tmp
            done[(id(ob), tuple(dets))] = True

    def _finalize(self, data, use_accel=False, **kwargs):
        if self.create_dist is not None:
            submaps = np.arange(self._n_submap, dtype=np.int64)[self._local_submaps == 1]
            data[self.create_dist] = PixelDistribution(
                n_pix=self._n_pix, n_submap=self._n_submap, local_submaps=submaps,
                comm=data.comm.comm_world)
            data[self.create_dist].nest = bool(self.nest)

    def _requires(self):
        return self.detector_pointing.requires()

    def _provides(self):
        prov = {"detdata": [self.pixels, self.detector_pointing.quats]}
        if self.create_dist is not None:
            prov["global"] = [self.create_dist]
        return prov


class StokesWeights(Operator):
    _defaults = dict(detector_pointing=None, mode="I", view=None, hwp_angle=None,
                     weights="weights", single_precision=False, cal=None, IAU=False,
                     focalplane_key="focalplane")

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        if self.detector_pointing is None:
            raise RuntimeError("The detector_pointing trait must be set")
        if self.mode not in ("I", "IQU"):
            raise NotImplementedError(f"mode '{self.mode}' is not implemented")
        dp = self.detector_pointing
        view = self.view if self.view is not None else dp.view
        nnz = 1 if self.mode == "I" else 3
        if self.mode == "IQU":
            dp.apply(data, detectors=detectors, use_accel=use_accel)
        for ob in data.obs:
            dets = ob.select_local_detectors(detectors, flagmask=dp.det_mask)
            if len(dets) == 0:
                continue
            alld = ob.select_local_detectors(None, flagmask=dp.det_mask)
            shape = () if nnz == 1 else (nnz,)
            exists = ob.detdata.ensure(self.weights, sample_shape=shape, dtype=np.float64,
                                       detectors=alld, accel=use_accel)
            done = self.__dict__.setdefault("_done", {})
            if exists and done.get((id(ob), tuple(dets))):
                continue  # stokes_weights.py:210-218
            fp = ob[self.focalplane_key]
            cal = np.array([fp[d]["cal"] if self.cal is None else ob[self.cal][d] for d in dets],
                           dtype=np.float64)
            widx = ob.detdata[self.weights].indices(dets)
            iv = _view_intervals(ob, view)
            if self.mode == "I":
                K.stokes_weights_I(widx, ob.detdata[self.weights].data, iv, cal, use_accel)
            else:
                eps = np.array([fp[d]["epsilon"] for d in dets], dtype=np.float64)
                gamma = np.array([fp[d]["gamma"] for d in dets], dtype=np.float64)
                hwp = ob.shared[self.hwp_angle] if self.hwp_angle is not None else np.zeros(1)
                K.stokes_weights_IQU(ob.detdata[dp.quats].indices(dets), ob.detdata[dp.quats].data,
                                     widx, ob.detdata[self.weights].data, hwp, iv, eps, gamma, cal,
                                     bool(self.IAU), use_accel)
            done[(id(ob), tuple(dets))] = True

    def _requires(self):
        req = self.detector_pointing.requires()
        if self.hwp_angle is not None:
            req.setdefault("shared", []).append(self.hwp_angle)
        return req

    def _provides(self):
        return {"detdata": [self.weights, self.detector_pointing.quats]}
