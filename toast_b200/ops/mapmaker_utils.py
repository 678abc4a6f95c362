"""NoiseWeight, ScanMap, BuildNoiseWeighted, BuildHitMap, BuildInverseCovariance,
CovarianceAndHits and BinMap (``ops/noise_weight/noise_weight.py:76-136``,
``ops/scan_map/scan_map.py:93-181``, ``ops/mapmaker_utils/mapmaker_utils.py:114-206, 352-515,
559-960, 1131-1270``, ``ops/mapmaker_binning.py:179-294``) on the CUDA kernels."""

import numpy as np

from .. import _libtoast as K
from .. import kernels as KC
from ..covariance import covariance_apply, covariance_invert
from ..pixels import PixelData
from .operator import Operator, Pipeline


def _dets(ob, detectors, det_mask):
    return ob.select_local_detectors(detectors, flagmask=det_mask)


class NoiseWeight(Operator):
    _defaults = dict(noise_model="noise_model", view=None, det_data="signal", det_mask=1,
                     det_flag_mask=1)

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        for ob in data.obs:
            dets = _dets(ob, detectors, self.det_mask)
            if len(dets) == 0:
                continue
            if self.noise_model not in ob:
                raise RuntimeError(f"Noise model {self.noise_model} not in observation {ob.name}")
            noise = ob[self.noise_model]
            w = np.array([noise.detector_weight(d) for d in dets], dtype=np.float64)
            K.noise_weight(ob.detdata[self.det_data].data, ob.detdata[self.det_data].indices(dets),
                           ob.intervals[self.view], w, use_accel)

    def _requires(self):
        return {"detdata": [self.det_data]}

    def _provides(self):
        return {"detdata": [self.det_data]}


class ScanMap(Operator):
    _defaults = dict(det_data="signal", det_mask=1, det_flag_mask=1, view=None, pixels="pixels",
                     weights=None, map_key=None, subtract=False, zero=False)

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        if self.map_key not in data:
            raise RuntimeError(f"The map_key '{self.map_key}' does not exist in the data")
        m = data[self.map_key]
        dist = m.distribution
        fn = getattr(K, f"ops_scan_map_{m.dtype.name}")
        for ob in data.obs:
            dets = _dets(ob, detectors, self.det_mask)
            if len(dets) == 0:
                continue
            ob.detdata.ensure(self.det_data, detectors=ob.local_detectors, accel=use_accel)
            if self.weights is None:
                raise RuntimeError("ScanMap without weights is not implemented")
            fn(dist.global_submap_to_local, dist.n_pix_submap, m.data,
               ob.detdata[self.det_data].data, ob.detdata[self.det_data].indices(dets),
               ob.detdata[self.pixels].data, ob.detdata[self.pixels].indices(dets),
               ob.detdata[self.weights].data, ob.detdata[self.weights].indices(dets),
               ob.intervals[self.view], 1.0, bool(self.zero), bool(self.subtract), False,
               use_accel)

    def _requires(self):
        return {"global": [self.map_key], "detdata": [self.pixels, self.weights, self.det_data]}

    def _provides(self):
        return {"detdata": [self.det_data]}


class ScanMask(Operator):
    """ops/scan_map/scan_map.py:216-357: raise `det_flags_value` in the detector flags wherever
    the sample's pixel has any of `mask_bits` set in the integer mask map.  The reference does
    this on the host with numpy; here the mask is gathered with the scan_map kernel."""

    _defaults = dict(det_mask=1, det_flags="flags", det_flags_value=1, det_flag_mask=1,
                     view=None, pixels="pixels", mask_key=None, mask_bits=255)

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        if self.mask_key not in data:
            raise RuntimeError(f"The mask_key '{self.mask_key}' does not exist in the data")
        mask = data[self.mask_key]
        dist = mask.distribution
        if mask.n_value != 1:
            raise RuntimeError("The mask map must have one value per pixel")
        hit = PixelData(dist, np.float64, n_value=1)
        hit.data[:] = (mask.data.astype(np.int64) & int(self.mask_bits)) != 0
        for ob in data.obs:
            dets = _dets(ob, detectors, self.det_mask)
            if len(dets) == 0:
                continue
            ob.detdata.ensure(self.det_flags, dtype=np.uint8, detectors=ob.local_detectors)
            pix = ob.detdata[self.pixels]
            tmp = np.zeros((len(dets), ob.n_local_samples))
            ones = np.ones((len(dets), ob.n_local_samples))
            didx = np.arange(len(dets), dtype=np.int32)
            K.ops_scan_map_float64(dist.global_submap_to_local, dist.n_pix_submap, hit.data, tmp,
                                   didx, pix.data, pix.indices(dets), ones, didx,
                                   ob.intervals[self.view], 1.0, True, False, False, False)
            fl = ob.detdata[self.det_flags]
            rows = fl.indices(dets)
            fl.data[rows] |= np.where(tmp != 0, np.uint8(self.det_flags_value), np.uint8(0))

    def _requires(self):
        return {"global": [self.mask_key], "detdata": [self.pixels]}

    def _provides(self):
        return {"detdata": [self.det_flags]}


class BuildNoiseWeighted(Operator):
    _defaults = dict(pixel_dist=None, zmap=None, view=None, det_data="signal", det_mask=1,
                     det_flags="flags", det_flag_mask=1, shared_flags="flags",
                     shared_flag_mask=1, pixels="pixels", weights="weights",
                     noise_model="noise_model", sync_type="alltoallv")

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        if self.pixel_dist not in data:
            raise RuntimeError(f"Pixel distribution '{self.pixel_dist}' does not exist")
        dist = data[self.pixel_dist]
        for ob in data.obs:
            dets = _dets(ob, detectors, self.det_mask)
            if len(dets) == 0:
                continue
            wts = ob.detdata[self.weights]
            nnz = 1 if len(wts.detector_shape) == 1 else wts.detector_shape[1]
            if self.zmap not in data:
                data[self.zmap] = PixelData(dist, np.float64, n_value=nnz)
                if use_accel:
                    data[self.zmap].accel_create(self.zmap)
                    data[self.zmap].accel_update_device(self.zmap)
            zmap = data[self.zmap]
            noise = ob[self.noise_model]
            scale = np.array([noise.detector_weight(d) for d in dets], dtype=np.float64)
            if self.det_flags is not None:
                fidx = ob.detdata[self.det_flags].indices(dets)
                fdata = ob.detdata[self.det_flags].data
            else:  # mapmaker_utils.py:836-838
                fidx = np.array([-1], dtype=np.int32)
                fdata = np.zeros((1, 1), dtype=np.uint8)
            sflags = ob.shared[self.shared_flags] if self.shared_flags is not None else \
                np.zeros(1, dtype=np.uint8)
            K.build_noise_weighted(
                dist.global_submap_to_local, zmap.data, ob.detdata[self.pixels].indices(dets),
                ob.detdata[self.pixels].data, wts.indices(dets), wts.data,
                ob.detdata[self.det_data].indices(dets), ob.detdata[self.det_data].data, fidx,
                fdata, scale, self.det_flag_mask, ob.intervals[self.view], sflags,
                self.shared_flag_mask, use_accel)

    def _finalize(self, data, use_accel=False, **kwargs):
        if self.zmap in data:
            z = data[self.zmap]
            # mapmaker_utils.py:885-925: host bounce around the collective
            if use_accel and z.accel_exists():
                z.accel_update_host(self.zmap)
            if self.sync_type == "alltoallv":
                z.sync_alltoallv()
            else:
                z.sync_allreduce()
            if use_accel and z.accel_exists():
                z.accel_update_device(self.zmap)

    def _requires(self):
        req = {"global": [self.pixel_dist], "detdata": [self.pixels, self.weights, self.det_data],
               "shared": []}
        if self.det_flags is not None:
            req["detdata"].append(self.det_flags)
        if self.shared_flags is not None:
            req["shared"].append(self.shared_flags)
        return req

    def _provides(self):
        return {"global": [self.zmap]}


class _CovAccum(Operator):
    """Shared body of BuildHitMap / BuildInverseCovariance: one fused accumulation kernel
    instead of the reference's per-detector host loop over global_pixel_to_submap."""

    _defaults = dict(pixel_dist=None, view=None, det_mask=1, det_flags="flags", det_flag_mask=1,
                     shared_flags="flags", shared_flag_mask=1, pixels="pixels",
                     weights="weights", noise_model="noise_model", sync_type="alltoallv")

    def _accum(self, data, detectors, hits, invcov):
        dist = data[self.pixel_dist]
        for ob in data.obs:
            dets = _dets(ob, detectors, self.det_mask)
            if len(dets) == 0:
                continue
            wts = ob.detdata[self.weights] if invcov is not None else None
            nnz = 3
            if wts is not None:
                nnz = 1 if len(wts.detector_shape) == 1 else wts.detector_shape[1]
            scale = None
            if invcov is not None:
                noise = ob[self.noise_model]
                scale = np.array([noise.detector_weight(d) for d in dets], dtype=np.float64)
            fdata = ob.detdata[self.det_flags].data if self.det_flags is not None else None
            fidx = ob.detdata[self.det_flags].indices(dets) if self.det_flags is not None else None
            sflags = ob.shared[self.shared_flags] if self.shared_flags is not None else None
            KC.cov_accum(dist.global_submap_to_local, dist.n_local_submap, dist.n_pix_submap, nnz,
                         hits.raw if hits is not None else None,
                         invcov.raw if invcov is not None else None,
                         ob.detdata[self.pixels].indices(dets), ob.detdata[self.pixels].data,
                         wts.indices(dets) if wts is not None else None,
                         wts.data if wts is not None else None, fidx, fdata, scale,
                         self.det_flag_mask, ob.intervals[self.view], sflags,
                         self.shared_flag_mask)


class BuildHitMap(_CovAccum):
    _defaults = dict(hits=None)

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        dist = data[self.pixel_dist]
        if self.hits not in data:
            data[self.hits] = PixelData(dist, np.int64, n_value=1)
        self._accum(data, detectors, data[self.hits], None)

    def _finalize(self, data, **kwargs):
        if self.hits in data:
            data[self.hits].sync_allreduce()


class BuildInverseCovariance(_CovAccum):
    _defaults = dict(inverse_covariance=None)

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        dist = data[self.pixel_dist]
        if self.inverse_covariance not in data:
            nnz = None
            for ob in data.obs:
                shp = ob.detdata[self.weights].detector_shape
                nnz = 1 if len(shp) == 1 else shp[1]
                break
            data[self.inverse_covariance] = PixelData(dist, np.float64,
                                                      n_value=nnz * (nnz + 1) // 2)
        self._accum(data, detectors, None, data[self.inverse_covariance])

    def _finalize(self, data, **kwargs):
        if self.inverse_covariance in data:
            data[self.inverse_covariance].sync_allreduce()


class CovarianceAndHits(_CovAccum):
    """mapmaker_utils.py:1131-1270: hits + inverse covariance + inverted covariance + rcond."""

    _defaults = dict(hits=None, inverse_covariance=None, covariance=None, rcond=None,
                     rcond_threshold=1.0e-8)

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        dist = data[self.pixel_dist]
        nnz = None
        for ob in data.obs:
            shp = ob.detdata[self.weights].detector_shape
            nnz = 1 if len(shp) == 1 else shp[1]
            break
        for key in (self.hits, self.inverse_covariance, self.covariance, self.rcond):
            if key is not None and key in data:
                del data[key]
        data[self.hits] = PixelData(dist, np.int64, n_value=1)
        data[self.inverse_covariance] = PixelData(dist, np.float64, n_value=nnz * (nnz + 1) // 2)
        self._accum(data, detectors, data[self.hits], data[self.inverse_covariance])

    def _finalize(self, data, **kwargs):
        data[self.hits].sync_allreduce()
        inv = data[self.inverse_covariance]
        inv.sync_allreduce()
        cov = PixelData(inv.distribution, np.float64, n_value=inv.n_value)
        cov.data[:] = inv.data
        rc = PixelData(inv.distribution, np.float64, n_value=1)
        covariance_invert(cov, self.rcond_threshold, rcond=rc)
        data[self.covariance] = cov
        if self.rcond is not None:
            data[self.rcond] = rc


class BinMap(Operator):
    """mapmaker_binning.py:27-294: noise-weighted map of a timestream times the pixel
    covariance.  ``full_pointing=False`` expands pointing one detector at a time."""

    _defaults = dict(pixel_dist=None, covariance=None, binned="binned", det_data="signal",
                     det_mask=1, det_flags="flags", det_flag_mask=1, shared_flags="flags",
                     shared_flag_mask=1, pixel_pointing=None, stokes_weights=None,
                     pre_process=None, noise_model="noise_model", sync_type="alltoallv",
                     full_pointing=False, noiseweighted=None)

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        for trait in ("pixel_dist", "covariance", "pixel_pointing", "stokes_weights"):
            if getattr(self, trait) is None:
                raise RuntimeError(f"You must set the '{trait}' trait before calling exec()")
        if self.pixel_dist not in data:
            raise RuntimeError(f"Pixel distribution '{self.pixel_dist}' does not exist")
        if self.covariance not in data:
            raise RuntimeError(f"Pixel covariance '{self.covariance}' does not exist")
        cov = data[self.covariance]
        zkey = self.noiseweighted if self.noiseweighted is not None else f"{self.name}_zmap"
        if zkey in data:
            del data[zkey]
        if self.binned in data:
            del data[self.binned]
        build = BuildNoiseWeighted(
            pixel_dist=self.pixel_dist, zmap=zkey, view=self.pixel_pointing.view,
            pixels=self.pixel_pointing.pixels, weights=self.stokes_weights.weights,
            noise_model=self.noise_model, det_data=self.det_data, det_mask=self.det_mask,
            det_flags=self.det_flags, det_flag_mask=self.det_flag_mask,
            shared_flags=self.shared_flags, shared_flag_mask=self.shared_flag_mask,
            sync_type=self.sync_type)
        ops = []
        if self.pre_process is not None:
            ops.append(self.pre_process)
        ops += [self.pixel_pointing, self.stokes_weights, build]
        pipe = Pipeline(operators=ops,
                        detector_sets=["ALL"] if self.full_pointing else ["SINGLE"])
        pipe.apply(data, detectors=detectors, use_accel=use_accel)
        z = data[zkey]
        covariance_apply(cov, z)
        data[self.binned] = z
        if self.noiseweighted is None:
            del data[zkey]

    def _requires(self):
        return {"global": [self.pixel_dist, self.covariance], "detdata": [self.det_data]}

    def _provides(self):
        return {"global": [self.binned]}
