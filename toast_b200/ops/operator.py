"""Operator base class and Pipeline (``ops/operator.py:29-258``, ``ops/pipeline.py:18-389``).

Traits become plain keyword attributes with the reference's trait names.  ``apply`` =
``exec`` + ``finalize``; ``use_accel=True`` makes the kernels look their large buffers up in the
device table, which ``Pipeline`` fills from ``requires()`` and drains from ``provides()`` exactly
where the reference stages data (``pipeline.py:208-303``).
"""


class Operator:
    _defaults = {}

    def __init__(self, name=None, **kwargs):
        self.name = name if name is not None else type(self).__name__
        merged = {}
        for klass in reversed(type(self).__mro__):
            merged.update(getattr(klass, "_defaults", {}))
        for k, v in merged.items():
            setattr(self, k, v)
        for k, v in kwargs.items():
            if k not in merged:
                raise AttributeError(f"{type(self).__name__} has no trait '{k}'")
            setattr(self, k, v)

    # -- reference interface ------------------------------------------------------------------
    def exec(self, data, detectors=None, use_accel=None, **kwargs):
        return self._exec(data, detectors=detectors, use_accel=bool(use_accel), **kwargs)

    def finalize(self, data, use_accel=None, **kwargs):
        return self._finalize(data, use_accel=bool(use_accel), **kwargs)

    def apply(self, data, detectors=None, use_accel=None, **kwargs):
        self.exec(data, detectors=detectors, use_accel=use_accel, **kwargs)
        return self.finalize(data, use_accel=use_accel, **kwargs)

    def requires(self):
        return self._requires()

    def provides(self):
        return self._provides()

    def supports_accel(self):
        return True

    def _finalize(self, data, **kwargs):
        return None

    def _requires(self):
        return dict()

    def _provides(self):
        return dict()

    def duplicate(self):
        import copy

        return copy.copy(self)


def _merge(a, b):
    for k, v in b.items():
        a.setdefault(k, [])
        for x in v:
            if x not in a[k]:
                a[k].append(x)
    return a


class Pipeline(Operator):
    """Run operators over detector sets, staging their buffers to the device once."""

    _defaults = dict(operators=None, detector_sets=("ALL",))

    def _exec(self, data, detectors=None, use_accel=False, **kwargs):
        ops = list(self.operators or [])
        staged = []
        if use_accel:
            req = dict()
            for op in ops:
                _merge(req, op.requires())
                _merge(req, op.provides())
            staged = self._stage(data, req)
        try:
            for dset in self.detector_sets:
                if dset == "ALL":
                    for op in ops:
                        op.exec(data, detectors=detectors, use_accel=use_accel)
                elif dset == "SINGLE":
                    for det in data.all_local_detectors(selection=detectors):
                        for op in ops:
                            op.exec(data, detectors=[det], use_accel=use_accel)
                else:
                    for op in ops:
                        op.exec(data, detectors=dset, use_accel=use_accel)
        finally:
            self._staged = staged

    def _finalize(self, data, use_accel=False, **kwargs):
        result = None
        for op in self.operators or []:
            result = op.finalize(data, use_accel=use_accel)
        for obj, nm in getattr(self, "_staged", []):
            obj.accel_update_host(nm)
            obj.accel_delete(nm)
        self._staged = []
        return result

    @staticmethod
    def _stage(data, req):
        from .. import _libtoast as K

        class _Buf:
            def __init__(self, arr):
                self.arr = arr

            def accel_update_host(self, nm):
                K.accel_update_host(self.arr, nm)

            def accel_delete(self, nm):
                K.accel_delete(self.arr, nm)

        staged = []
        for ob in data.obs:
            for key in req.get("detdata", []):
                if key in ob.detdata and not ob.detdata[key].accel_exists():
                    ob.detdata[key].accel_create(key)
                    ob.detdata[key].accel_update_device(key)
                    staged.append((ob.detdata[key], key))
            for key in req.get("shared", []):
                if key in ob.shared and not K.accel_present(ob.shared[key], key):
                    K.accel_create(ob.shared[key], key)
                    K.accel_update_device(ob.shared[key], key)
                    staged.append((_Buf(ob.shared[key]), key))
        for key in req.get("global", []):
            if key in data and hasattr(data[key], "accel_create") and not data[key].accel_exists():
                data[key].accel_create(key)
                data[key].accel_update_device(key)
                staged.append((data[key], key))
        return staged
