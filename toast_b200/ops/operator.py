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


def _ident(obj):
    arr = getattr(obj, "arr", None)
    return id(arr) if arr is not None else id(obj)


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
        self._req, self._prov, self._pre = dict(), dict(), set()
        if use_accel:
            for op in ops:
                _merge(self._req, op.requires())
                _merge(self._prov, op.provides())
            # buffers somebody else already put on the device stay there afterwards
            self._pre = {_ident(o) for o, _ in self._device_objects(data, self._req, self._prov)
                         if o.accel_exists()}
            self._stage(data)
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

    def _finalize(self, data, use_accel=False, **kwargs):
        result = None
        for op in self.operators or []:
            result = op.finalize(data, use_accel=use_accel)
        if use_accel:
            # pipeline.py:266-303: copy provides() back, drop what this pipeline staged/created
            prov_ids = {_ident(o) for o, _ in self._device_objects(data, self._prov, {})}
            for obj, nm in self._device_objects(data, self._req, self._prov):
                if not obj.accel_exists():
                    continue
                if _ident(obj) in prov_ids:
                    obj.accel_update_host(nm)
                if _ident(obj) not in self._pre:
                    obj.accel_delete(nm)
        return result

    def _stage(self, data):
        for obj, nm in self._device_objects(data, self._req, {}):
            if not obj.accel_exists():
                obj.accel_create(nm)
                obj.accel_update_device(nm)

    @staticmethod
    def _device_objects(data, req, prov):
        """(object, name) for every existing buffer named by the requires / provides dicts."""
        from .. import _libtoast as K

        class _Buf:
            def __init__(self, arr):
                self.arr = arr

            def accel_exists(self):
                return K.accel_present(self.arr, "shared")

            def accel_create(self, nm):
                K.accel_create(self.arr, nm)

            def accel_update_device(self, nm):
                K.accel_update_device(self.arr, nm)

            def accel_update_host(self, nm):
                K.accel_update_host(self.arr, nm)

            def accel_delete(self, nm):
                K.accel_delete(self.arr, nm)

        keys = dict()
        _merge(keys, req)
        _merge(keys, prov)
        out, seen = [], set()

        def add(obj, nm):
            ident = _ident(obj)
            if ident not in seen:
                seen.add(ident)
                out.append((obj, nm))

        for ob in data.obs:
            for key in keys.get("detdata", []):
                if key is not None and key in ob.detdata:
                    add(ob.detdata[key], key)
            for key in keys.get("shared", []):
                if key is not None and key in ob.shared:
                    add(_Buf(ob.shared[key]), key)
        for key in keys.get("global", []):
            if key is not None and key in data and hasattr(data[key], "accel_create"):
                add(data[key], key)
        return out
