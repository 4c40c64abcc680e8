"""The CUDA backend module -- the drop-in boundary of the reference's backend protocol
(reference cuda.py:17-32,91-96):

    run(ex) -> array      evaluate a graph node, return the backend (device) array
    is_ndarray(obj)       is obj a backend array?
    np                    namespace with matmul/add/multiply/subtract/true_divide, pi, random
    fallback              namespace with the creation functions and eager helpers
    fft                   fft.fft entry point

``np``/``fallback`` are CuPy in the reference; here they are small namespaces over
DeviceArray + the engine.  Without a GPU the creation functions hand back host arrays so that
graphs can still be *built, planned and compiled* (tests do that); evaluating one raises.
"""
import types

import numpy as _np

from . import _lib, engine
from .device import DeviceArray


def is_ndarray(arr):
    return isinstance(arr, DeviceArray)


def run(ex):
    return engine.run(ex)


def run_many(nodes):
    return engine.run_many(nodes)


def _to_device(host):
    host = _np.asarray(host)
    if _lib.gpu_available() or engine.is_dry():
        return DeviceArray.from_host(host)
    return host                      # graph building only; evaluation needs a device


def _filled(shape, value, dtype):
    if not (_lib.gpu_available() or engine.is_dry()):
        return _np.full(shape, value, dtype=dtype)
    out = DeviceArray.empty(shape if not isinstance(shape, (int, _np.integer)) else (int(shape),),
                            dtype)
    if out.size:
        out.fill(value)
    return out


def _like_shape(a):
    return tuple(a.shape)


def _host_ctor(fn):
    def make(*args, **kwargs):
        return _to_device(fn(*args, **kwargs))
    make.__name__ = fn.__name__
    return make


def _array(obj, dtype=None, copy=True, **kw):
    from .delayarray import DelayArray
    if isinstance(obj, DelayArray):
        dev = obj._force()
        return dev.astype(dtype) if dtype is not None and _np.dtype(dtype) != dev.dtype else (
            dev.copy() if copy else dev)
    if isinstance(obj, DeviceArray):
        return obj.astype(dtype) if dtype is not None else (obj.copy() if copy else obj)
    if isinstance(obj, _np.ndarray) and (dtype is None or _np.dtype(dtype) == obj.dtype):
        return _to_device(obj)           # the upload itself is the copy
    return _to_device(_np.array(obj, dtype=dtype))


def _asarray(obj, dtype=None, **kw):
    return _array(obj, dtype=dtype, copy=False)


def _empty(shape, dtype=float, **kw):
    if not (_lib.gpu_available() or engine.is_dry()):
        return _np.empty(shape, dtype=dtype)
    return DeviceArray.empty(shape if not isinstance(shape, (int, _np.integer)) else (int(shape),),
                             dtype)


fallback = types.SimpleNamespace(
    newaxis=None,
    empty=_empty,
    empty_like=lambda a, dtype=None, **k: _empty(_like_shape(a), dtype or a.dtype),
    ones=lambda shape, dtype=float, **k: _filled(shape, 1, dtype),
    ones_like=lambda a, dtype=None, **k: _filled(_like_shape(a), 1, dtype or a.dtype),
    zeros=lambda shape, dtype=float, **k: _filled(shape, 0, dtype),
    zeros_like=lambda a, dtype=None, **k: _filled(_like_shape(a), 0, dtype or a.dtype),
    full=lambda shape, fill_value, dtype=None, **k: _filled(
        shape, fill_value, dtype if dtype is not None else _np.asarray(fill_value).dtype),
    full_like=lambda a, fill_value, dtype=None, **k: _filled(
        _like_shape(a), fill_value, dtype or a.dtype),
    array=_array, asarray=_asarray, asanyarray=_asarray, ascontiguousarray=_asarray,
    copy=lambda a, **k: _array(a, copy=True),
    eye=_host_ctor(_np.eye), identity=_host_ctor(_np.identity), arange=_host_ctor(_np.arange),
    linspace=_host_ctor(_np.linspace), logspace=_host_ctor(_np.logspace),
    tri=_host_ctor(_np.tri),
    tril=lambda a, k=0: _to_device(_np.tril(_np.asarray(a), k)),
    triu=lambda a, k=0: _to_device(_np.triu(_np.asarray(a), k)),
)


def _eager(ufunc):
    def call(a, b):
        from .delayarray import arg_to_numpy_ex, create_ex
        return create_ex(ufunc, [arg_to_numpy_ex(a), arg_to_numpy_ex(b)])._force()
    return call


np = types.SimpleNamespace(
    pi=_np.pi, add=_eager(_np.add), multiply=_eager(_np.multiply), subtract=_eager(_np.subtract),
    true_divide=_eager(_np.true_divide), matmul=lambda a, b: (
        __import__("delayrepay_b200").delayarray.arg_to_numpy_ex(a) @ b)._force(),
)
