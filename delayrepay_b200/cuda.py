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
    """A backend array: a DeviceArray, or the row-sharded ShardView of sharding.py."""
    return isinstance(arr, DeviceArray) or getattr(arr, "_is_shard_view", False) is True


def run(ex):
    """Backend protocol entry (reference delayarray.py:43 -> cuda.py:91-96).  ``ex`` is normally
    one of this package's nodes; a node built by the REFERENCE's own front-end (this module
    dropped in as ``delayrepay/cuda.py`` under the unmodified reference ``delayarray.py``) is
    translated first, see ``import_foreign``."""
    from .delayarray import DelayArray
    if not isinstance(ex, DelayArray):
        return import_foreign(ex)._force()
    return engine.run(ex)


def import_foreign(ex, _memo=None):
    """Reference front-end node -> engine node, by the attributes the reference's own emitter
    reads (cuda.py:55-84): ``NPArray.array``, ``Scalar.val``, ``.func`` + ``.children`` of
    BinaryNumpyEx / UnaryFuncEx / BinaryFuncEx, ``NPRef.ref``.  A node the reference has already
    evaluated (``self.array`` cached, delayarray.py:38-44) enters as a leaf.  Unknown node
    classes raise NotImplementedError (the reference emitter returns NotImplemented for
    ReduceEx / NPRef, cuda.py:80-88)."""
    from . import delayarray as da
    memo = {} if _memo is None else _memo
    hit = memo.get(id(ex))
    if hit is not None:
        return hit
    name = type(ex).__name__
    if isinstance(ex, da.DelayArray):
        node = ex
    elif name == "Scalar":
        node = da.Scalar(ex.val)
    elif name == "NPRef":
        node = import_foreign(ex.ref, memo)
    elif "array" in getattr(ex, "__dict__", {}) and (is_ndarray(ex.array) or isinstance(ex.array, _np.ndarray)):
        node = da.NPArray(ex.array)
    elif name in ("BinaryNumpyEx", "UnaryFuncEx", "BinaryFuncEx"):
        node = da.create_ex(ex.func, [import_foreign(c, memo) for c in ex.children])
    else:
        raise NotImplementedError(f"cannot evaluate a foreign {name} node")
    memo[id(ex)] = node
    return node


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
    """Eager binary ufunc on backend arrays (what cupy.add etc. are to the reference's broadcast
    escape, delayarray.py:47-55,189-193): capture + force, returns a backend array."""
    def call(a, b, *args, **kwargs):
        from .delayarray import arg_to_numpy_ex, create_ex
        return create_ex(ufunc, [arg_to_numpy_ex(a), arg_to_numpy_ex(b)])._force()
    call.__name__ = ufunc.__name__
    return call


def _forced(res):
    from .delayarray import DelayArray
    if isinstance(res, DelayArray):
        return res._force()
    if isinstance(res, tuple):
        return tuple(_forced(r) for r in res)
    return res


def _eager_fn(np_func):
    """Eager array function on backend arrays (cupy.sum, cupy.roll ... in the reference,
    delayarray.py:511-568,608-615): runs this package's device handler and returns the backend
    array, so the unmodified reference handlers (`_backend.fallback.sum(arr.__array__(), ...)`)
    work on top of this module."""
    def call(*args, **kwargs):
        from . import delayarray as da
        lifted = [da.NPArray(a) if isinstance(a, DeviceArray) else a for a in args]
        return _forced(da.HANDLED_FUNCTIONS[np_func](*lifted, **kwargs))
    call.__name__ = np_func.__name__
    return call


def _dot(a, b, out=None):
    from . import delayarray as da
    left = da.arg_to_numpy_ex(a)
    return left._dot([left, da.arg_to_numpy_ex(b)])._force()


for _name in ("var", "sum", "transpose", "roll", "max", "min", "mean", "average", "repeat", "cumsum",
              "tile", "diag", "diagflat", "where", "prod", "std", "argmax", "argmin"):
    setattr(fallback, _name, _eager_fn(getattr(_np, _name)))
fallback.dot = _dot
fallback.matmul = _dot
for _name in ("maximum", "minimum", "greater", "less", "add", "multiply", "subtract", "true_divide"):
    setattr(fallback, _name, _eager(getattr(_np, _name)))
fallback.pi = _np.pi


def _rand_fn(name):
    def call(*args, **kwargs):
        from . import random as _random           # imports this module: resolved at call time
        return _forced(getattr(_random, name)(*args, **kwargs))
    call.__name__ = name
    return call


_random_ns = types.SimpleNamespace(**{n: _rand_fn(n) for n in
                                      ("rand", "randn", "random", "seed", "randint", "choice")})
fallback.random = _random_ns


def _fft_fn(name):
    def call(a, *args, **kwargs):
        from . import fft as _fft
        return getattr(_fft, name)(a, *args, **kwargs)._force()
    call.__name__ = name
    return call


# `fft.fft` on the backend MODULE itself: reference fft.py:7,12 calls backend.fft.fft(...)
fft = types.SimpleNamespace(fft=_fft_fn("fft"), ifft=_fft_fn("ifft"))
fallback.fft = fft

np = types.SimpleNamespace(
    pi=_np.pi, add=_eager(_np.add), multiply=_eager(_np.multiply), subtract=_eager(_np.subtract),
    true_divide=_eager(_np.true_divide), matmul=_dot, random=_random_ns, fft=fft,
)
