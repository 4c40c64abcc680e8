"""delayrepay_b200 -- a B200-native lazy-evaluation array engine behind DelayRepay's drop-in
NumPy API (reference delayrepay/__init__.py:4-17).  ``import delayrepay_b200 as np`` (or the
alias package ``import delayrepay as np``).
"""
from . import backend                       # noqa: F401
from .delayarray import *                   # noqa: F401,F403
from .delayarray import (DelayArray, NPArray, Scalar, BinaryNumpyEx, UnaryFuncEx,  # noqa: F401
                         BinaryFuncEx, ReduceEx, DotEx, MVEx, MMEx, NPRef, Memoiser, reset, cast,
                         evaluate, implements, HANDLED_FUNCTIONS, sum, max, min, abs)
from .device import (DeviceArray, set_device, current_device, synchronize,  # noqa: F401
                     pinned_empty)
from . import random, fft                   # noqa: F401
from .stream import map_chunks              # noqa: F401
from . import sharding                      # noqa: F401
from .sharding import shard                 # noqa: F401

pi = backend.backend.np.pi


def __getattr__(name):
    """Names this module does not define resolve to NumPy's (constants, dtypes, every ufunc and
    array function): called with a DelayArray they dispatch into the engine through
    __array_ufunc__ / __array_function__ (KeyError when there is no device implementation --
    nothing evaluates device data on the host); called with host data they are NumPy."""
    import numpy as _np
    if name.startswith("__"):
        raise AttributeError(name)
    try:
        return getattr(_np, name)
    except AttributeError:
        raise AttributeError(f"module 'delayrepay' has no attribute {name!r}") from None
__version__ = "0.1.0"
