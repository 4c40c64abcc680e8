"""Drop-in alias: ``import delayrepay as np`` resolves to the B200 engine (delayrepay_b200),
so scripts and the reference's own tests/test.py run unchanged (reference __init__.py:4-17)."""
import sys as _sys

import delayrepay_b200 as _impl
from delayrepay_b200 import *               # noqa: F401,F403
from delayrepay_b200 import (NPArray, array, full, ones, sum, random, fft, pi,  # noqa: F401
                             backend, delayarray)

_sys.modules[__name__ + ".delayarray"] = _impl.delayarray
_sys.modules[__name__ + ".backend"] = _impl.backend
_sys.modules[__name__ + ".random"] = _impl.random
_sys.modules[__name__ + ".fft"] = _impl.fft


def __getattr__(name):
    return getattr(_impl, name)         # NumPy pass-through for everything else (see _impl)
