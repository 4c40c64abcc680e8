"""delayrepay.random entry points  (reference random.py:8-13: cast-wrapped cuRAND calls).

Values are produced ON THE DEVICE by a Philox4x32-10 counter generator (extras.philox), one
zero-read kernel per call: no host generation and no H2D copy, which is what makes 2^30-element
random inputs practical.  Bit-stream parity with NumPy's MT19937 is not achievable (different
generator); parity is statistical (tests check moments and ranges) -- SURVEY.md section 8f rank 2.
``choice`` keeps NumPy's host implementation (small index draws) and uploads the result.
"""
import numpy as _np

from . import extras
from .cuda import _to_device
from .delayarray import NPArray, cast


def seed(s=None):
    extras.seed(s)
    _np.random.seed(s)


def _shape(args):
    if len(args) == 1 and isinstance(args[0], (tuple, list)):
        return tuple(args[0])
    return tuple(int(a) for a in args)


def rand(*shape):
    return NPArray(extras.philox(_shape(shape), _np.float64, 0))


def randn(*shape):
    return NPArray(extras.philox(_shape(shape), _np.float64, 1))


def random(size=None):
    return NPArray(extras.philox(() if size is None else size, _np.float64, 0))


def randint(low, high=None, size=None, dtype=_np.int64):
    if high is None:
        low, high = 0, low
    if high <= low:
        raise ValueError("low >= high")
    return NPArray(extras.philox(() if size is None else size, dtype, 2, int(low), int(high) - int(low)))


choice = cast(lambda *a, **k: _to_device(_np.random.choice(*a, **k)))
