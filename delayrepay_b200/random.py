"""delayrepay.random entry points  (reference random.py:8-13: cast-wrapped backend RNG).

Round 1: values come from NumPy's host generator (bit-identical streams to the reference's
CPU path for a given seed) and are uploaded once; the device Philox producer is the next row
(SURVEY.md section 8f rank 2).
"""
import numpy as _np

from .delayarray import cast
from .cuda import _to_device

seed = _np.random.seed
rand = cast(lambda *a, **k: _to_device(_np.random.rand(*a, **k)))
randn = cast(lambda *a, **k: _to_device(_np.random.randn(*a, **k)))
random = cast(lambda *a, **k: _to_device(_np.random.random(*a, **k)))
randint = cast(lambda *a, **k: _to_device(_np.random.randint(*a, **k)))
choice = cast(lambda *a, **k: _to_device(_np.random.choice(*a, **k)))
