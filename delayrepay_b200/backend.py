"""Backend selector  (reference backend.py:1-11).

The reference picks cpu.py / rise.py / cuda.py from DELAY_CPU / DELAY_LIFT at import time.
This package has exactly one backend -- the B200 CUDA engine -- and no CPU fallback:
the reference's CPU path lives on only as the parity oracle under oracle/ (test code).
"""
import os

if "DELAY_CPU" in os.environ or "DELAY_LIFT" in os.environ:
    import warnings
    warnings.warn("delayrepay_b200 has a single backend (CUDA, sm_100a); DELAY_CPU / DELAY_LIFT "
                  "are ignored -- there is no CPU fallback and no RISE emitter", stacklevel=2)

from . import cuda as be  # noqa: E402

backend = be
