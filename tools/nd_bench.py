"""Row/column-broadcast normalisation of a 16384 x 16384 float32 matrix (nd family): GB/s with
and without the inner-dimension vectors (DR_NO_NDVEC=1)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
from delayrepay_b200._lib import lib, check
dr.set_device(0)
n = 16384
rng = np.random.default_rng(0)
X = dr.tile(dr.array(rng.standard_normal((1024, 1024)).astype(np.float32)), (n // 1024, n // 1024))
mu = dr.array(rng.standard_normal(n).astype(np.float32))
sd = dr.array(rng.uniform(0.5, 2.0, n).astype(np.float32))


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = C.c_uint64(), C.c_uint64()
    check(lib.drc_event_create(0, C.byref(a))); check(lib.drc_event_create(0, C.byref(b)))
    dr.synchronize()
    check(lib.drc_event_record(0, 0, a.value))
    for _ in range(reps):
        fn()
    check(lib.drc_event_record(0, 0, b.value)); check(lib.drc_event_sync(0, b.value))
    ms = C.c_float(); check(lib.drc_event_elapsed_ms(0, a.value, b.value, C.byref(ms)))
    return ms.value / reps


ms = timed(lambda: ((X - mu[None, :]) / sd[:, None]).run())
print(f"(X - mu[None,:]) / sd[:,None]  {n}x{n} f32: {ms:.3f} ms  {n * n * 8 / ms / 1e6:.0f} GB/s", os.environ.get("DR_NO_NDVEC", ""))
ms = timed(lambda: (X[1:-1, 4:-4] * 2.0 + X[2:, 4:-4]).run())
print(f"strided views                  {ms:.3f} ms  {(n - 2) * (n - 8) * 12 / ms / 1e6:.0f} GB/s")
