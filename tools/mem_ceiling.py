"""Memory-bound ceiling of a 3-input / 2-output float32 flat kernel (the Black-Scholes traffic
pattern, 20 B per element) with trivial arithmetic: what the load/store path can sustain."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
from delayrepay_b200._lib import lib, check
dr.set_device(0)
n = 1 << 30
S, K, T = (dr.NPArray(dr.DeviceArray.empty((n,), np.float32)) for _ in range(3))
for a in (S, K, T):
    (a * 0.0).run()


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


def light():
    dr.evaluate(S + K + T, S - K)


def one_out():
    (S + K + T).run()


def copy1():
    (S + 1.0).run()


for name, fn, bytes_per in (("3 in / 2 out, adds only", light, 20), ("3 in / 1 out", one_out, 16),
                            ("1 in / 1 out (copy-like)", copy1, 8)):
    ms = timed(fn)
    print(f"{name:28s} {ms:7.3f} ms  {n * bytes_per / ms / 1e6:7.0f} GB/s")
# the same pattern sustained (power-capped regime): 200 back-to-back launches
ms = timed(light, 200)
print(f"{'3 in / 2 out, 200 launches':28s} {ms:7.3f} ms  {n * 20 / ms / 1e6:7.0f} GB/s")
