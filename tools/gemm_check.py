"""Standalone check + timing of the tcgen05 3xTF32 GEMM (run under `timeout`)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
from delayrepay_b200._lib import lib, check
dr.set_device(0)
rng = np.random.default_rng(0)
for (m, k, n) in [(128, 64, 128), (512, 256, 384), (1000, 777, 650), (4096, 4096, 4096)]:
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((k, n)).astype(np.float32)
    A, B = dr.array(a), dr.array(b)
    got = (A @ B).get()
    want = a.astype(np.float64) @ b.astype(np.float64)
    ref32 = a @ b
    scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    err = np.abs(got - want).max() / scale.max()
    err_np = np.abs(ref32 - want).max() / scale.max()
    print(f"{m}x{k}x{n}: max err/scale ours {err:.2e}  numpy-f32 {err_np:.2e}  allclose(rtol1e-5,atol=1e-5*scale) "
          f"{bool(np.all(np.abs(got - want) <= 1e-5 * scale + 1e-6))}", flush=True)
m = k = n = 4096
ev = [C.c_uint64(), C.c_uint64()]
for e in ev: check(lib.drc_event_create(0, C.byref(e)))
for _ in range(2): (A @ B).run()
dr.synchronize()
check(lib.drc_event_record(0, 0, ev[0].value))
for _ in range(10): (A @ B).run()
check(lib.drc_event_record(0, 0, ev[1].value)); check(lib.drc_event_sync(0, ev[1].value))
ms = C.c_float(); check(lib.drc_event_elapsed_ms(0, ev[0].value, ev[1].value, C.byref(ms)))
per = ms.value / 10
print(f"4096^3 incl. split pre-pass: {per:.3f} ms  -> {2*m*n*k/per/1e9:.1f} TFLOP/s fp32-equivalent")
