"""GB/s of the operations around the hot path (axis reductions, matvec, scans, gathers, copies)
on 16384 x 16384 matrices / 2^28 vectors: finds the families that are still far from the HBM
roofline.  Algorithmic bytes = operands read once + result written once."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
from delayrepay_b200._lib import lib, check
dr.set_device(0)
n = 16384
rng = np.random.default_rng(0)
only = sys.argv[1:] if len(sys.argv) > 1 else None


def timed(fn, reps=5):
    for _ in range(2):
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


def report(name, nbytes, fn, reps=5):
    if only and not any(o in name for o in only):
        return
    try:
        ms = timed(fn, reps)
        print(f"{name:44s} {ms:9.3f} ms  {nbytes / ms / 1e6:8.0f} GB/s", flush=True)
    except Exception as e:                                      # noqa: BLE001
        print(f"{name:44s} FAILED {type(e).__name__}: {str(e)[:120]}", flush=True)


for dt, w in ((np.float32, 4), (np.float64, 8)):
    t = dt.__name__
    X = dr.tile(dr.array(rng.standard_normal((1024, 1024)).astype(dt)), (n // 1024, n // 1024)).run()
    v = dr.array(rng.standard_normal(n).astype(dt))
    big = X.reshape(-1)
    N = n * n
    report(f"sum(X, axis=1) {t}", N * w, lambda: np.sum(X, axis=1).run())
    report(f"sum(X, axis=0) {t}", N * w, lambda: np.sum(X, axis=0).run())
    report(f"max(X, axis=0) {t}", N * w, lambda: np.max(X, axis=0).run())
    report(f"mean(X*X, axis=1) {t}", N * w, lambda: np.mean(X * X, axis=1).run())
    report(f"sum(X) {t}", N * w, lambda: np.sum(X).run())
    report(f"X @ v {t}", N * w, lambda: (X @ v).run())
    report(f"v @ X {t}", N * w, lambda: (v @ X).run())
    report(f"var(X, axis=0) {t}", 2 * N * w, lambda: np.var(X, axis=0).run())
    report(f"argmax(big) {t}", N * w, lambda: np.argmax(big).run())
    report(f"cumsum(big) {t}", 2 * N * w, lambda: np.cumsum(big).run())
    report(f"cumsum(X, axis=0) {t}", 2 * N * w, lambda: np.cumsum(X, axis=0).run())
    report(f"cumsum(X, axis=1) {t}", 2 * N * w, lambda: np.cumsum(X, axis=1).run())
    report(f"X.T.copy() {t}", 2 * N * w, lambda: X.T.copy())
    report(f"X.T.reshape(-1) {t}", 2 * N * w, lambda: X.T.reshape(-1))
    report(f"X.T + X {t}", 3 * N * w, lambda: (X.T + X).run())
    report(f"X[:, ::2] * 2 {t}", N * w, lambda: (X[:, ::2] * 2.0).run())
    report(f"concatenate([X, X]) {t}", 4 * N * w, lambda: np.concatenate([X, X]))
    report(f"roll(big, 12345) {t}", 2 * N * w, lambda: np.roll(big, 12345))
    report(f"where(X > 0, X, 0) {t}", 2 * N * w, lambda: np.where(X > 0, X, 0.0).run())
    idx = dr.array(rng.integers(0, N, N // 4))
    report(f"big[idx] random gather N/4 {t}", (N // 4) * (8 + 2 * w), lambda: big[idx])
    rows = dr.array(rng.integers(0, n, n))
    report(f"X[rows] row gather {t}", 2 * N * w, lambda: X[rows])
    report(f"big[big > 0] {t}", N * w + N // 2 * w, lambda: big[big > 0])
    report(f"X @ X[:, :64] {t}", N * w, lambda: (X @ X[:, :64]).run(), reps=2)
    del X, big
