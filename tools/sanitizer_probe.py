"""Small invocations of the round-2 kernels (tile family, row scan, chained scan, vector gather,
vector argmax) for compute-sanitizer (memcheck / racecheck); results are checked against NumPy.
usage: compute-sanitizer --tool racecheck python tools/sanitizer_probe.py [chain]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
dr.set_device(0)
rng = np.random.default_rng(1)
for dt, (r, c) in ((np.float32, (130, 257)), (np.float32, (128, 192)), (np.float64, (70, 99)), (np.int32, (64, 64))):
    a = (rng.standard_normal((r, c)) * 9).astype(dt); b = (rng.standard_normal((c, r)) * 9).astype(dt)
    assert np.array_equal((dr.array(b).T + dr.array(a)).get(), b.T + a)
    assert np.array_equal(dr.array(b).T.copy().get(), b.T)
m = rng.integers(-9, 9, (70, 4100)).astype(np.int32)
assert np.array_equal(np.cumsum(dr.array(m), axis=1).get(), np.cumsum(m, axis=1))
x = (rng.standard_normal((40, 64)) * 9).astype(np.float32)
idx = rng.integers(-40, 40, 77)
assert np.array_equal(dr.array(x)[dr.array(idx)].get(), x[idx])
v = rng.standard_normal(100_003).astype(np.float32); v[5000] = np.nan
assert int(np.argmax(dr.array(v))) == 5000 and int(np.argmin(dr.array(v[:4999]))) == int(np.argmin(v[:4999]))
if len(sys.argv) > 1 and sys.argv[1] == "chain":
    xi = rng.integers(-9, 9, (1 << 20) + 3).astype(np.int32)
    assert np.array_equal(np.cumsum(dr.array(xi)).get(), np.cumsum(xi))
dr.synchronize()
print("probe ok")
