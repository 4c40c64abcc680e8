"""cumsum over 2^28 float32 (or float64 with argv[1] = f64) a few times: the target of the ncu
captures of the one-pass chained scan (profiles/r2_scan_chain_*.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
dr.set_device(0)
dt = np.float64 if len(sys.argv) > 1 and sys.argv[1] == "f64" else np.float32
x = dr.tile(dr.array(np.random.default_rng(0).standard_normal(1 << 20).astype(dt)), 256).run()
X = x.reshape(16384, 16384)
for _ in range(3):
    y = np.cumsum(x).run()
    z = np.cumsum(X, axis=1).run()
dr.synchronize()
print("done", y.shape)
