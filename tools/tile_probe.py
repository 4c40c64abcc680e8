"""`X.T + X` on 16384^2 float32 a few times: target of the ncu capture of the tile family
(profiles/r2_tile_family_xt_plus_x.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
dr.set_device(0)
X = dr.tile(dr.array(np.random.default_rng(0).standard_normal((1024, 1024)).astype(np.float32)), (16, 16)).run()
for _ in range(3):
    y = (X.T + X).run()
dr.synchronize()
print("done", y.shape)
