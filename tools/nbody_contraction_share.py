"""How much of the n-body step (BASELINE config 5) is the contraction?  Times the fused
mm_skinny kernel as shipped and with its contraction FMAs removed (DR_SK_NOCONTRACT=1, results
wrong on purpose): the difference bounds what moving `W @ pos` to tcgen05 could gain.
Run each variant in its own process:  python tools/nbody_contraction_share.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r"""
import sys, ctypes as C
sys.path.insert(0, %r)
import delayrepay_b200 as dr, workloads as wl
from delayrepay_b200._lib import lib, check
dr.set_device(0)
i = wl.make_inputs("nbody", 65536)
pos, m = dr.array(i["pos"]), dr.array(i["m"])
for _ in range(3):
    wl.nbody_acc(dr, pos, m).run()
a, b = C.c_uint64(), C.c_uint64()
check(lib.drc_event_create(0, C.byref(a))); check(lib.drc_event_create(0, C.byref(b)))
dr.synchronize(); check(lib.drc_event_record(0, 0, a.value))
for _ in range(10):
    wl.nbody_acc(dr, pos, m).run()
check(lib.drc_event_record(0, 0, b.value)); check(lib.drc_event_sync(0, b.value))
ms = C.c_float(); check(lib.drc_event_elapsed_ms(0, a.value, b.value, C.byref(ms)))
print("%%.4f" %% (ms.value / 10))
""" % ROOT
res = {}
for name, env in (("full", {}), ("producer_only", {"DR_SK_NOCONTRACT": "1"})):
    out = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, **env), capture_output=True, text=True)
    res[name] = float(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else out.stderr[-400:]
    print(name, res[name], "ms per step (65536 bodies, 4.29 G pairs)")
if all(isinstance(v, float) for v in res.values()):
    print("contraction share of the step: %.1f %%" % (100 * (1 - res["producer_only"] / res["full"])))
