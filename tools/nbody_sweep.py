"""Time the n-body step at N = 65536 for the current DR_SK_* settings (CUDA events, 5 steps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delayrepay_b200 as dr
import workloads as wl
from delayrepay_b200._lib import lib, check
dr.set_device(0)
nb = 65536
i = wl.make_inputs("nbody", nb)
pos, m = dr.array(i["pos"]), dr.array(i["m"])
for _ in range(2):
    wl.nbody_acc(dr, pos, m).run()
a, b = C.c_uint64(), C.c_uint64()
check(lib.drc_event_create(0, C.byref(a))); check(lib.drc_event_create(0, C.byref(b)))
dr.synchronize()
check(lib.drc_event_record(0, 0, a.value))
for _ in range(5):
    wl.nbody_acc(dr, pos, m).run()
check(lib.drc_event_record(0, 0, b.value)); check(lib.drc_event_sync(0, b.value))
ms = C.c_float(); check(lib.drc_event_elapsed_ms(0, a.value, b.value, C.byref(ms)))
per = ms.value / 5
print({k: v for k, v in os.environ.items() if k.startswith("DR_SK")}, f"{per:.3f} ms/step  {nb*nb/per/1e6:.0f} Gpair/s")
