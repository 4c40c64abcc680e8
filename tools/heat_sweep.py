"""Time the heat stencil at 32768^2 for the current DR_ST_* settings (CUDA events, 20 steps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delayrepay_b200 as dr
import workloads as wl
from delayrepay_b200._lib import lib, check
dr.set_device(0)
g = 32768
u = dr.tile(dr.array(wl.make_inputs("heat", 2048)["u"]), (g // 2048, g // 2048))
wl.heat(dr, u, 3)
a, b = C.c_uint64(), C.c_uint64()
check(lib.drc_event_create(0, C.byref(a))); check(lib.drc_event_create(0, C.byref(b)))
dr.synchronize()
check(lib.drc_event_record(0, 0, a.value))
wl.heat(dr, u, 20)
check(lib.drc_event_record(0, 0, b.value)); check(lib.drc_event_sync(0, b.value))
ms = C.c_float(); check(lib.drc_event_elapsed_ms(0, a.value, b.value, C.byref(ms)))
per = ms.value / 20
print({k: v for k, v in os.environ.items() if k.startswith("DR_ST")}, f"{per:.3f} ms/step  {g*g*8/per/1e6:.0f} GB/s")
