"""Where does a sharded heat step spend its time?  1 GPU: unsharded vs an in-process mesh whose
ranks share the device (same kernels, same flags, no NVLink), plus the host-only cost per step
(dry run).   python tools/heat_shard_probe.py [rows] [ranks]"""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
from delayrepay_b200 import engine, sharding
from delayrepay_b200._lib import lib, check
import workloads as wl

g = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dr.set_device(0)
h0 = wl.make_inputs("heat", 2048)["u"]


def ev():
    e = C.c_uint64()
    check(lib.drc_event_create(0, C.byref(e)))
    return e.value


def timed(fn, steps):
    fn(3)
    a, b = ev(), ev()
    dr.synchronize()
    check(lib.drc_event_record(0, 0, a))
    t0 = time.perf_counter()
    fn(steps)
    host = time.perf_counter() - t0
    check(lib.drc_event_record(0, 0, b))
    check(lib.drc_event_sync(0, b))
    ms = C.c_float()
    check(lib.drc_event_elapsed_ms(0, a, b, C.byref(ms)))
    return ms.value / steps, host / steps * 1e3


u = dr.tile(dr.array(h0), (g // 2048, 16))
print("unsharded %d x 32768: device %.4f ms/step, host issue %.4f ms/step" % ((g,) + timed(lambda n: wl.heat(dr, u, n), 50)))
del u
mesh = sharding.init(devices=[0] * ranks)
su = sharding.from_global_fn(lambda r0, r1: dr.tile(dr.array(h0), ((r1 - r0) // 2048 + 2, 16))[(r0 % 2048):(r0 % 2048) + (r1 - r0)],
                             (g, 32768), np.float32)
print("%d ranks on one GPU:    device %.4f ms/step, host issue %.4f ms/step" % ((ranks,) + timed(lambda n: wl.heat(dr, su, n), 50)))
del su
sharding.shutdown()
with engine.dry_run():
    mesh = sharding.init(devices=[0])
    mesh.world = 1
    v = dr.array(h0)
    wl.heat(dr, v, 3)
    t0 = time.perf_counter(); wl.heat(dr, v, 200); print("host only, unsharded: %.1f us/step" % ((time.perf_counter() - t0) / 200 * 1e6))
    m2 = sharding.init(devices=[0, 0])
    s2 = dr.shard(h0)
    wl.heat(dr, s2, 3)
    t0 = time.perf_counter(); wl.heat(dr, s2, 200); print("host only, 2 local ranks: %.1f us/step" % ((time.perf_counter() - t0) / 200 * 1e6))
