"""Host-side cost of one axpy step (capture -> plan -> launch) on the GPU box, by phase.
usage: python tools/host_overhead.py"""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import delayrepay_b200 as dr
from delayrepay_b200 import planner, engine
import workloads as wl

dr.set_device(0)
n = 1 << 24
i = wl.make_inputs("axpy", n)
x, y = dr.array(i["x"]), dr.array(i["y"])
N = 2000


def t(label, fn, reps=N):
    fn()
    dr.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t1 = time.perf_counter()
    dr.synchronize()
    t2 = time.perf_counter()
    print(f"{label:42s} host {1e6 * (t1 - t0) / reps:7.1f} us/step   with drain {1e6 * (t2 - t0) / reps:7.1f}")


t("capture only (nodes die)", lambda: wl.axpy(dr, 1.5, x, y))
t("capture + run (full step)", lambda: wl.axpy(dr, 1.5, x, y).run())
r = wl.axpy(dr, 1.5, x, y)
t("build_program", lambda: planner.build_program([r]))
prog = planner.build_program([r])
outs = [dr.DeviceArray.empty(prog.shape, r.dtype, 0)]
t("run_program (layout+key+args+launch)", lambda: engine.run_program(prog, outs))
t("alloc + free", lambda: dr.DeviceArray.empty(prog.shape, r.dtype, 0))
del r
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    wl.axpy(dr, 1.5, x, y).run()
pr.disable()
dr.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
