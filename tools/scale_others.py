"""Multi-GPU timing of the sharded C3 (fused reduction + ncclAllReduce) and C4 (row-sharded heat
stencil with one halo row per neighbour per step) workloads; one process per GPU.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/scale_others.py

C3: weak scaling, 2^30 float64 elements per GPU.  C4: strong scaling, the 32768 x 32768 float32 grid
of BASELINE.json split into N row blocks, 100 steps.  Device time (CUDA events on the launch
stream), max over ranks."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as tdist
import delayrepay_b200 as dr
from delayrepay_b200 import dist as dd
import workloads as wl
from delayrepay_b200._lib import lib, check

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
dr.set_device(local)
comm = dd.NcclComm(rank, world, local) if world > 1 else None


def timed(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    a, b = C.c_uint64(), C.c_uint64()
    check(lib.drc_event_create(local, C.byref(a)))
    check(lib.drc_event_create(local, C.byref(b)))
    dr.synchronize()
    if world > 1:
        tdist.barrier()
    check(lib.drc_event_record(local, 0, a.value))
    for _ in range(reps):
        fn()
    check(lib.drc_event_record(local, 0, b.value))
    check(lib.drc_event_sync(local, b.value))
    ms = C.c_float()
    check(lib.drc_event_elapsed_ms(local, a.value, b.value, C.byref(ms)))
    t = torch.tensor([ms.value / reps], device="cuda")
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    return float(t.item())


out = {"n_gpus": world}
# ---- C3: ||a - b|| over 2^30 float64 per GPU
n = 1 << 30
i = wl.make_inputs("l2", 1 << 22, seed=3 + rank)
a = dr.tile(dr.array(i["a"]), n >> 22)
b = dr.tile(dr.array(i["b"]), n >> 22)
if world > 1:
    ms = timed(lambda: dd.sharded_l2_distance(dr, a, b, comm)._force(), 10, 2)
else:
    ms = timed(lambda: wl.l2_distance(dr, a, b).run(), 10, 2)
out["l2_distance_f64"] = {"ms": ms, "elems_per_s": world * n / (ms * 1e-3), "GBs_per_gpu": n * 16 / ms / 1e6}
del a, b
# ---- C4: heat 32768^2, 100 steps, row blocks
g = 32768
lo, hi = dd.shard_bounds(g, world, rank)
up, down, total = dd.halo_rows(rank, world, hi - lo)
u = dr.tile(dr.array(wl.make_inputs("heat", 2048)["u"]), (-(-total // 2048), g // 2048))
if u.shape[0] != total:
    u = dr.array(u._force()[:total].copy()) if hasattr(u, "_force") else u[:total]
steps = 100
if world > 1:
    dev = u._force()
    ms = timed(lambda: dd.sharded_heat(lambda blk: wl.heat_step(dr, u), dev, steps, comm,
                                       lambda blk, r: blk[r], lambda blk, r, buf: None), 1, 1)
else:
    ms = timed(lambda: wl.heat(dr, u, steps), 1, 1)
out["heat_f32_32768^2"] = {"ms_per_step": ms / steps, "cell_steps_per_s": g * g / (ms / steps * 1e-3),
                           "GBs_aggregate": g * g * 8 / (ms / steps) / 1e6}
if rank == 0:
    print(json.dumps(out))
if comm is not None:
    comm.close()
    tdist.destroy_process_group()
