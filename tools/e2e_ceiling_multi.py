"""The box's host<->device ceiling with N ranks copying at once (the bound of bench.py's e2e leg).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_ceiling_multi.py [log2n]

Every rank moves the Black-Scholes e2e traffic of 2^log2n options (12 B/option in, 8 B/option out)
between PINNED host buffers and its GPU through libdrcuda's two copy streams, both directions at
once, nothing else running -- no kernel, no Python per chunk.  Rank 0 prints one JSON line: the
aggregate rate over all ranks (wall clock between two cross-process barriers, best of 3).  The
ratio of these lines at N = 1, 2, 4, 8 is what e2e scaling can reach on this host at best.
No torch: ranks meet through the sharding layer's TCP rendezvous.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import delayrepay_b200 as dr
from delayrepay_b200._lib import check, lib
from delayrepay_b200.device import DeviceArray, bind_to_device_numa
from delayrepay_b200.sharding import Rendezvous
from delayrepay_b200.stream import H2D, D2H

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << log2n
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = int(os.environ.get("LOCAL_RANK", 0))
dr.set_device(dev)
numa = None if os.environ.get("DR_NO_NUMA_BIND") else bind_to_device_numa(dev)
rdv = Rendezvous(rank, world)
hin = [dr.pinned_empty(n, np.float32) for _ in range(3)]
hout = [dr.pinned_empty(n, np.float32) for _ in range(2)]
for h in hin:
    h[:] = 1.0
d_in = [DeviceArray.empty((n,), np.float32, dev) for _ in range(3)]
d_out = [DeviceArray.empty((n,), np.float32, dev) for _ in range(2)]


def sync():
    for s in (0, H2D, D2H):
        check(lib.drc_stream_sync(dev, s))


def both(piece=1 << 26):
    sync()
    rdv.barrier()
    t0 = time.perf_counter()
    for lo in range(0, n * 4, piece):
        m = min(piece, n * 4 - lo)
        for h, d in zip(hin, d_in):
            check(lib.drc_memcpy_h2d_async(dev, H2D, d.ptr + lo, h.ctypes.data + lo, m))
        for h, d in zip(hout, d_out):
            check(lib.drc_memcpy_d2h_async(dev, D2H, h.ctypes.data + lo, d.ptr + lo, m))
    sync()
    mine = time.perf_counter() - t0
    rdv.barrier()
    return time.perf_counter() - t0, mine


both()
runs = [both() for _ in range(3)]
wall = min(r[0] for r in runs)
per_rank = rdv.allgather(json.dumps(min(r[1] for r in runs)).encode())
if rank == 0:
    print(json.dumps({
        "n_ranks": world, "options_per_rank": n, "wall_ms": round(wall * 1e3, 2),
        "aggregate_options_per_s": world * n / wall,
        "aggregate_h2d_GBs": round(12 * n * world / wall / 1e9, 1),
        "aggregate_d2h_GBs": round(8 * n * world / wall / 1e9, 1),
        "per_rank_ms": [round(float(json.loads(b)) * 1e3, 1) for b in per_rank],
        "numa_cpus_rank0": None if numa is None else len(numa), "host_cpus": os.cpu_count()}), flush=True)
rdv.close()
