"""Host logic of the leading-axis sharding layer on CPU: world_size 2 over gloo.

The collective back end here is GlooComm on host arrays and the local compute step is the
oracle (NumPy); on the GPU box the same functions run with NcclComm and the CUDA engine
(tests/test_multi_gpu.py).  What is verified: block partitioning, partial-sum combination,
halo bookkeeping / exchange order, and that the sharded heat iteration is bit-identical to
the unsharded reference."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from delayrepay_b200 import dist as dd
    import workloads as wl
    from oracle import refcpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    comm = dd.GlooComm()
    res = {}
    # ---- reductions: shard, local partial, all-reduce
    n = 100003
    inp = wl.make_inputs("l2", n)
    lo, hi = dd.shard_bounds(n, world, rank)
    a, b = refcpu.leaf(inp["a"][lo:hi]), refcpu.leaf(inp["b"][lo:hi])
    res["l2"] = float(dd.sharded_l2_distance(refcpu, a, b, comm))
    res["dot"] = float(dd.sharded_dot(refcpu, a, b, comm))
    # ---- heat: row blocks with halos
    rows, cols, steps = 37, 29, 6
    u0 = wl.make_inputs("heat", 64)["u"][:rows, :cols].copy()
    lo, hi = dd.shard_bounds(rows, world, rank)
    up, down, total = dd.halo_rows(rank, world, hi - lo)
    block = u0[lo - int(up):hi + int(down)].copy()
    assert block.shape[0] == total
    leafed = refcpu.leaf(block)

    def step(u):
        wl.heat_step(refcpu, leafed)

    def getrow(u, i):
        return u[i]

    def setrow(u, i, buf):
        u[i] = buf
    dd.sharded_heat(step, block, steps, comm, getrow, setrow)
    res["heat_rows"] = (lo, hi)
    res["heat"] = block[int(up):block.shape[0] - int(down)].copy()
    out[rank] = res
    dist.destroy_process_group()


def test_shard_bounds_cover_the_axis():
    from delayrepay_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 1 << 30, (1 << 30) + 5):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_sharded_reductions_and_heat_world2():
    import torch.multiprocessing as mp
    import workloads as wl
    from oracle import refcpu
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    inp = wl.make_inputs("l2", 100003)
    want_l2 = float(wl.l2_distance(refcpu, refcpu.leaf(inp["a"]), refcpu.leaf(inp["b"])))
    want_dot = float(wl.dot(refcpu, refcpu.leaf(inp["a"]), refcpu.leaf(inp["b"])).get())
    for r in range(world):
        assert abs(out[r]["l2"] - want_l2) <= 1e-12 * want_l2
        assert abs(out[r]["dot"] - want_dot) <= 1e-12 * max(abs(want_dot), 1.0)
    u0 = wl.make_inputs("heat", 64)["u"][:37, :29].copy()
    want = wl.heat(refcpu, refcpu.leaf(u0.copy()), 6).get()
    got = np.concatenate([out[r]["heat"] for r in range(world)], axis=0)
    assert got.shape == want.shape and got.tobytes() == want.tobytes(), "sharded heat differs"
