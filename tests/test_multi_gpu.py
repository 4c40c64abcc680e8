"""Two-GPU run of the sharding layer with the real back ends: CUDA engine + NCCL through
libdrcuda (one process per GPU).  Needs >= 2 devices (gpurun --gpus 2); on a 1-GPU box the test
reports itself skipped -- the host logic is covered by tests/test_dist_gloo.py."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as tdist
    import delayrepay_b200 as dr
    from delayrepay_b200 import dist as dd
    import workloads as wl
    tdist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dr.set_device(rank)
    comm = dd.NcclComm(rank, world, rank)
    res = {}
    n = (1 << 22) + 11
    inp = wl.make_inputs("l2", n)
    lo, hi = dd.shard_bounds(n, world, rank)
    a, b = dr.array(inp["a"][lo:hi]), dr.array(inp["b"][lo:hi])
    res["l2"] = float(dd.sharded_l2_distance(dr, a, b, comm).get())
    res["dot"] = float(dd.sharded_dot(dr, a, b, comm).get())
    rows, cols, steps = 301, 517, 9
    u0 = np.random.default_rng(4).random((rows, cols), dtype=np.float32)
    lo, hi = dd.shard_bounds(rows, world, rank)
    up, down, total = dd.halo_rows(rank, world, hi - lo)
    block = dr.array(u0[lo - int(up):hi + int(down)].copy())
    dev = block._force()
    dd.sharded_heat(lambda u: wl.heat_step(dr, block), dev, steps, comm,
                    lambda u, i: u[i], lambda u, i, buf: None)
    res["heat"] = block.get()[int(up):total - int(down)].copy()
    out[rank] = res
    comm.close()
    tdist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_sharded_reductions_and_heat(gpu):
    from delayrepay_b200._lib import init
    if init() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import workloads as wl
    from oracle import refcpu
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    inp = wl.make_inputs("l2", (1 << 22) + 11)
    want_l2 = float(wl.l2_distance(refcpu, refcpu.leaf(inp["a"]), refcpu.leaf(inp["b"])))
    want_dot = float(wl.dot(refcpu, refcpu.leaf(inp["a"]), refcpu.leaf(inp["b"])).get())
    for r in range(world):
        assert abs(out[r]["l2"] - want_l2) <= 1e-12 * want_l2
        assert abs(out[r]["dot"] - want_dot) <= 1e-12 * max(abs(want_dot), np.sqrt(1 << 22) * 1e-3)
    u0 = np.random.default_rng(4).random((301, 517), dtype=np.float32)
    want = wl.heat(refcpu, refcpu.leaf(u0.copy()), 9).get()
    got = np.concatenate([out[r]["heat"] for r in range(world)], axis=0)
    assert got.tobytes() == want.tobytes(), "2-GPU sharded heat is not bit-identical to the oracle"
