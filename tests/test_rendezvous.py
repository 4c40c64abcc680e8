"""The sharding layer's process plumbing on CPU: TCP rendezvous between ranks (all-gather,
broadcast, barrier of small byte strings -- what carries NCCL ids and CUDA IPC handles) and the
block partition.  World sizes 2 and 3, one OS process per rank, no torch."""
import multiprocessing as mp
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from delayrepay_b200.sharding import Rendezvous
    rdv = Rendezvous(rank, world, "127.0.0.1", port, timeout=60)
    got = rdv.allgather(f"rank{rank}".encode() * (rank + 1))
    uid = rdv.bcast(b"\x07" * 128 if rank == 0 else b"ignored")
    for _ in range(3):
        rdv.barrier()
    big = rdv.allgather(bytes([rank]) * 100000)
    rdv.close()
    q.put((rank, got, uid, [len(b) for b in big], [b[:1] for b in big]))


@pytest.mark.parametrize("world", [2, 3])
def test_rendezvous_allgather_bcast_barrier(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, got, uid, lens, heads in res:
        assert got == [f"rank{r}".encode() * (r + 1) for r in range(world)]
        assert uid == b"\x07" * 128
        assert lens == [100000] * world and heads == [bytes([r]) for r in range(world)]


def test_shard_bounds_partition():
    from delayrepay_b200.sharding import shard_bounds
    for n in (0, 1, 7, 8, 100003, 32768):
        for world in (1, 2, 3, 8):
            parts = [shard_bounds(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_expressions_plan_and_compile_without_a_gpu():
    """Dry run: a 3-rank in-process mesh localises and compiles the heat stencil (halo variant),
    the fused reductions and the n-body step; launches are recorded, nothing executes."""
    import numpy as np
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine, sharding
    import workloads as wl
    with engine.dry_run() as log:
        sharding.init(devices=[0, 0, 0])
        try:
            u = dr.shard(wl.make_inputs("heat", 300)["u"])
            assert u.shape == (300, 300) and u.array.base.H == 1
            n0 = len(log)
            wl.heat(dr, u, 4)
            assert len(log) - n0 == 12 and all(k[0].name.startswith("dr_stencil_") for k in log[n0:])
            assert "dr_wait_epoch" in log[-1][0].source and "dr_st_release_sys" in log[-1][0].source
            a, b = dr.shard(np.arange(1000.0)), dr.shard(np.ones(1000))
            assert wl.l2_distance(dr, a, b).run().shape == ()
            assert wl.dot(dr, a, b).run().shape == ()
            assert type((a * 2 + b)._force()).__name__ == "ShardView"
            i = wl.make_inputs("nbody", 600)
            acc = wl.nbody_acc(dr, dr.shard(i["pos"], halo=0), dr.array(i["m"]))
            assert acc._force().shape == (600, 3)
            with pytest.raises(NotImplementedError):
                (u[3:] + u[:-3]).run()          # needs rows 3 away, the blocks hold 1
        finally:
            sharding.shutdown()
