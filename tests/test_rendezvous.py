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


def _spmd_dry_worker(rank, world, port, q):
    """One rank of an SPMD run in dry mode: everything but the kernels -- rendezvous, partition,
    neighbour wiring, localisation, planning and NVRTC compilation of the halo stencil."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import numpy as np
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine, sharding
    import workloads as wl
    with engine.dry_run() as log:
        mesh = sharding.init()
        assert mesh.spmd and mesh.world == world and mesh.local == [rank]
        h = wl.make_inputs("heat", 516)["u"][:515]          # 515 x 516: uneven row split, vectorisable columns
        u = dr.shard(h)
        base = u.array.base
        link = base.links[rank]
        n0 = len(log)
        wl.heat(dr, u, 4)
        launched = [k[0].name for k in log[n0:]]
        a, b = dr.shard(np.arange(1001.0)), dr.shard(np.ones(1001))
        wl.l2_distance(dr, a, b).run()
        mesh.barrier()
        q.put((rank, base.bounds, base.blocks[rank].shape, link.up is not None, link.dn is not None,
               (link.up or {}).get("rows"), (link.dn or {}).get("rows"), link.epoch, launched,
               a.array.base.bounds))
        sharding.shutdown()


def test_spmd_two_processes_plan_and_wire_without_a_gpu():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_spmd_dry_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, r1) = res
    assert r0[1] == r1[1] == [(0, 258), (258, 515)]                      # the same partition on both ranks
    assert r0[2] == (258 + 2, 516) and r1[2] == (257 + 2, 516)          # owned rows + one halo row per side
    assert (r0[3], r0[4], r1[3], r1[4]) == (False, True, True, False)   # neighbours: rank 0 below-only, rank 1 above-only
    assert r0[6] == 259 and r1[5] == 260                                # each knows the other's block height
    assert r0[7] == r1[7] == 4                                          # four halo-stencil steps, epochs agree
    assert all(len(r[8]) == 4 and all(n.startswith("dr_stencil_") for n in r[8]) for r in res)   # one launch per step
    assert r0[9] == r1[9] == [(0, 501), (501, 1001)]
