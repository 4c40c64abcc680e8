"""Leading-axis sharding across the GPUs of one box, one process per GPU (torchrun).

The reference has no multi-device support at all; this layer is the
north_star's item 4.  Arrays are split along axis 0 into contiguous blocks, one per rank:

  * elementwise regions run on the local block with ZERO communication;
  * reductions finish with one NCCL all-reduce of the per-GPU partial (a 0-d device array);
  * slice stencils exchange one boundary row with each neighbour per step (ncclSend/ncclRecv
    over NVLink), everything else is the local fused kernel.

Two communicator back ends behind one interface: ``NcclComm`` (libdrcuda, device buffers --
the product path) and ``GlooComm`` (torch.distributed on host arrays) which exists so the host
logic -- partitioning, partial combination, halo bookkeeping -- is testable on a CPU box with
world_size 2.
"""
import ctypes as C
import os

import numpy as np

_NCCL_DT = {"float32": 0, "float64": 1, "int32": 2, "int64": 3, "uint8": 4}
_OPS = {"sum": 0, "prod": 1, "max": 2, "min": 3}


def shard_bounds(n, world, rank):
    """Contiguous block split of ``n`` leading-axis units: the first n % world ranks get one
    extra unit.  Returns (lo, hi)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


class GlooComm:
    """Host-array communicator over an initialised torch.distributed (gloo) group."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allreduce(self, value, op="sum"):
        import torch
        t = torch.from_numpy(np.array(value, copy=True, ndmin=1))
        self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX,
                                    "min": self.dist.ReduceOp.MIN, "prod": self.dist.ReduceOp.PRODUCT}[op])
        out = t.numpy()
        return out.reshape(np.shape(value))

    def sendrecv(self, send, send_peer, recv_like, recv_peer):
        """Exchange with neighbours; a peer of -1 means "no neighbour on that side"."""
        import torch
        reqs, out = [], None
        if recv_peer >= 0:
            out = torch.empty(tuple(recv_like.shape), dtype=torch.from_numpy(np.empty(0, recv_like.dtype)).dtype)
            reqs.append(self.dist.irecv(out, src=recv_peer))
        if send_peer >= 0:
            reqs.append(self.dist.isend(torch.from_numpy(np.ascontiguousarray(send)), dst=send_peer))
        for r in reqs:
            r.wait()
        return None if out is None else out.numpy()

    def barrier(self):
        self.dist.barrier()


class NcclComm:
    """Device-buffer communicator: NCCL through libdrcuda, stream-ordered on the compute stream.
    The 128-byte unique id travels through ``exchange`` (rank 0's bytes -> every rank), by default
    a torch.distributed broadcast on the already initialised default group."""

    def __init__(self, rank, world, dev, exchange=None):
        from ._lib import lib, check, init
        init()
        self.lib, self.check = lib, check
        self.rank, self.world, self.dev = rank, world, dev
        uid = (C.c_uint8 * 128)()
        if rank == 0:
            check(lib.drc_nccl_get_unique_id(uid))
        raw = bytes(uid)
        raw = (exchange or self._torch_exchange)(raw)
        buf = (C.c_uint8 * 128).from_buffer_copy(raw)
        comm = C.c_uint64()
        check(lib.drc_nccl_init_rank(dev, world, rank, buf, C.byref(comm)))
        self.comm = comm.value

    @staticmethod
    def _torch_exchange(raw):
        import torch
        import torch.distributed as dist
        t = torch.tensor(list(raw), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    def allreduce(self, arr, op="sum"):
        """In-place all-reduce of a contiguous DeviceArray; returns it."""
        self.check(self.lib.drc_nccl_allreduce(self.comm, self.dev, 0, arr.ptr, arr.ptr, arr.size,
                                               _NCCL_DT[arr.dtype.name], _OPS[op]))
        return arr

    def sendrecv(self, send, send_peer, recv, recv_peer):
        """send / recv: contiguous DeviceArrays (or None with peer -1)."""
        self.check(self.lib.drc_nccl_sendrecv(
            self.comm, self.dev, 0, send.ptr if send is not None else 0,
            send.nbytes if send is not None else 0, send_peer,
            recv.ptr if recv is not None else 0, recv.nbytes if recv is not None else 0, recv_peer))
        return recv

    def barrier(self):
        from .device import synchronize
        synchronize(self.dev)

    def group(self):
        """Context manager: the send/recv pairs issued inside become ONE NCCL launch."""
        comm = self

        class _Group:
            def __enter__(self):
                comm.check(comm.lib.drc_nccl_group_start())

            def __exit__(self, *exc):
                comm.check(comm.lib.drc_nccl_group_end())
        return _Group()

    def close(self):
        if getattr(self, "comm", 0):
            self.lib.drc_nccl_destroy(self.comm)
            self.comm = 0


# ------------------------------------------------------------------------------ sharded workloads
def sharded_l2_distance(xp, a_local, b_local, comm):
    """|| a - b ||_2 over arrays sharded on axis 0: local fused (a-b)^2 sum, one all-reduce."""
    part = xp.sum((a_local - b_local) ** 2)
    if isinstance(comm, NcclComm):
        dev = part._force()
        comm.allreduce(dev, "sum")
        return xp.sqrt(xp.NPArray(dev))
    return np.sqrt(comm.allreduce(np.asarray(part), "sum"))


def sharded_dot(xp, a_local, b_local, comm):
    part = xp.dot(a_local, b_local)
    if isinstance(comm, NcclComm):
        return xp.NPArray(comm.allreduce(part._force(), "sum"))
    return comm.allreduce(np.asarray(part), "sum")


def halo_rows(rank, world, rows_local):
    """Row layout of a rank's block WITH halos: returns (has_up, has_down, total_rows); the
    block is stored as [up halo?] + rows_local + [down halo?]."""
    up, down = rank > 0, rank < world - 1
    return up, down, rows_local + int(up) + int(down)


def exchange_halos(u, comm, getrow, setrow):
    """One halo exchange for a row-sharded 2-d block ``u`` (with halo rows, see halo_rows).
    ``getrow(u, i)`` returns row i as a contiguous buffer, ``setrow(u, i, buf)`` stores one;
    the same code drives device blocks (NcclComm) and host blocks (GlooComm)."""
    up, down, _ = halo_rows(comm.rank, comm.world, 0)
    n = u.shape[0]
    first, last = (1 if up else 0), (n - 2 if down else n - 1)
    if isinstance(comm, NcclComm):
        # both directions in one NCCL group: one launch per step, receives land in the halo rows
        with comm.group():
            comm.sendrecv(getrow(u, first) if up else None, comm.rank - 1 if up else -1,
                          getrow(u, n - 1) if down else None, comm.rank + 1 if down else -1)
            comm.sendrecv(getrow(u, last) if down else None, comm.rank + 1 if down else -1,
                          getrow(u, 0) if up else None, comm.rank - 1 if up else -1)
        return u
    # phase 1: send my first interior row up, receive my down halo from below
    got = comm.sendrecv(getrow(u, first) if up else None, comm.rank - 1 if up else -1,
                        getrow(u, n - 1) if down else None, comm.rank + 1 if down else -1)
    if down:
        setrow(u, n - 1, got)
    # phase 2: send my last interior row down, receive my up halo from above
    got = comm.sendrecv(getrow(u, last) if down else None, comm.rank + 1 if down else -1,
                        getrow(u, 0) if up else None, comm.rank - 1 if up else -1)
    if up:
        setrow(u, 0, got)
    return u


def sharded_heat(step_fn, u, steps, comm, getrow, setrow):
    """Row-sharded Jacobi iteration: exchange halos, then ``step_fn(u)`` updates the block's
    interior (the same slice arithmetic as the unsharded workload, fused into one kernel)."""
    for _ in range(steps):
        exchange_halos(u, comm, getrow, setrow)
        step_fn(u)
    return u
