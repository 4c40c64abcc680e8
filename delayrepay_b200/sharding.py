"""Leading-axis sharding layer (north_star item 4).

The reference has no multi-device support at all.  Here an array can be
split along axis 0 into one contiguous row block per GPU of the box and then used through the
SAME drop-in API: ``u = dr.shard(u)`` and the script goes on unchanged.

How it plugs into the engine
  * ``ShardView`` is the "backend array" of a sharded graph leaf (``NPArray(ShardView)``): global
    shape and dtype, no data of its own; every node captured above such a leaf carries the
    mesh (``_mesh``), and ``engine.run`` hands those nodes to ``run`` below.
  * elementwise regions are LOCALISED: the same expression is rebuilt over each rank's local
    row block (plain DeviceArray views) and evaluated by the ordinary planner / code generator /
    plan cache -- zero communication;
  * reductions and ``dot`` finish with ONE all-reduce of the per-GPU partials through
    libdrcuda's NCCL communicator (``drc_nccl_allreduce``), the scalar epilogue (``sqrt``) then
    runs replicated;
  * slice stencils ``u[1:-1,1:-1] = f(u[2:,1:-1], u[:-2,1:-1], ...)`` run the TMA stencil kernel
    on every block in its halo variant (codegen.gen_stencil(halo=True)): blocks carry H halo rows
    per side, the kernel computes its boundary rows FIRST, stores them a second time straight into
    the neighbour GPU's block over NVLink (peer mapping) and publishes the step with a
    release/acquire flag while the interior is still being computed.  No NCCL, no host round
    trip, one launch per block per step.  A host-synchronised peer-copy exchange
    (``drc_memcpy_peer_async``) is the fallback after any other write to the array.

Two ways to own the GPUs, one code path:
  * SPMD   -- one process per GPU (torchrun): RANK / WORLD_SIZE / LOCAL_RANK from the environment;
              neighbours' blocks are mapped with CUDA IPC handles; the 64/128-byte handles and ids
              travel over a small TCP rendezvous on MASTER_ADDR (no torch anywhere);
  * in-process -- ``init(devices=[0, 1, ...])``: one process drives every device (peer access +
              ``drc_nccl_init_all``); a device may appear several times, which runs the whole
              protocol -- partitioning, localisation, halo pushes, flags -- on a single GPU.
"""
import ctypes as C
import os
import pickle
import socket
import struct
import time
import weakref

import numpy as np

from . import _lib
from ._lib import check, lib
from .device import DeviceArray, DeviceBuffer, current_device

_NCCL_DT = {"float32": 0, "float64": 1, "int32": 2, "int64": 3, "uint8": 4}
_OPS = {"sum": 0, "prod": 1, "max": 2, "min": 3}


def shard_bounds(n, world, rank):
    """Contiguous block split of ``n`` leading-axis units: the first n % world ranks get one
    extra unit.  Returns (lo, hi)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# ------------------------------------------------------------------------------ rendezvous
class Rendezvous:
    """All-gather of small byte strings between the ranks' processes over TCP (star through
    rank 0).  Carries the NCCL unique id, CUDA IPC handles and row counts -- never array data."""

    def __init__(self, rank, world, addr=None, port=None, timeout=180.0):
        self.rank, self.world = rank, world
        addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(port or os.environ.get("DR_RDV_PORT") or int(os.environ.get("MASTER_PORT", "29500")) + 211)
        self.peers = []
        if world == 1:
            return
        if rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(world)
            srv.settimeout(timeout)
            conns = {}
            while len(conns) < world - 1:
                c, _ = srv.accept()
                c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                r = struct.unpack("<i", self._recv_exact(c, 4))[0]
                conns[r] = c
            srv.close()
            self.peers = [conns[r] for r in range(1, world)]
        else:
            deadline = time.time() + timeout
            while True:
                try:
                    c = socket.create_connection((addr, port), timeout=5.0)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            c.settimeout(timeout)
            c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            c.sendall(struct.pack("<i", rank))
            self.peers = [c]

    @staticmethod
    def _recv_exact(c, n):
        buf = bytearray()
        while len(buf) < n:
            part = c.recv(n - len(buf))
            if not part:
                raise ConnectionError("rendezvous peer closed the connection")
            buf += part
        return bytes(buf)

    def _send(self, c, blob):
        c.sendall(struct.pack("<q", len(blob)) + blob)

    def _recv(self, c):
        n = struct.unpack("<q", self._recv_exact(c, 8))[0]
        return self._recv_exact(c, n)

    def allgather(self, blob):
        if self.world == 1:
            return [blob]
        if self.rank == 0:
            parts = [blob] + [self._recv(c) for c in self.peers]
            packed = pickle.dumps(parts)
            for c in self.peers:
                self._send(c, packed)
            return parts
        self._send(self.peers[0], blob)
        return pickle.loads(self._recv(self.peers[0]))

    def bcast(self, blob):
        return self.allgather(blob if self.rank == 0 else b"")[0]

    def barrier(self):
        self.allgather(b"")

    def close(self):
        for c in self.peers:
            try:
                c.close()
            except OSError:
                pass
        self.peers = []


# ------------------------------------------------------------------------------ mesh
class Mesh:
    """The ranks of one sharded computation and who drives them.  ``local`` lists the ranks this
    process launches for, ``devs[rank]`` is the CUDA device of a rank."""

    def __init__(self, world, local, devs, rdv=None, comms=None):
        self.world, self.local, self.devs = world, list(local), dict(devs)
        self.rdv, self.comms = rdv, comms or {}
        self.spmd = rdv is not None and world > 1
        self._peer_enabled = set()

    def bounds(self, n, rank=None):
        return shard_bounds(n, self.world, self.local[0] if rank is None else rank)

    # ---- host-level collectives on small metadata
    def allgather_obj(self, per_local):
        """per_local: {rank: picklable}; returns the list over all ranks."""
        if self.spmd:
            return [pickle.loads(b) for b in self.rdv.allgather(pickle.dumps(per_local[self.local[0]]))]
        return [per_local[r] for r in range(self.world)]

    def sync_streams(self):
        for r in self.local:
            if self.devs[r] >= 0:
                check(lib.drc_stream_sync(self.devs[r], 0))

    def barrier(self):
        """Every rank's launch stream has drained, on every process."""
        self.sync_streams()
        if self.spmd:
            self.rdv.barrier()

    def enable_peer(self, a, b):
        """In-process meshes: device a may address device b's peer allocations."""
        if a != b and (a, b) not in self._peer_enabled and a >= 0:
            check(lib.drc_enable_peer_access(a, b))
            self._peer_enabled.add((a, b))

    # ---- data collectives
    def allreduce(self, parts, op="sum"):
        """parts: {rank: contiguous DeviceArray} for the local ranks, all of one shape/dtype;
        reduced in place over ALL ranks."""
        if self.world == 1:
            return parts
        first = parts[self.local[0]]
        if first.dev < 0:
            return parts                                     # dry run: planning only
        if self.comms:
            dt = _NCCL_DT[first.dtype.name]
            grouped = len(self.local) > 1
            if grouped:
                check(lib.drc_nccl_group_start())
            for r in self.local:
                p = parts[r]
                check(lib.drc_nccl_allreduce(self.comms[r], p.dev, 0, p.ptr, p.ptr, p.size, dt, _OPS[op]))
            if grouped:
                check(lib.drc_nccl_group_end())
            return parts
        # in-process mesh whose ranks share devices (NCCL refuses duplicate devices): fold the
        # partials in rank order on the first rank's device, then hand the total back
        from .delayarray import NPArray
        uf = {"sum": np.add, "prod": np.multiply, "max": np.maximum, "min": np.minimum}[op]
        home = first.dev
        total = None
        for r in self.local:
            node = NPArray(self._on_device(parts[r], home))
            total = node if total is None else uf(total, node)
        res = total._force()
        for r in self.local:
            self._copy(parts[r], res)
        return parts

    def _on_device(self, arr, dev):
        if arr.dev == dev:
            return arr
        out = DeviceArray.empty(arr.shape, arr.dtype, dev)
        self._copy(out, arr)
        return out

    def _copy(self, dst, src):
        """dst[...] = src for contiguous equally sized arrays, across devices if need be."""
        if dst.nbytes == 0 or dst.dev < 0:
            return
        if dst.dev == src.dev:
            check(lib.drc_memcpy_d2d_async(dst.dev, 0, dst.ptr, src.ptr, dst.nbytes))
        else:
            # the source device's stream carries the copy, so it is ordered after the producer;
            # the consumer device then waits for it
            self.enable_peer(src.dev, dst.dev)
            check(lib.drc_memcpy_peer_async(dst.dev, dst.ptr, src.dev, src.ptr, dst.nbytes, src.dev, 0))
            check(lib.drc_stream_sync(src.dev, 0))

    def close(self):
        for c in set(self.comms.values()):
            lib.drc_nccl_destroy(c)
        self.comms = {}
        if self.rdv is not None:
            self.rdv.close()


_state = {"mesh": None}


def init(devices=None, rank=None, world=None, local_rank=None):
    """Create (or return) the process-wide mesh.

    ``devices=[...]``: in-process mesh, rank i on device devices[i] (repeats allowed).
    Otherwise SPMD from RANK / WORLD_SIZE / LOCAL_RANK (torchrun); world 1 = a trivial mesh."""
    if _state["mesh"] is not None and devices is None and rank is None:
        return _state["mesh"]
    shutdown()                                   # a new mesh replaces the old one (its communicators are destroyed)
    from . import engine
    dry = engine.is_dry()
    if devices is not None:
        devices = [(-1 if dry else int(d)) for d in devices]
        world = len(devices)
        comms = {}
        if not dry and world > 1 and len(set(devices)) == world:
            _lib.init()
            arr = (C.c_int * world)(*devices)
            out = (C.c_uint64 * world)()
            check(lib.drc_nccl_init_all(world, arr, out))
            comms = {r: out[r] for r in range(world)}
        mesh = Mesh(world, range(world), dict(enumerate(devices)), None, comms)
    else:
        rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
        dev = int(os.environ.get("LOCAL_RANK", str(current_device()))) if local_rank is None else local_rank
        if dry:
            dev = -1
        rdv, comms = None, {}
        if world > 1:
            rdv = Rendezvous(rank, world)
            if not dry:            # (dry run: the ranks still meet and exchange metadata, nothing touches a GPU)
                _lib.init()
                uid = (C.c_uint8 * 128)()
                if rank == 0:
                    check(lib.drc_nccl_get_unique_id(uid))
                raw = rdv.bcast(bytes(uid))
                comm = C.c_uint64()
                check(lib.drc_nccl_init_rank(dev, world, rank, (C.c_uint8 * 128).from_buffer_copy(raw), C.byref(comm)))
                comms = {rank: comm.value}
        mesh = Mesh(world, [rank], {rank: dev}, rdv, comms)
    _state["mesh"] = mesh
    return mesh


def shutdown():
    m, _state["mesh"] = _state["mesh"], None
    if m is not None:
        m.close()


def current_mesh():
    return _state["mesh"] or init()


# ------------------------------------------------------------------------------ halo link
_FLAG_BYTES = 256          # u32 slots 64 B apart: [0] from-up flag, [16] from-down flag, [32], [48] tile counters


class HaloLink:
    """What the halo variant of the stencil kernel needs to know about one block's neighbours
    (DeviceBuffer.halo).  Allocation i of a neighbour (i = 0: the one it started on, 1: its
    ping-pong partner) is addressable from this rank's device; a stencil step number ``epoch``
    (equal on all ranks: SPMD programs take the same steps) decides which one is its output."""

    def __init__(self, base, rank):
        self.base = weakref.ref(base)
        self.rank, self.H = rank, base.H
        self.epoch = 0
        self.waited = 0                   # last step whose pushed halo rows a wait kernel has covered
        self.dirty = False                # halo rows stale: something else wrote the array
        self._partner = None
        self.flags = None                 # DeviceBuffer (peer): my flags + counters
        self.up = self.dn = None          # dict(bufs=[ptr0, ptr1], rows=total rows, flags=ptr)

    @property
    def partner(self):
        return self._partner

    def before_stencil(self, max_dy):
        if max_dy > self.H:
            raise ValueError(f"the stencil reads {max_dy} rows away but the array was sharded with "
                             f"halo={self.H}; use dr.shard(x, halo={max_dy})")
        base = self.base()
        if base.halo_dirty:
            base.exchange_halos()

    def after_stencil(self):
        self.epoch += 1
        self.dirty = False

    def kernel_args(self, rows, pitch, TH, tiles_x):
        """The Halo_<kernel> struct of codegen.gen_stencil(halo=True)."""
        H, out_idx = self.H, (self.epoch + 1) % 2
        up_rows = self.up["bufs"][out_idx] + (self.up["rows"] - H) * pitch if self.up else 0
        dn_rows = self.dn["bufs"][out_idx] if self.dn else 0
        up_flag = self.up["flags"] + 64 if self.up else 0          # their "from below" flag
        dn_flag = self.dn["flags"] if self.dn else 0               # their "from above" flag
        # tile rows that hold rows pushed to a neighbour (counted for the "boundary done" signal) ...
        up_set = sorted({r // TH for r in range(H, 2 * H)}) if self.up else []
        dn_set = sorted({r // TH for r in range(rows - 2 * H, rows - H)}) if self.dn else []
        # ... and all tile rows the edge loop must own: those plus the ones holding halo rows
        # (which this block never writes)
        edge = set(up_set) | set(dn_set)
        if self.up:
            edge |= {r // TH for r in range(0, H)}
        if self.dn:
            edge |= {r // TH for r in range(rows - H, rows)}
        prio = sorted(edge)
        assert len(prio) <= 4
        f = self.flags.ptr
        return struct.pack("<7QI4i4i4x", up_rows, dn_rows, up_flag, dn_flag, f, f + 64, f + 128,
                           (self.epoch + 1) & 0xFFFFFFFF, H, len(up_set) * tiles_x, len(dn_set) * tiles_x,
                           len(prio), *(prio + [0] * (4 - len(prio))))


# ------------------------------------------------------------------------------ sharded storage
class ShardedBase:
    """Storage of one sharded array: per local rank a block of (rows owned + 2 H halo rows) x tail.
    ``bounds[r]`` = global rows owned by rank r."""

    def __init__(self, mesh, gshape, dtype, bounds, H=0):
        self.mesh, self.gshape, self.dtype = mesh, tuple(int(s) for s in gshape), np.dtype(dtype)
        self.bounds, self.H = [tuple(b) for b in bounds], int(H)
        self.blocks = {}
        self.links = {}
        tail = 1
        for s in self.gshape[1:]:
            tail *= s
        self.pitch = tail * self.dtype.itemsize
        peer = self.H > 0 and mesh.world > 1
        self._kind = "peer" if (peer and not os.environ.get("DR_SHARD_POOL_MEMORY")) else "pool"   # knob: experiments only
        for r in mesh.local:
            lo, hi = self.bounds[r]
            rows = hi - lo + 2 * self.H
            buf = DeviceBuffer(rows * self.pitch, mesh.devs[r], kind=self._kind)
            self.blocks[r] = DeviceArray(buf, (rows,) + self.gshape[1:], self.dtype)
        if peer:
            self._link_neighbours()

    @classmethod
    def adopt(cls, mesh, gshape, dtype, bounds, blocks):
        """A halo-less base over blocks that already exist (results of local evaluations)."""
        self = cls.__new__(cls)
        self.mesh, self.gshape, self.dtype = mesh, tuple(int(s) for s in gshape), np.dtype(dtype)
        self.bounds, self.H, self.links = [tuple(b) for b in bounds], 0, {}
        tail = 1
        for s in self.gshape[1:]:
            tail *= s
        self.pitch = tail * self.dtype.itemsize
        self.blocks = {r: (b if b.is_contiguous else b.copy()) for r, b in blocks.items()}
        return self

    # ---- geometry
    def owned(self, r):
        """The rows rank r owns, as a view of its block."""
        lo, hi = self.bounds[r]
        return self.blocks[r][self.H:self.H + hi - lo]

    @property
    def halo_dirty(self):
        return any(l.dirty for l in self.links.values())

    def mark_written(self):
        for l in self.links.values():
            l.dirty = True

    # ---- neighbour wiring
    def _link_neighbours(self):
        mesh, H = self.mesh, self.H
        info = {}
        for r in mesh.local:
            blk = self.blocks[r]
            dev = blk.dev
            link = self.links[r] = HaloLink(self, r)
            link._partner = DeviceBuffer(blk.buf.nbytes, dev, kind=self._kind)
            link.flags = DeviceBuffer(_FLAG_BYTES, dev, kind=self._kind)
            blk.buf.halo = link
            lo, hi = self.bounds[r]
            if hi - lo < 2 * H:
                raise ValueError(f"rank {r} owns {hi - lo} rows: a block needs at least 2 x halo = {2 * H}")
            rec = {"rows": blk.shape[0], "dev": dev, "pid": os.getpid()}
            if dev >= 0:
                check(lib.drc_memset_async(dev, 0, link.flags.ptr, 0, _FLAG_BYTES))
                check(lib.drc_memset_async(dev, 0, link._partner.ptr, 0, link._partner.nbytes))
                rec["ptrs"] = (blk.buf.ptr, link._partner.ptr, link.flags.ptr)
                if mesh.spmd:
                    hs = []
                    for p in rec["ptrs"]:
                        h = (C.c_uint8 * 64)()
                        check(lib.drc_ipc_get_handle(dev, p, h))
                        hs.append(bytes(h))
                    rec["handles"] = hs
            else:
                rec["ptrs"] = (blk.buf.ptr, link._partner.ptr, link.flags.ptr)
            info[r] = rec
        mesh.sync_streams()
        table = mesh.allgather_obj(info)
        self._foreign = []
        for r in mesh.local:
            link, dev = self.links[r], self.blocks[r].dev
            for side, nb in (("up", r - 1), ("dn", r + 1)):
                if not 0 <= nb < mesh.world:
                    continue
                rec = table[nb]
                if mesh.spmd and dev >= 0:
                    ptrs = []
                    for h in rec["handles"]:
                        p = C.c_uint64()
                        check(lib.drc_ipc_open_handle(dev, (C.c_uint8 * 64).from_buffer_copy(h), C.byref(p)))
                        ptrs.append(p.value)
                        self._foreign.append((dev, p.value))
                else:
                    ptrs = list(rec["ptrs"])
                    mesh.enable_peer(dev, rec["dev"])
                setattr(link, side, {"bufs": ptrs[:2], "rows": rec["rows"], "flags": ptrs[2], "dev": rec["dev"]})
        mesh.barrier()

    def __del__(self):
        for dev, p in getattr(self, "_foreign", ()):
            try:
                lib.drc_ipc_close_handle(dev, p)
            except Exception:
                pass

    # ---- explicit halo exchange (fallback: after a write that was not the halo stencil)
    def exchange_halos(self):
        """Peer copies of the first / last H owned rows into the neighbours' current blocks,
        bracketed by barriers.  The hot loop never comes here: the stencil kernel pushes its
        boundary rows itself."""
        mesh, H, pitch = self.mesh, self.H, self.pitch
        mesh.barrier()                                  # neighbours no longer read their halos
        for r in mesh.local:
            link, blk = self.links[r], self.blocks[r]
            cur = link.epoch % 2
            n = blk.shape[0]
            if blk.dev < 0:
                continue
            for nb, src_row, dst_row in ((link.up, H, None), (link.dn, n - 2 * H, 0)):
                if nb is None:
                    continue
                drow = nb["rows"] - H if dst_row is None else dst_row
                dst = nb["bufs"][cur] + drow * pitch
                src = blk.buf.ptr + src_row * pitch
                if mesh.spmd or nb["dev"] == blk.dev:
                    check(lib.drc_memcpy_d2d_async(blk.dev, 0, dst, src, H * pitch))
                else:
                    check(lib.drc_memcpy_peer_async(nb["dev"], dst, blk.dev, src, H * pitch, blk.dev, 0))
        mesh.barrier()
        for l in self.links.values():
            l.dirty = False
            l.waited = l.epoch


# ------------------------------------------------------------------------------ the backend array
class ShardView:
    """Rows [r0, r1) of a ShardedBase, each row seen through a strided sub-view ``tail`` -- the
    backend array of a sharded leaf.  Slicing makes views; nothing here holds data."""

    _is_shard_view = True
    __array_priority__ = 60.0

    def __init__(self, base, r0=0, r1=None, tail=None):
        self.base = base
        self.r0, self.r1 = int(r0), int(base.gshape[0] if r1 is None else r1)
        if tail is None:
            shp = base.gshape[1:]
            st, acc = [], base.dtype.itemsize
            for n in reversed(shp):
                st.append(acc)
                acc *= max(n, 1)
            tail = (0, tuple(shp), tuple(reversed(st)))
        self.tail = tail                      # (byte offset, shape, byte strides) inside a row
        self.shape = (self.r1 - self.r0,) + tuple(tail[1])
        self.dtype = base.dtype

    # ---- what the capture layer asks of a backend array
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def mesh(self):
        return self.base.mesh

    def layout_key(self):
        return ("shard", id(self.base), self.r0, self.r1) + self.tail + (self.dtype.str,)

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"ShardView(shape={self.shape}, dtype={self.dtype}, ranks={self.base.mesh.world}, halo={self.base.H})"

    def __getattr__(self, name):
        raise AttributeError(f"'{name}' is not supported on a sharded array (sharded arrays support "
                             "elementwise expressions, reductions, dot/@ with a sharded left operand, "
                             "basic slicing and slice assignment); gather it first with dr.array(x.get())")

    # ---- local pieces
    def rows_of(self, r):
        """View-row range [i0, i1) whose data rank r OWNS."""
        lo, hi = self.base.bounds[r]
        return max(lo - self.r0, 0), max(min(hi - self.r0, self.r1 - self.r0), max(lo - self.r0, 0))

    def local(self, r, i0, i1):
        """DeviceArray view of view-rows [i0, i1) on rank r (may reach into its halo rows)."""
        base = self.base
        lo, hi = base.bounds[r]
        H = base.H
        l0, l1 = i0 + self.r0 - lo + H, i1 + self.r0 - lo + H
        blk = base.blocks[r]
        if H and (l0 < H or l1 > blk.shape[0] - H):
            _halo_reads.add(base)             # the caller refreshes stale halo rows before it launches
        if l0 < 0 or l1 > blk.shape[0] or l1 < l0:
            raise NotImplementedError(
                f"rows {i0 + self.r0}..{i1 + self.r0} are not resident on rank {r} (it holds "
                f"{lo - H}..{hi + H}): the expression needs a redistribution or a wider halo "
                f"(dr.shard(x, halo=k))")
        off, shp, st = self.tail
        return DeviceArray(blk.buf, (l1 - l0,) + shp, self.dtype, (base.pitch,) + st, blk.offset + l0 * base.pitch + off)

    def local_rows(self, g0, g1):
        """Global rows [g0, g1) of the BASE on the (first) local rank that owns them (tests)."""
        for r in self.base.mesh.local:
            lo, hi = self.base.bounds[r]
            if lo <= g0 and g1 <= hi:
                return ShardView(self.base, g0, g1).local(r, 0, g1 - g0)
        raise IndexError(f"rows {g0}..{g1} are not owned by a local rank")

    # ---- views
    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - sum(k is not None for k in key) + 1) + key[i + 1:]
        if not key:
            return self
        k0, rest = key[0], key[1:]
        if isinstance(k0, slice) and k0.step in (None, 1):
            a, b, _ = k0.indices(self.shape[0])
            b = max(a, b)
            tail = self.tail if not rest else self._tail_view(rest)
            return ShardView(self.base, self.r0 + a, self.r0 + b, tail)
        if isinstance(k0, (int, np.integer)):
            i = int(k0) + (self.shape[0] if k0 < 0 else 0)
            if not 0 <= i < self.shape[0]:
                raise IndexError(f"index {k0} is out of bounds for axis 0 with size {self.shape[0]}")
            row = ShardView(self.base, self.r0 + i, self.r0 + i + 1, self.tail if not rest else self._tail_view(rest))
            return _RowPick(row)
        if k0 is None:
            # x[None, ...]: the rows stop being the leading axis -> every rank needs all of them
            return replicate(self)[key]
        raise NotImplementedError("sharded arrays support unit-stride row slices, integer rows and newaxis "
                                  "as the first index")

    def _tail_view(self, rest):
        # stride arithmetic on a 1-element shadow of one row (no data is touched)
        off, shp, st = self.tail
        if any(isinstance(k, (list, np.ndarray, DeviceArray)) or hasattr(k, "kind") for k in rest):
            raise NotImplementedError("advanced indexing of a sharded array")
        base = np.empty(1, dtype=self.dtype)
        shadow = np.lib.stride_tricks.as_strided(base, shp, st)
        view = shadow[tuple(rest) + ((Ellipsis,) if not any(k is Ellipsis for k in rest) else ())]
        delta = view.__array_interface__["data"][0] - base.__array_interface__["data"][0]
        return (off + delta, tuple(view.shape), tuple(view.strides))

    def __setitem__(self, key, value):
        try:
            target = self._targets.get(key)
        except (TypeError, AttributeError):
            target = None
        if target is None:
            target = self if (key is Ellipsis or (isinstance(key, slice) and key == slice(None))) else self[key]
            try:
                hash(key)
                if "_targets" not in self.__dict__:
                    self.__dict__["_targets"] = {}
                if len(self._targets) < 64:
                    self._targets[key] = target
            except TypeError:
                pass
        if isinstance(target, _RowPick):
            target[...] = value
        elif isinstance(target, DeviceArray):
            raise NotImplementedError("assignment through a replicated view of a sharded array")
        else:
            assign(target, value)

    def fill(self, value):
        assign(self, value)

    # ---- data movement
    def get(self, out=None):
        """Gather to the host: every rank ends up with the whole array."""
        host = replicate(self).get()
        if out is not None:
            out[...] = host
            return out
        return host

    def __array__(self, dtype=None, copy=None):
        host = self.get()
        return host if dtype is None else host.astype(dtype)

    def copy(self):
        from .delayarray import NPArray, create_ex
        return run(create_ex(np.positive, [NPArray(self)]))

    def astype(self, dtype, copy=True):
        from .delayarray import CastEx, NPArray
        dtype = np.dtype(dtype)
        if dtype == self.dtype:
            return self.copy() if copy else self
        return run(CastEx(NPArray(self), dtype))


class _RowPick:
    """x[i] / x[i, ...] of a sharded array: one row, owned by one rank.  Assignable; reading it
    replicates the row."""

    _is_shard_view = False

    def __init__(self, row):
        self.row = row                           # ShardView with exactly one row

    def __setitem__(self, key, value):
        tgt = self.row
        if not (key is Ellipsis or key == slice(None)):
            tgt = self.row[(slice(None),) + (key if isinstance(key, tuple) else (key,))]
        assign(tgt, value, drop_row_axis=True)

    def materialise(self):
        return replicate(self.row)[0]


# ------------------------------------------------------------------------------ replicate / gather
_gathered = weakref.WeakKeyDictionary()      # base -> {view layout: (generation, DeviceArray)}


def replicate(view):
    """The whole of ``view`` as an ordinary DeviceArray on every local rank's device (returned:
    the first local rank's copy).  One all-gather (NCCL) or peer copies (in-process mesh).  The
    result is remembered until the array is written again (`generation` counts writes on every
    rank alike, so all ranks hit or miss together and the collective stays matched): config 5
    gathers `pos` once, not once per step."""
    base = view.base
    gen = getattr(base, "generation", 0)
    per = _gathered.setdefault(base, {})
    hit = per.get(view.layout_key())
    if hit is not None and hit[0] == gen:
        return hit[1]
    out = _replicate(view)
    if out.nbytes <= (64 << 20):               # big gathers (.get() of a whole grid) are not kept
        if len(per) > 32:
            per.clear()
        per[view.layout_key()] = (gen, out)
    return out


def _replicate(view):
    mesh, base = view.mesh, view.base
    counts = []
    for r in range(mesh.world):
        i0, i1 = view.rows_of(r)
        counts.append(i1 - i0)
    row_shape = view.shape[1:]
    row_items = 1
    for s in row_shape:
        row_items *= s
    row_bytes = row_items * view.dtype.itemsize
    pieces = {}
    for r in mesh.local:
        i0, i1 = view.rows_of(r)
        if i1 > i0:
            loc = view.local(r, i0, i1)
            pieces[r] = loc if loc.is_contiguous else loc.copy()
        else:
            dev = mesh.devs[r]
            pieces[r] = DeviceArray.empty((0,) + tuple(view.shape[1:]), view.dtype, dev if dev >= 0 else None)
    home = mesh.devs[mesh.local[0]]
    total = view.shape[0]
    out = DeviceArray.empty((total,) + row_shape, view.dtype, home if home >= 0 else None)
    if home < 0 or total == 0 or row_bytes == 0:
        return out
    if mesh.spmd:
        r = mesh.local[0]
        cmax = max(counts)
        if cmax * row_bytes == 0:
            return out
        send = pieces[r]
        if counts[r] < cmax:
            padded = DeviceArray.empty((cmax,) + row_shape, view.dtype, home)
            if send.nbytes:
                check(lib.drc_memcpy_d2d_async(home, 0, padded.ptr, send.ptr, send.nbytes))
            send = padded
        if all(c == cmax for c in counts):
            check(lib.drc_nccl_allgather(mesh.comms[r], home, 0, send.ptr, out.ptr, cmax * row_bytes))
            return out
        stage = DeviceArray.empty((mesh.world, cmax) + row_shape, view.dtype, home)
        check(lib.drc_nccl_allgather(mesh.comms[r], home, 0, send.ptr, stage.ptr, cmax * row_bytes))
        at = 0
        for q in range(mesh.world):
            if counts[q]:
                check(lib.drc_memcpy_d2d_async(home, 0, out.ptr + at * row_bytes,
                                               stage.ptr + q * cmax * row_bytes, counts[q] * row_bytes))
            at += counts[q]
        return out
    at = 0
    for r in range(mesh.world):
        if counts[r]:
            dst = DeviceArray(out.buf, (counts[r],) + row_shape, view.dtype, None, out.offset + at * row_bytes)
            mesh._copy(dst, pieces[r])
        at += counts[r]
    return out


_halo_reads = set()        # bases whose halo rows were addressed since the last _refresh_halos()


_WAIT_SRC = r"""
// One thread: spin until both neighbours have published stencil step `epoch` (or there is no
// neighbour on that side: null pointer).  Launched in front of a kernel that reads halo rows which
// the neighbours' stencil kernels wrote, when that kernel is not the halo stencil itself (which
// performs the same acquire inside).
extern "C" __global__ void NAME(const unsigned* flag_up, const unsigned* flag_dn, unsigned epoch) {
  const unsigned* f[2] = {flag_up, flag_dn};
  for (int s = 0; s < 2; ++s) {
    if (f[s] == nullptr) continue;
    unsigned long long t0 = 0;
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f[s]) : "memory");
      if ((int)(v - epoch) >= 0) break;
      __nanosleep(100);
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now; else if (now - t0 > 120000000000ull) __trap();
    }
  }
}
"""


def _wait_for_pushed_halos(base):
    """The halo rows of `base` were written by the neighbours' stencil kernels (no host sync): a
    consumer other than the next halo stencil of the same array must not start before their
    flags say so.  One 1-thread kernel per local block, once per stencil step."""
    from . import engine
    for r, link in base.links.items():
        if link.epoch == link.waited or (link.up is None and link.dn is None):
            continue
        blk = base.blocks[r]
        kern = engine.get_kernel(("halo_wait",), lambda name: _WAIT_SRC.replace("NAME", name))
        a = engine.Args()
        a.ptr(link.flags.ptr if link.up else 0)
        a.ptr(link.flags.ptr + 64 if link.dn else 0)
        a.scalar(link.epoch & 0xFFFFFFFF, np.uint32)
        engine.launch(kern, blk.dev, 1, 1, a)
        link.waited = link.epoch


def _refresh_halos(bases=None, stencil_target=None):
    """Before halo rows are read: the peer-copy exchange for every base whose halo copies are stale
    (collective: every rank reaches this point with the same set -- same program, same state),
    and a device-side wait for halo rows that were pushed by the neighbours' stencil kernels --
    except for `stencil_target`, whose own halo stencil performs that wait itself."""
    todo = list(_halo_reads if bases is None else bases)
    if bases is None:
        _halo_reads.clear()
    for b in todo:
        if b.halo_dirty:
            b.exchange_halos()
        elif b is not stencil_target and b.links:
            _wait_for_pushed_halos(b)
    return todo


_replicas = weakref.WeakKeyDictionary()      # DeviceBuffer -> {dev: (version, DeviceArray)}


def _on_device(mesh, arr, dev):
    """A replicated operand on the device of the rank that needs it (in-process meshes)."""
    if arr.dev == dev or arr.dev < 0 or dev < 0:
        return arr
    per = _replicas.setdefault(arr.buf, {})
    hit = per.get((dev, arr.layout_key()))
    if hit is not None and hit[0] == arr.buf.version:
        return hit[1]
    src = arr if arr.is_contiguous else arr.copy()
    out = DeviceArray.empty(src.shape, src.dtype, dev)
    mesh._copy(out, src)
    per[(dev, arr.layout_key())] = (arr.buf.version, out)
    return out


# ------------------------------------------------------------------------------ localisation
def _shard_leaves(node, found, seen):
    """Every sharded leaf view below ``node`` (cut nodes are evaluated separately)."""
    stack = [node]
    while stack:
        n = stack.pop()
        if id(n) in seen:
            continue
        seen.add(id(n))
        arr = n.__dict__.get("array")
        if arr is not None and getattr(arr, "_is_shard_view", False):
            found.append(arr)
            continue
        if n.kind == "ewise" and arr is None:
            stack.extend(n.children)
        elif arr is None and n.kind in ("reduce", "matmul") and n.__dict__.get("_mesh") is not None:
            arr = n._force()                  # a sharded cut point below: evaluate it now
            if getattr(arr, "_is_shard_view", False):
                found.append(arr)


def _partition(node, views, target=None):
    """Which result rows each rank computes: {rank: (i0, i1)} over ALL ranks.  With a target view
    the owner of each target row computes it; otherwise the rows follow the first sharded
    operand's base, shifted to the middle of the row offsets with which that base is read."""
    if target is not None:
        return {r: target.rows_of(r) for r in range(target.mesh.world)}
    first = views[0]
    n = first.shape[0]
    offs = [v.r0 for v in views if v.base is first.base]
    anchor = (min(offs) + max(offs)) // 2
    part = {}
    for r in range(first.mesh.world):
        lo, hi = first.base.bounds[r]
        i0 = min(max(lo - anchor, 0), n)
        part[r] = (i0, max(min(hi - anchor, n), i0))
    # the rows in front of the first block / behind the last one (anchor shifts) go to the edge ranks
    world = first.mesh.world
    part[0] = (0, part[0][1])
    part[world - 1] = (part[world - 1][0], n)
    return part


def _localise(node, r, i0, i1, nrows, memo, mesh):
    """``node`` rebuilt over rank r's rows [i0, i1) of the result."""
    from . import delayarray as da
    hit = memo.get(id(node))
    if hit is not None:
        return hit
    kind = node.kind
    if kind == "scalar":
        out = node
    else:
        arr = node.__dict__.get("array")
        if arr is None and kind in ("reduce", "matmul"):
            arr = node._force()                          # cut point: evaluated on its own
        if arr is not None:
            if getattr(arr, "_is_shard_view", False):
                if arr.shape[0] != nrows or len(arr.shape) != len(memo["__shape__"]):
                    raise NotImplementedError("a sharded operand must span the rows of the result "
                                              "(no broadcasting along or into the sharded axis)")
                out = da.NPArray(arr.local(r, i0, i1))
            else:
                if isinstance(arr, np.ndarray):
                    arr = node._force()
                full = len(arr.shape) == len(memo["__shape__"]) and arr.shape[0] == nrows and nrows != 1
                loc = _on_device(mesh, arr, mesh.devs[r])
                out = da.NPArray(loc[i0:i1]) if full else (node if loc is arr and kind == "leaf" else da.NPArray(loc))
        elif kind == "ewise":
            kids = [_localise(k, r, i0, i1, nrows, memo, mesh) for k in node.children]
            if isinstance(node, da.WhereEx):
                out = da.WhereEx(*kids)
            elif isinstance(node, da.CastEx):
                out = da.CastEx(kids[0], node.dtype)
            elif isinstance(node, da.RawOp):
                out = da.RawOp(node.op, kids[0])
            else:
                out = type(node)(node.func, *kids)
        else:
            raise NotImplementedError(f"cannot localise a {kind} node")
    memo[id(node)] = out
    return out


def _scalar_sig(ops):
    return tuple((type(o.val).__name__, o.val) for o in ops if o.kind == "scalar")


# ------------------------------------------------------------------------------ evaluation
def run(node):
    """engine.run for a node that has a sharded leaf below it."""
    kind = node.kind
    if kind == "leaf":
        return node.array
    if kind == "ewise":
        return _run_ewise(node)
    if kind == "reduce":
        return _run_reduce(node)
    if kind == "matmul":
        return _run_contraction(node)
    raise NotImplementedError(kind)


def run_many(nodes):
    """engine.run_many for sharded nodes: lazy elementwise roots of one shape are localised TOGETHER
    and co-evaluated per row block (Black-Scholes' call and put: one two-output kernel per block,
    20 B/option, exactly as unsharded).  Everything else is evaluated on its own."""
    from . import engine
    groups = {}
    for n in nodes:
        if n.kind == "ewise" and n.__dict__.get("array") is None:
            g = groups.setdefault(tuple(n.shape), [])
            if all(n is not m for m in g):
                g.append(n)
        else:
            n._force()
    for group in groups.values():
        if len(group) == 1:
            group[0]._force()
            continue
        views = []
        for n in group:
            _shard_leaves(n, views, set())
        if not views:
            for n in group:
                n._force()
            continue
        mesh = views[0].mesh
        part = _partition(group[0], views)
        nrows = group[0].shape[0]
        _halo_reads.clear()
        local = {}
        for r in mesh.local:
            i0, i1 = part[r]
            if i1 > i0:
                memo = {"__shape__": group[0].shape}          # shared: common subexpressions stay common
                local[r] = [_localise(n, r, i0, i1, nrows, memo, mesh) for n in group]
        _refresh_halos()
        results = {r: None for r in mesh.local}
        for r, roots in local.items():
            engine.run_many(roots)
            results[r] = [x._force() for x in roots]
        for j, n in enumerate(group):
            blocks = {}
            for r in mesh.local:
                if results[r] is not None:
                    blocks[r] = results[r][j]
                else:
                    dev = mesh.devs[r]
                    blocks[r] = DeviceArray.empty((0,) + tuple(n.shape[1:]), n.dtype, dev if dev >= 0 else None)
            n.array = ShardView(ShardedBase.adopt(mesh, n.shape, n.dtype, [part[q] for q in range(mesh.world)], blocks))


def _run_ewise(node):
    from . import engine
    views = []
    _shard_leaves(node, views, set())
    if not views:
        # everything sharded below sits behind replicated cut points: an ordinary evaluation
        mesh = node._mesh
        return _localise(node, mesh.local[0], 0, 0, -1, {"__shape__": node.shape}, mesh)._force()
    mesh = views[0].mesh
    nrows = node.shape[0]
    part = _partition(node, views)
    return ShardView(_evaluate_blocks(mesh, node, part, lambda r, i0, i1: _localise(
        node, r, i0, i1, nrows, {"__shape__": node.shape}, mesh), key=_graph_key("ew", node)))


_local_graphs = {}      # structural key -> {rank: (local root, extra lazy nodes to reset)}


def _graph_key(tag, node, *more):
    """Key under which the LOCAL lazy graphs of a sharded evaluation can be reused the next time
    the same expression comes by: the expression's structural signature (which names every sharded
    leaf by its storage), the identity of every other leaf, the scalar values.  None = do not
    cache (cut points below, host operands not uploaded yet, ...)."""
    d = node.__dict__
    sig = d.get("_psig")
    if sig is None or os.environ.get("DR_SHARD_NO_GRAPH_CACHE"):
        return None
    ident = []
    for o in d["_pops"]:
        if o.kind == "scalar":
            ident.append((type(o.val).__name__, o.val))
        else:
            arr = o.array
            if getattr(arr, "_is_shard_view", False):
                ident.append(arr.layout_key())
            elif isinstance(arr, DeviceArray):
                ident.append(("D", id(arr.buf), arr.offset))
            else:
                return None                  # a host operand: a new leaf (and upload) per capture
    return (tag, sig, tuple(ident)) + more


def _evaluate_blocks(mesh, node, part, build, key=None):
    """One local evaluation per rank (``build(r, i0, i1)`` -> the local lazy node, or a tuple
    (node, extra nodes that receive results)); the fresh results become the blocks of a new
    sharded array -- no copy, plan cache and all.  With a ``key`` the local lazy graphs are kept:
    the next evaluation of the same expression over the same storage skips localisation and node
    construction (their leaves are views of the blocks, which live as long as the arrays; a
    replayed root is reset to "not evaluated" first)."""
    _halo_reads.clear()
    cache = _local_graphs.get(key) if key is not None else None
    if cache is None:
        cache = {}
        for r in mesh.local:
            if part[r][1] > part[r][0]:
                got = build(r, *part[r])
                cache[r] = got if isinstance(got, tuple) else (got, ())
        cache["__reads__"] = set(_halo_reads)
        cache["__part__"] = dict(part)
        if key is not None:
            if len(_local_graphs) > 128:
                _local_graphs.clear()
            _local_graphs[key] = cache
    else:
        _halo_reads.update(cache["__reads__"])
    _refresh_halos()
    blocks, extras = {}, {}
    for r in mesh.local:
        if r in cache:
            root, more = cache[r]
            blocks[r] = root._force()
            extras[r] = [m.__dict__.get("array") for m in more]
            for n in (root,) + tuple(more):          # lazy again for the next replay
                n.__dict__.pop("array", None)
        else:
            dev = mesh.devs[r]
            blocks[r] = DeviceArray.empty((0,) + tuple(node.shape[1:]), node.dtype, dev if dev >= 0 else None)
    base = ShardedBase.adopt(mesh, node.shape, node.dtype, [part[r] for r in range(mesh.world)], blocks)
    base._extras = extras
    return base


def _run_reduce(node):
    from . import delayarray as da
    child = node.children[0]
    if node.op == "sum" and node.axes == (1,) and child.ndim == 2 and not node.keepdims \
            and child.kind == "ewise" and child.__dict__.get("array") is None:
        # a pending A @ B over the same lazy producer computes this row sum in the same pass
        # (the counterpart of the hook in engine._run_reduce)
        for cons in list(getattr(child, "_consumers", ())):
            if isinstance(cons, da.MMEx) and cons.arg1 is child and cons.__dict__.get("array") is None \
                    and cons.shape[1] <= 6:
                cons._force()
                if node.__dict__.get("array") is not None:
                    return node.array
    views = []
    _shard_leaves(child, views, set())
    if not views:                               # the sharded part is behind a cut: already replicated
        return da.ReduceEx(node.func, da.NPArray(child._force()), node.axes or None, node.keepdims, node.post)._force()
    mesh = views[0].mesh
    nrows = child.shape[0]
    part = _partition(child, views)
    over_rows = 0 in node.axes
    if not over_rows:
        # rows stay sharded: a local reduction per block
        return ShardView(_evaluate_blocks(mesh, node, part, lambda r, i0, i1: da.ReduceEx(
            node.func, _localise(child, r, i0, i1, nrows, {"__shape__": child.shape}, mesh),
            node.axes, node.keepdims, node.post)))
    _halo_reads.clear()
    lazy = {}
    for r in mesh.local:
        i0, i1 = part[r]
        if i1 <= i0:
            raise NotImplementedError(f"rank {r} holds no rows of the reduced array")
        local = _localise(child, r, i0, i1, nrows, {"__shape__": child.shape}, mesh)
        lazy[r] = da.ReduceEx(node.func, local, node.axes, node.keepdims, None)
    _refresh_halos()
    parts = {r: n._force() for r, n in lazy.items()}
    mesh.allreduce(parts, node.op)
    res = parts[mesh.local[0]]
    for r in mesh.local[1:]:
        _replicas.setdefault(res.buf, {})[(mesh.devs[r], res.layout_key())] = (res.buf.version, parts[r])
    if node.post == "mean":
        count = 1
        for ax in node.axes:
            count *= child.shape[ax]
        res = da.as_dtype(da.NPArray(res) / float(count), node.dtype)._force()
    return res


def _run_contraction(node):
    from . import delayarray as da
    from . import engine
    a, b = node.arg1, node.arg2

    def sharded(x):
        found = []
        _shard_leaves(x, found, set())
        return found
    va, vb = sharded(a), sharded(b)
    if isinstance(node, da.DotEx):
        red = da.ReduceEx(np.add, da.BinaryNumpyEx(np.multiply, a, b), None, False)
        out = red._force()
        return out if out.dtype == node.dtype else out.astype(node.dtype)
    if not va:
        raise NotImplementedError("contraction over a sharded axis with a replicated left operand")
    mesh = va[0].mesh
    if vb:                                      # the right operand is needed whole on every rank
        bnode = b if b.kind == "leaf" else da.NPArray(b._force())
        b = da.NPArray(replicate(bnode.array))
    nrows = a.shape[0]
    part = _partition(a, va)
    # a pending A.sum(1) over the same lazy producer (config 5: W @ pos and W.sum(1)) rides along
    # as a column of ones in the local skinny kernel (engine._try_mm_skinny finds the LOCAL
    # ReduceEx through the local producer's consumer set), so W is evaluated once per block
    rowsum = None
    if isinstance(node, da.MMEx) and a.kind == "ewise" and a.__dict__.get("array") is None:
        for cons in list(getattr(a, "_consumers", ())):
            if isinstance(cons, da.ReduceEx) and cons.op == "sum" and cons.axes == (1,) and not cons.keepdims \
                    and cons.post is None and cons.__dict__.get("array") is None and cons.dtype == node.dtype:
                rowsum = cons
                break
    def build(r, i0, i1):
        la = _localise(a, r, i0, i1, nrows, {"__shape__": a.shape}, mesh)
        extra = ()
        if rowsum is not None:
            extra = (da.ReduceEx(np.add, la, 1, False),)              # receives its result from the MMEx pass
        return type(node)(la, _localise(b, r, 0, b.shape[0], -1, {"__shape__": b.shape}, mesh)), extra
    key = None
    if a.kind == "ewise" and b.kind == "leaf" and isinstance(b.array, DeviceArray):
        key = _graph_key("mm", a, type(node).__name__, id(b.array.buf), b.array.offset, rowsum is not None)
    base = _evaluate_blocks(mesh, node, part, build, key=key)
    if rowsum is not None and base._extras and all(v and v[0] is not None for v in base._extras.values()):
        blocks = {}
        for r in mesh.local:
            if r in base._extras:
                blocks[r] = base._extras[r][0]
            else:
                dev = mesh.devs[r]
                blocks[r] = DeviceArray.empty((0,), rowsum.dtype, dev if dev >= 0 else None)
        rowsum.array = ShardView(ShardedBase.adopt(mesh, rowsum.shape, rowsum.dtype,
                                                   [part[r] for r in range(mesh.world)], blocks))
    return ShardView(base)


# ------------------------------------------------------------------------------ assignment
_assign_plans = {}


def assign(target, value, drop_row_axis=False):
    """target[...] = value for a ShardView target: every rank writes the target rows it owns,
    reading only resident rows (its own and its halo).  Shifted reads of the target's own array
    run as the halo-pushing stencil kernel (one launch per block per step)."""
    from . import delayarray as da
    from . import engine
    from numbers import Number
    base, mesh = target.base, target.mesh
    if isinstance(value, Number):
        node = None
    elif isinstance(value, da.DelayArray):
        node = value
    else:
        node = da.arg_to_numpy_ex(value if isinstance(value, (DeviceArray, np.ndarray)) else np.asarray(value))
    key = None
    if node is not None and node.kind == "ewise" and node.__dict__.get("_psig") is not None and not drop_row_axis:
        key = (node._psig, target.layout_key(), _scalar_sig(node._pops))
        plan = _assign_plans.get(key)
        if plan is not None:
            _refresh_halos(plan[1], stencil_target=plan[3])
            _write(base, plan[0])
            plan[2][0] = node             # keeps the global right-hand side hash-consed
            return
    nrows = target.shape[0]
    todo = []
    _halo_reads.clear()
    for r in mesh.local:
        i0, i1 = target.rows_of(r)
        if i1 <= i0:
            continue
        tloc = target.local(r, i0, i1)
        if drop_row_axis:
            tloc = tloc[0]
        if node is None:
            vloc = value
        elif node.kind == "scalar":
            vloc = node.val
        else:
            sharded_operand = node.__dict__.get("_mesh") is not None or \
                getattr(node.__dict__.get("array"), "_is_shard_view", False)
            if sharded_operand:
                vloc = _localise(node, r, i0, i1, nrows, {"__shape__": node.shape}, mesh)
            else:
                arr = node._force()
                full = arr.ndim == target.ndim and arr.shape[0] == nrows and nrows != 1
                loc = _on_device(mesh, arr, mesh.devs[r])
                vloc = da.NPArray(loc[i0:i1] if full else loc)
        todo.append((tloc, vloc))
    # a self-stencil on an array with links runs as the halo kernel, which waits for the pushed rows
    # itself; the first time (no plan yet) an extra wait kernel is harmless
    reads = _refresh_halos()
    before = {r: l.epoch for r, l in base.links.items()}
    _write(base, todo)
    stepped = bool(base.links) and all(l.epoch != before[r] for r, l in base.links.items())
    if key is not None:
        if len(_assign_plans) > 256:
            _assign_plans.clear()
        _assign_plans[key] = (todo, reads, [node], base if stepped else None)


def _write(base, todo):
    base.generation = getattr(base, "generation", 0) + 1      # on every rank, whatever its share of the rows
    _write_local(base, todo)


def _write_local(base, todo):
    """Run the local assignments.  Halo bookkeeping: a block whose link did not step (the write
    was not the halo-pushing stencil kernel) leaves the neighbours' halo copies stale."""
    from . import engine
    links = base.links
    before = {r: l.epoch for r, l in links.items()}
    for tloc, vloc in todo:
        engine.assign(tloc, vloc)
    if links:
        stepped = [l.epoch != before[r] for r, l in links.items()]
        if any(stepped) and not all(stepped):
            raise RuntimeError("the stencil kernel ran on some blocks of a sharded array but not on "
                               "others (blocks too small for the tile?): the ranks' halos have diverged")
        for l in links.values():
            l.dirty = not stepped[0]


# ------------------------------------------------------------------------------ constructors
def _default_halo(shape, dtype, halo):
    if halo is not None:
        return int(halo)
    # (1-d arrays: no halo by default -- one element in front of the block would leave the owned rows
    # misaligned for 128-bit accesses; a 1-d stencil asks for it with halo=k)
    return 1 if (len(shape) == 2 and np.dtype(dtype).kind == "f") else 0


def shard(x, halo=None, mesh=None):
    """Split ``x`` (a host array, a DeviceArray or a lazy array holding the WHOLE array; under
    SPMD every rank passes the same one) along axis 0 over the mesh and return a sharded lazy
    array.  2-d float arrays get one halo row per side by default, so slice stencils need no
    change; ``halo=k`` for wider stencils, ``halo=0`` for none."""
    from . import delayarray as da
    mesh = mesh or current_mesh()
    if isinstance(x, da.DelayArray):
        arr = x._force()
        if getattr(arr, "_is_shard_view", False):
            return x
    else:
        arr = x if isinstance(x, DeviceArray) else np.asarray(x)
    if arr.ndim == 0:
        raise ValueError("cannot shard a 0-d array")
    g = arr.shape[0]
    return from_global_fn(lambda r0, r1: arr[r0:r1], arr.shape, arr.dtype, halo=halo, mesh=mesh,
                          bounds=[shard_bounds(g, mesh.world, r) for r in range(mesh.world)])


def from_global_fn(fn, gshape, dtype, halo=None, mesh=None, bounds=None):
    """Sharded array whose rows [r0, r1) are produced by ``fn(r0, r1)`` (array-like) on the rank
    that needs them -- including its halo rows, so the first stencil step needs no exchange."""
    from . import delayarray as da
    from . import engine
    mesh = mesh or current_mesh()
    gshape = tuple(int(s) for s in gshape)
    H = _default_halo(gshape, dtype, halo) if mesh.world > 1 else 0
    g = gshape[0]
    bounds = bounds or [shard_bounds(g, mesh.world, r) for r in range(mesh.world)]
    base = ShardedBase(mesh, gshape, dtype, bounds, H=H)
    for r in mesh.local:
        lo, hi = bounds[r]
        a, b = max(lo - H, 0), min(hi + H, g)
        blk = base.blocks[r]
        if b > a:
            rows = fn(a, b)
            if isinstance(rows, da.DelayArray):
                rows = rows._force()
            if isinstance(rows, DeviceArray):
                rows = _on_device(mesh, rows, blk.dev)
            elif blk.dev >= 0:
                rows = DeviceArray.from_host(np.ascontiguousarray(rows, dtype=base.dtype), blk.dev)   # upload to the block's own device
            else:
                rows = np.ascontiguousarray(rows)
            engine.assign(blk[a - (lo - H):b - (lo - H)], rows)
        for s0, s1 in ((0, a - (lo - H)), (b - (lo - H), blk.shape[0])):      # unused edge halos
            if s1 > s0:
                engine.assign(blk[s0:s1], 0)
    for l in base.links.values():
        l.dirty = False
    if base.links:
        mesh.barrier()
    return da.NPArray(ShardView(base))


def from_local(block, halo=0, mesh=None):
    """Sharded array from the row block(s) the ranks already hold: ``block`` is this rank's
    array (SPMD) or ``{rank: array}`` / a list (in-process mesh).  No data moves."""
    from . import delayarray as da
    mesh = mesh or current_mesh()
    if not isinstance(block, (dict, list, tuple)):
        block = {mesh.local[0]: block}
    elif not isinstance(block, dict):
        block = dict(enumerate(block))
    devs = {}
    for r in mesh.local:
        b = block[r]
        devs[r] = b._force() if isinstance(b, da.DelayArray) else b
    counts = mesh.allgather_obj({r: devs[r].shape[0] for r in mesh.local})
    bounds, at = [], 0
    for c in counts:
        bounds.append((at, at + c))
        at += c
    first = devs[mesh.local[0]]
    if halo:
        raise NotImplementedError("from_local with a halo: use shard() / from_global_fn(), which fill the halo rows")
    base = ShardedBase.__new__(ShardedBase)
    base.mesh, base.gshape, base.dtype = mesh, (at,) + tuple(first.shape[1:]), first.dtype
    base.bounds, base.H, base.links = bounds, 0, {}
    tail = 1
    for s in base.gshape[1:]:
        tail *= s
    base.pitch = tail * base.dtype.itemsize
    base.blocks = {r: (devs[r] if devs[r].is_contiguous else devs[r].copy()) for r in mesh.local}
    return da.NPArray(ShardView(base))
