"""Execution engine: plan -> generate -> compile (cubin cache) -> launch.

Replaces ``cuda.run`` of the reference (cuda.py:91-96: emit, make_kernel, kern(*inputs)).
The cubin cache has two levels: an in-process table keyed by the region's *structural* key
(program + layout class, never sizes / pointers / scalar values) and an on-disk cache keyed
by sha256(source + options); steady state is one dict lookup and one ctypes launch.
All launches go through libdrcuda (include/drcuda.h); a missing GPU raises.
"""
import ctypes as C
import hashlib
import os
import weakref

import numpy as np

from . import _lib, codegen, planner, ranges
from ._lib import check, lib
from .device import DeviceArray, DeviceBuffer, current_device

_HERE = os.path.dirname(os.path.abspath(__file__))
ARCH = "sm_100a"
NVRTC_OPTIONS = [f"--gpu-architecture={ARCH}", "--std=c++17", "--fmad=false",
                 "--generate-line-info", "-default-device"]
CACHE_DIR = os.environ.get("DR_CACHE_DIR", os.path.join(_HERE, "_cache"))
DUMP_SRC = os.environ.get("DR_DUMP_SRC")

with open(os.path.join(_HERE, "csrc", "math_tables.cuh")) as _f:      # generated coefficient tables
    PRELUDE = _f.read()
with open(os.path.join(_HERE, "csrc", "prelude.cuh")) as _f:
    PRELUDE += _f.read()

stats = {"compiled": 0, "disk_hits": 0, "mem_hits": 0, "launches": 0, "compile_ms": 0.0}


# --------------------------------------------------------------------------- kernels
class Kernel:
    __slots__ = ("name", "source", "cubin", "funcs", "meta", "occ")

    def __init__(self, name, source, cubin, meta):
        self.name, self.source, self.cubin, self.meta = name, source, cubin, meta
        self.funcs = {}           # device -> CUfunction handle
        self.occ = {}

    def func(self, dev):
        f = self.funcs.get(dev)
        if f is None:
            mod, fn = C.c_uint64(), C.c_uint64()
            check(lib.drc_module_load(dev, self.cubin, len(self.cubin), C.byref(mod)))
            check(lib.drc_module_get_function(dev, mod, self.name.encode(), C.byref(fn)))
            f = self.funcs[dev] = fn.value
        return f

    def blocks_per_sm(self, dev, threads, smem=0):
        k = (dev, threads, smem)
        v = self.occ.get(k)
        if v is None:
            out = C.c_int()
            check(lib.drc_occupancy(dev, self.func(dev), threads, smem, C.byref(out)))
            v = self.occ[k] = max(out.value, 1)
        return v


_kernels = {}


def kernel_name(key):
    """dr_<family>_<structural hash>: the family (flat, nd, rows, cols, stencil, mm_skinny) is
    readable in profiler launch lists."""
    family = key[0] if isinstance(key, tuple) and isinstance(key[0], str) else "k"
    return f"dr_{family}_" + hashlib.sha256(repr(key).encode()).hexdigest()[:16]


_NVRTC_TAG = []


def _nvrtc_tag():
    """'nvrtc-12.9': part of the cubin cache key (code generation differs between releases)."""
    if not _NVRTC_TAG:
        try:
            ma, mi, _ = _lib.nvrtc_version()
            _NVRTC_TAG.append(f"nvrtc-{ma}.{mi}")
        except Exception:
            _NVRTC_TAG.append("nvrtc-?")
    return _NVRTC_TAG[0]


def compile_source(name, body_source):
    """Source text -> cubin through the on-disk cache (works without a GPU)."""
    import time
    source = PRELUDE + "\n" + body_source
    digest = hashlib.sha256((source + "\0" + " ".join(NVRTC_OPTIONS) + "\0" + _nvrtc_tag()).encode()).hexdigest()
    path = os.path.join(CACHE_DIR, f"{name}-{digest[:16]}.cubin")
    if os.path.exists(path):
        if os.environ.get("DR_CACHE_TOUCH"):
            os.utime(path)                      # tools/prune_cache.sh: mark the entries a run needs
        with open(path, "rb") as f:
            stats["disk_hits"] += 1
            return source, f.read()
    t0 = time.perf_counter()
    cubin, _log = _lib.compile_cubin(source, name + ".cu", NVRTC_OPTIONS)
    stats["compile_ms"] += (time.perf_counter() - t0) * 1e3
    stats["compiled"] += 1
    os.makedirs(CACHE_DIR, exist_ok=True)
    tmp = f"{path}.{os.getpid()}.tmp"
    with open(tmp, "wb") as f:
        f.write(cubin)
    os.replace(tmp, path)
    if DUMP_SRC:
        os.makedirs(DUMP_SRC, exist_ok=True)
        with open(os.path.join(DUMP_SRC, name + ".cu"), "w") as f:
            f.write(source)
    return source, cubin


def get_kernel(key, generate, meta=None):
    """``generate(name) -> CUDA source`` is only called on a structural-key miss."""
    k = _kernels.get(key)
    if k is not None:
        stats["mem_hits"] += 1
        return k
    name = kernel_name(key)
    source, cubin = compile_source(name, generate(name))
    k = _kernels[key] = Kernel(name, source, cubin, meta or {})
    return k


# --------------------------------------------------------------------------- device state
class _DevState:
    def __init__(self, dev):
        sm, major, minor, l2, smem = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        total = C.c_size_t()
        check(lib.drc_device_attr(dev, C.byref(sm), C.byref(major), C.byref(minor),
                                  C.byref(total), C.byref(l2), C.byref(smem)))
        self.sm_count, self.cc, self.l2_bytes = sm.value, (major.value, minor.value), l2.value
        self.max_smem = smem.value
        self.total_mem = total.value
        if self.cc[0] < 10:
            raise _lib.DrcError(f"device {dev} is sm_{major.value}{minor.value}; this engine "
                                f"targets {ARCH} only")
        self.scratch = DeviceBuffer(MAX_PARTIALS * 16 + 256, dev)
        check(lib.drc_memset_async(dev, 0, self.scratch.ptr, 0, self.scratch.nbytes))

    @property
    def partials_ptr(self):
        return self.scratch.ptr + 256

    @property
    def counter_ptr(self):
        return self.scratch.ptr


MAX_PARTIALS = 8192
_dev_state = {}
dry_log = []          # (kernel, grid, block) recorded instead of launched in dry mode


class _DryState:
    """Stand-in device description for compile-only planning (no GPU in the build container)."""
    sm_count, cc, l2_bytes, max_smem, total_mem = 148, (10, 0), 126 << 20, 232448, 180 << 30
    partials_ptr, counter_ptr = 0x7E0000000100, 0x7E0000000000


class dry_run:
    """Context manager: plan, generate and NVRTC-compile regions without touching a device.
    Arrays get placeholder addresses; launches are recorded in ``dry_log``.  Used by the CPU
    test-suite and by __graft_entry__.build() to pre-compile the workload kernels."""

    def __enter__(self):
        from . import device
        self._prev = device._current["dry"]
        device._current["dry"] = True
        return dry_log

    def __exit__(self, *exc):
        from . import device
        device._current["dry"] = self._prev


def is_dry():
    from . import device
    return device._current["dry"]


def dev_state(dev):
    if dev < 0:
        return _DryState
    st = _dev_state.get(dev)
    if st is None:
        _lib.init()
        st = _dev_state[dev] = _DevState(dev)
    return st


# --------------------------------------------------------------------------- launch
_PACK_ALIGN = 8


_SCALAR_FMT = {"f4": ("f", 4), "f8": ("d", 8), "i4": ("i", 4), "i8": ("q", 8), "u4": ("I", 4),
               "u8": ("Q", 8), "i2": ("h", 2), "u2": ("H", 2), "i1": ("b", 1), "u1": ("B", 1),
               "b1": ("?", 1)}
_arg_layouts = {}            # format string -> (struct.Struct, ctypes offsets array, count)


class Args:
    """Kernel arguments: a struct format (explicit padding, natural alignment per argument --
    the CUDA parameter layout) plus the values; packed once at launch.  The compiled Struct and
    the ctypes offset table are cached per format string, i.e. per kernel signature."""

    __slots__ = ("fmt", "vals", "size", "offsets")

    def __init__(self):
        self.fmt, self.vals, self.size, self.offsets = [], [], 0, []

    def _put(self, code, size, align, value):
        pad = (-self.size) % align
        if pad:
            self.fmt.append(f"{pad}x")
            self.size += pad
        self.offsets.append(self.size)
        self.fmt.append(code)
        self.vals.append(value)
        self.size += size

    def raw(self, data, align):
        self._put(f"{len(data)}s", len(data), align, bytes(data))

    def i64(self, v):
        self._put("q", 8, 8, int(v))

    def ptr(self, v):
        self._put("Q", 8, 8, int(v))

    def f64(self, v):
        self._put("d", 8, 8, float(v))

    def scalar(self, value, dtype):
        dt = np.dtype(dtype)
        hit = _SCALAR_FMT.get(dt.str[1:])
        if hit is not None and dt.kind != "c":
            code, size = hit
            if dt.kind == "f":
                # route through the dtype so that float32 rounding / inf / nan match NumPy's
                value = float(dt.type(value))
            elif dt.kind == "b":
                value = bool(value)
            else:
                value = int(dt.type(value))
            self._put(code, size, size, value)
        else:
            data = np.asarray(value, dtype=dt).tobytes()
            self.raw(data, max(min(len(data), 8), 1))

    def pack(self):
        key = "".join(self.fmt)
        lay = _arg_layouts.get(key)
        if lay is None:
            import struct
            lay = _arg_layouts[key] = (struct.Struct("<" + key),
                                       (C.c_uint32 * len(self.offsets))(*self.offsets),
                                       len(self.offsets))
        return lay[0].pack(*self.vals), lay[1], lay[2]


_last_kernel = [None]


def last_kernel_name():
    """Name (dr_<family>_<structural hash>) of the kernel launched last; bench.py ties its ncu
    traffic figure to it."""
    return _last_kernel[0]


def launch(kernel, dev, grid, block, args, smem=0, stream=0, cluster=1):
    _last_kernel[0] = kernel.name
    if dev < 0:
        dry_log.append((kernel, grid, block))
        return
    blob, offs, n = args.pack()
    gx, gy, gz = (tuple(grid) + (1, 1))[:3] if not isinstance(grid, int) else (grid, 1, 1)
    if gz != 1:                 # generated code uses gridDim.z - 1 as a zero ptxas cannot fold (codegen.emit_expr)
        raise ValueError("kernels are launched with gridDim.z == 1")
    bx, by, bz = (tuple(block) + (1, 1))[:3] if not isinstance(block, int) else (block, 1, 1)
    check(lib.drc_launch_packed(dev, stream, kernel.func(dev), gx, gy, gz, bx, by, bz, smem,
                                cluster, blob, offs, n))
    stats["launches"] += 1


def _grid_for(kernel, dev, threads, work_items, smem=0):
    st = dev_state(dev)
    cap = st.sm_count * (kernel.blocks_per_sm(dev, threads, smem) if dev >= 0 else 8)
    need = max(1, -(-work_items // threads))
    return min(need, cap)


# --------------------------------------------------------------------------- regions
def _stamp_of(prog):
    return [(weakref.ref(b), b.version) for b in prog.leaf_bufs]


def _geo_blob(total, shape, strides_per_operand):
    vals = [total] + list(shape)
    for st in strides_per_operand:
        vals += list(st)
    return np.asarray(vals, dtype=np.int64).tobytes()


def run_program(prog, outs, reduce=None, inplace=False):
    """Launch the fused kernel of ``prog``.  ``outs``: DeviceArrays to write (elementwise), or
    [] with ``reduce`` = (op, acc_dt, res_dt, count, result DeviceArray) for a full reduction."""
    dev = outs[0].dev if outs else (prog.arrays[0].dev if prog.arrays else current_device())
    st = dev_state(dev)
    lay = planner.resolve_layout(prog, outs, inner_vectors=reduce is None and not os.environ.get("DR_NO_NDVEC"))
    if lay.total == 0:
        return
    out_dts = tuple(o.dtype for o in outs)
    red_key = None if reduce is None else (reduce[0], reduce[1].str, reduce[2].str)
    moved = sum(a.dtype.itemsize for a, c in zip(prog.arrays, lay.in_class) if c != "b") \
        + sum(o.dtype.itemsize for o in outs)
    stream_hint = (not inplace) and lay.total * max(moved, 1) > 2 * st.l2_bytes
    if lay.family == "flat":
        # float32 scalars are classified by magnitude ('n' = 2^-24 <= |s| < 2^24): the interval
        # analysis that places the range guards (ranges.py) may rely on the class, so it is
        # part of the structural key; values themselves stay kernel arguments
        scl = ranges.scalar_classes(prog) if codegen.has_lane_fast(prog) else None
        key = ("flat", prog.key(), lay.in_class, tuple(d.str for d in out_dts), lay.vec_ok,
               stream_hint, red_key, inplace, scl)
        if scl is not None and codegen.ERF3 and codegen.uses_erf_table(prog):
            key += ("erf3",)               # the table generation is part of the kernel's identity
        gen_reduce = None if reduce is None else (reduce[0], reduce[1], reduce[2], None)
        meta = {}
        kern = get_kernel(key, lambda name: codegen.gen_flat(
            name, prog, lay.in_class, out_dts, lay.vec_ok, stream_hint, gen_reduce, meta=meta,
            sclasses=scl), meta)
        smem = kern.meta.get("smem", 0)
        threads = kern.meta.get("threads", 256)
        a = Args()
        a.i64(lay.total)
        head = ("i64", lay.total)
        for arr in prog.arrays:
            a.ptr(arr.ptr)
        for val, dt in prog.scalars:
            a.scalar(val, dt)
        for o in outs:
            a.ptr(o.ptr)
        widest = max([x.dtype.itemsize for x, c in zip(prog.arrays, lay.in_class) if c == "c"]
                     + [d.itemsize for d in out_dts] + [1])
        vec = max(1, 16 // widest) if lay.vec_ok else 1
        if smem > 40 * 1024 and dev >= 0 and dev not in kern.meta.setdefault("smem_set_devs", set()):
            check(lib.drc_func_set_max_dynamic_smem(dev, kern.func(dev), smem))
            kern.meta["smem_set_devs"].add(dev)
        grid = _grid_for(kern, dev, threads, -(-lay.total // vec), smem)
    elif reduce is None and not inplace and not os.environ.get("DR_NO_TILE") \
            and planner.tile_classes(prog, outs, lay):
        # a transposed operand among row-major ones (`X.T + X`): shared-memory tiles, every
        # global access coalesced
        cls, T, W = planner.tile_classes(prog, outs, lay)
        smem, threads, scl = 0, 256, None
        key = ("tile", prog.key(), cls, tuple(d.str for d in out_dts), T, W)
        kern = get_kernel(key, lambda name: codegen.gen_tile(name, prog, cls, out_dts, T=T, W=W))
        R, C = lay.shape
        tiles_c = -(-C // T)
        ntiles = tiles_c * -(-R // T)
        vals = [R, C, tiles_c, ntiles]
        for stv in list(lay.in_strides) + list(lay.out_strides):
            vals += list(stv)
        head = ("raw", np.asarray(vals, dtype=np.int64).tobytes())
        a = Args()
        a.raw(head[1], 8)
        for arr in prog.arrays:
            a.ptr(arr.ptr)
        for val, dt in prog.scalars:
            a.scalar(val, dt)
        for o in outs:
            a.ptr(o.ptr)
        grid = _grid_for(kern, dev, 256, ntiles * 256)
    else:
        smem = 0
        threads = 256
        wide = lay.total >= (1 << 32) or any(
            abs(s) * n >= (1 << 62) for stv in lay.in_strides for s, n in zip(stv, lay.shape))
        vec = int(lay.vec_ok) if (lay.vec_ok and reduce is None) else 0
        scl = ranges.scalar_classes(prog) if (vec == 4 and codegen.has_lane_fast(prog)) else None
        key = ("nd", prog.key(), len(lay.shape), lay.in_class, tuple(d.str for d in out_dts),
               red_key, wide, vec, scl)
        gen_reduce = None if reduce is None else (reduce[0], reduce[1], reduce[2], None)
        kern = get_kernel(key, lambda name: codegen.gen_nd(
            name, prog, len(lay.shape), lay.in_class, out_dts, gen_reduce, wide_index=wide, vec=vec,
            sclasses=scl))
        a = Args()
        operands = list(lay.in_strides) + (list(lay.out_strides) if reduce is None else [])
        if not operands:
            operands = [(0,) * len(lay.shape)]
        head = ("raw", _geo_blob(lay.total, lay.shape, operands))
        a.raw(head[1], 8)
        for arr in prog.arrays:
            a.ptr(arr.ptr)
        for val, dt in prog.scalars:
            a.scalar(val, dt)
        for o in outs:
            a.ptr(o.ptr)
        grid = _grid_for(kern, dev, 256, lay.total)
    if reduce is not None:
        grid = min(grid, MAX_PARTIALS)
        a.ptr(st.partials_ptr)
        a.ptr(st.counter_ptr)
        a.ptr(reduce[4].ptr)
        a.f64(reduce[3])
    launch(kern, dev, grid, threads, a, smem=smem)
    return kern, grid, threads, smem, head, scl


# --------------------------------------------------------------------------- prepared launches
# Plan cache: roots' structural signatures (delayarray._plan_info) -> everything a launch needs
# except pointers and scalar values.  Only plain elementwise regions with fresh outputs.
_plans = {}
_PLAN_CACHE = os.environ.get("DR_PLAN_CACHE", "1") != "0"
_PLAN_LIMIT = 1024


class _Plan:
    __slots__ = ("kern", "grid", "threads", "smem", "head", "arr_idx", "sc_idx", "sc_dt", "leaf_sigs",
                 "shape", "out_dts", "dev", "scl", "layout")


def _plan_key(nodes):
    """(key, operand nodes) of the roots, or (None, None) when any root is not eligible."""
    if len(nodes) == 1:
        d = nodes[0].__dict__
        sig = d.get("_psig")
        if sig is None:
            return None, None
        return (sig, nodes[0].shape, nodes[0].dtype), d["_pops"]
    from .delayarray import _merge_ops
    parts, ops = [], ()
    for n in nodes:
        d = n.__dict__
        sig = d.get("_psig")
        if sig is None:
            return None, None
        if not ops:
            ops = d["_pops"]
            parts.append((sig, n.shape, n.dtype))
        else:
            ops, idx = _merge_ops(ops, d["_pops"])
            parts.append((sig, n.shape, n.dtype, idx))
    return tuple(parts), ops


def _plan_launch(plan, ops):
    """Launch a prepared plan against the current operands; returns (outs, stamp) or None when
    an operand no longer has the layout (or scalar class) the plan was made for."""
    from .delayarray import _leaf_sig
    arrays = []
    for i, want in zip(plan.arr_idx, plan.leaf_sigs):
        leaf = ops[i]
        if _leaf_sig(leaf) != want:
            return None
        arrays.append(leaf._force())
    if plan.scl is not None:
        for j, (i, dt) in enumerate(zip(plan.sc_idx, plan.sc_dt)):
            if ranges.scalar_class(ops[i].val, dt) != plan.scl[j]:
                return None
    dev = plan.dev
    outs = [DeviceArray.empty(plan.shape, dt, dev if dev >= 0 else None) for dt in plan.out_dts]
    lay = plan.layout
    if lay is None:
        a = Args()
        if plan.head[0] == "i64":
            a.i64(plan.head[1])
        else:
            a.raw(plan.head[1], 8)
        for arr in arrays:
            a.ptr(arr.ptr)
        for i, dt in zip(plan.sc_idx, plan.sc_dt):
            a.scalar(ops[i].val, dt)
        for o in outs:
            a.ptr(o.ptr)
        launch(plan.kern, dev, plan.grid, plan.threads, a, smem=plan.smem)
        if all(np.dtype(dt).kind == "f" for dt in plan.sc_dt) and isinstance(plan.grid, int):
            a.pack()
            plan.layout = _arg_layouts["".join(a.fmt)]      # the argument layout of a plan is fixed
    else:
        vals = [plan.head[1]]
        for arr in arrays:
            vals.append(arr.buf.ptr + arr.offset)
        for i, dt in zip(plan.sc_idx, plan.sc_dt):
            vals.append(float(dt.type(ops[i].val)))
        for o in outs:
            vals.append(o.buf.ptr)
        _last_kernel[0] = plan.kern.name
        if dev >= 0:
            check(lib.drc_launch_packed(dev, 0, plan.kern.func(dev), plan.grid, 1, 1, plan.threads, 1, 1,
                                        plan.smem, 1, lay[0].pack(*vals), lay[1], lay[2]))
            stats["launches"] += 1
        else:
            dry_log.append((plan.kern, plan.grid, plan.threads))
    seen, stamp = set(), []
    for arr in arrays:
        b = arr.buf
        if id(b) not in seen:
            seen.add(id(b))
            stamp.append((weakref.ref(b), b.version))
    stats["plan_hits"] = stats.get("plan_hits", 0) + 1
    return outs, stamp


def _plan_record(key, ops, prog, outs, rec):
    """Remember the launch `rec` of `prog` under `key` if every kernel operand maps to one of
    the roots' operand nodes."""
    from .delayarray import _leaf_sig
    kern, grid, threads, smem, head, scl = rec
    arr_idx, leaf_sigs = [], []
    for arr in prog.arrays:
        for i, o in enumerate(ops):
            if o.kind == "leaf" and o._force() is arr:
                sig = _leaf_sig(o)
                if sig is None:
                    return
                arr_idx.append(i)
                leaf_sigs.append(sig)
                break
        else:
            return
    sc_idx = []
    for node in prog.scalar_nodes:
        for i, o in enumerate(ops):
            if o is node:
                sc_idx.append(i)
                break
        else:
            return
    if len(_plans) >= _PLAN_LIMIT:
        _plans.clear()
    p = _Plan()
    p.kern, p.grid, p.threads, p.smem, p.head, p.scl = kern, grid, threads, smem, head, scl
    p.arr_idx, p.leaf_sigs, p.sc_idx = tuple(arr_idx), tuple(leaf_sigs), tuple(sc_idx)
    p.sc_dt = tuple(np.dtype(dt) for _, dt in prog.scalars)
    p.shape, p.out_dts, p.dev = tuple(outs[0].shape), tuple(o.dtype for o in outs), outs[0].dev
    p.layout = None
    _plans[key] = p


def _reduce_plan_key(src, op, res_dt, post):
    from .delayarray import _leaf_sig
    if src.kind == "leaf":
        sig = _leaf_sig(src)
        if sig is None:
            return None, None
        return ("reduce", "leaf", sig, op, res_dt.str, post), (src,)
    d = src.__dict__
    sig = d.get("_psig")
    if src.kind != "ewise" or sig is None:
        return None, None
    return ("reduce", sig, src.shape, src.dtype, op, res_dt.str, post), d["_pops"]


def _reduce_plan_launch(plan, ops, shape, res_dt, post):
    """_plan_launch for a fused full reduction: operands, then the partials / ticket / result
    pointers and the post scale (run_program's argument order)."""
    from .delayarray import _leaf_sig
    arrays = []
    for i, want in zip(plan.arr_idx, plan.leaf_sigs):
        leaf = ops[i]
        if _leaf_sig(leaf) != want:
            return None
        arrays.append(leaf._force())
    if plan.scl is not None:
        for j, (i, dt) in enumerate(zip(plan.sc_idx, plan.sc_dt)):
            if ranges.scalar_class(ops[i].val, dt) != plan.scl[j]:
                return None
    dev = plan.dev
    st = dev_state(dev)
    result = DeviceArray.empty(shape, res_dt, dev if dev >= 0 else None)
    a = Args()
    if plan.head[0] == "i64":
        a.i64(plan.head[1])
    else:
        a.raw(plan.head[1], 8)
    for arr in arrays:
        a.ptr(arr.ptr)
    for i, dt in zip(plan.sc_idx, plan.sc_dt):
        a.scalar(ops[i].val, dt)
    a.ptr(st.partials_ptr)
    a.ptr(st.counter_ptr)
    a.ptr(result.ptr)
    a.f64(post)
    launch(plan.kern, dev, plan.grid, plan.threads, a, smem=plan.smem)
    seen, stamp = set(), []
    for arr in arrays:
        b = arr.buf
        if id(b) not in seen:
            seen.add(id(b))
            stamp.append((weakref.ref(b), b.version))
    stats["plan_hits"] = stats.get("plan_hits", 0) + 1
    return result, stamp


def evaluate_nodes(nodes, outs=None, inplace=False):
    """Fuse ``nodes`` (same iteration shape) into one kernel; returns the output arrays."""
    key = ops = None
    if outs is None and not inplace and _PLAN_CACHE:
        key, ops = _plan_key(nodes)
        if key is not None:
            plan = _plans.get(key)
            if plan is not None:
                done = _plan_launch(plan, ops)
                if done is not None:
                    return done
    prog = planner.build_program(nodes)
    fresh = outs is None
    if fresh:
        dev = prog.arrays[0].dev if prog.arrays else current_device()
        outs = [DeviceArray.empty(prog.shape, n.dtype, dev) for n in nodes]
    else:
        prog.shape = tuple(outs[0].shape)
    rec = run_program(prog, outs, inplace=inplace)
    if key is not None and rec is not None and fresh and prog.arrays:
        _plan_record(key, ops, prog, outs, rec)
    stamp = _stamp_of(prog)
    return outs, stamp


def run(node):
    """Backend protocol entry: evaluate one node, return its DeviceArray  (cuda.py:91-96)."""
    kind = node.kind
    if node.__dict__.get("_mesh") is not None:
        from . import sharding
        return sharding.run(node)            # a row-sharded operand below: localise per block
    if kind == "leaf":
        return node._force()
    if kind == "scalar":
        raise TypeError("cannot evaluate a bare Scalar")
    if kind == "ewise":
        outs, stamp = evaluate_nodes([node])
        node._stamp = stamp
        return outs[0]
    if kind == "reduce":
        return _run_reduce(node)
    if kind == "matmul":
        return _run_contraction(node)
    raise NotImplementedError(kind)


def run_many(nodes):
    """Co-evaluate: elementwise nodes sharing a shape go into ONE multi-output kernel."""
    groups = {}
    sharded = [n for n in nodes if n.__dict__.get("_mesh") is not None]
    if sharded:
        from . import sharding
        sharding.run_many(sharded)           # co-evaluated per row block, one kernel per block and shape
    for n in nodes:
        if n.__dict__.get("_mesh") is not None:
            continue
        elif n.kind == "ewise" and n.__dict__.get("array") is None:
            groups.setdefault(tuple(n.shape), [])
            if all(n is not m for m in groups[tuple(n.shape)]):
                groups[tuple(n.shape)].append(n)
        else:
            n._force()
    for group in groups.values():
        outs, stamp = evaluate_nodes(group)
        for n, o in zip(group, outs):
            n.array, n._stamp = o, stamp


# --------------------------------------------------------------------------- reductions
def _group_strides(arr_strides, shape, lo, hi):
    """Collapse dims [lo, hi) of one operand to a single stride, or None if impossible."""
    dims = [d for d in range(lo, hi) if shape[d] != 1]
    if not dims:
        return 0
    for a, b in zip(dims, dims[1:]):
        if arr_strides[a] != arr_strides[b] * shape[b]:
            return None
    return arr_strides[dims[-1]]


def _run_reduce(node):
    from .delayarray import MMEx, NPArray, ReduceEx
    child = node.children[0]
    if node.op == "sum" and node.axes == (1,) and child.ndim == 2 and not node.keepdims:
        # a pending A @ B over the same lazy producer A: one pass computes both (the row sum
        # rides along as a virtual column of ones), see _try_mm_skinny
        for cons in list(getattr(child, "_consumers", ())):
            if isinstance(cons, MMEx) and cons.arg1 is child and cons.__dict__.get("array") is None \
                    and cons.shape[1] <= 6 and child.kind == "ewise":
                cons._force()
                if node.__dict__.get("array") is not None:
                    return node.array
    op = node.op
    res_dt = node.dtype
    nd = child.ndim
    axes = node.axes
    count = 1
    for ax in axes:
        count *= child.shape[ax]
    post = float(count) if node.post == "mean" else 1.0
    src = child
    if src.dtype != res_dt:
        from .delayarray import as_dtype
        src = as_dtype(child, res_dt)       # e.g. bool/int8 sums accumulate as int64
    acc_dt = codegen.acc_dtype(op, res_dt)
    full = len(axes) == nd
    # non-contiguous axis groups: peel the last contiguous run first
    if not full and axes and any(b != a + 1 for a, b in zip(axes, axes[1:])):
        run_start = len(axes) - 1
        while run_start > 0 and axes[run_start - 1] == axes[run_start] - 1:
            run_start -= 1
        inner = ReduceEx(node.func, child, tuple(axes[run_start:]), True)
        outer = ReduceEx(node.func, inner, tuple(axes[:run_start]), True)
        res = outer._force()
        if node.post == "mean":
            from .delayarray import as_dtype
            res = as_dtype(NPArray(res) / float(count), res_dt)._force()
        return res.reshape(node.shape)
    rkey = rops = None
    if _PLAN_CACHE and full and axes and child.size:
        # prepared launches for full reductions: the producer's structural signature (or the
        # leaf's layout) + the reduction; a hit packs pointers and launches, nothing else
        rkey, rops = _reduce_plan_key(src, op, res_dt, node.post)
        plan = _plans.get(rkey) if rkey is not None else None
        if plan is not None:
            done = _reduce_plan_launch(plan, rops, node.shape, res_dt, post)
            if done is not None:
                node._stamp = done[1]
                return done[0]
    prog = planner.build_program([src])
    prog.shape = tuple(child.shape)
    dev = prog.arrays[0].dev if prog.arrays else current_device()
    result = DeviceArray.empty(node.shape, res_dt, dev)
    node._stamp = _stamp_of(prog)
    if child.size == 0:
        # NumPy's identities: sum 0, prod 1, mean nan (0/0); max/min of nothing is an error
        if count == 0 and op in ("max", "min"):
            raise ValueError(f"zero-size array to reduction operation {op}imum which has no identity")
        if result.size:
            result.fill(float("nan") if node.post == "mean" and res_dt.kind == "f"
                        else 1 if op == "prod" else 0)
        return result
    if full or not axes:
        if not axes:        # reduction over nothing: a copy
            outs, _ = evaluate_nodes([src], [result])
            return result
        rec = run_program(prog, [], reduce=(op, acc_dt, res_dt, post, result))
        if rkey is not None and rec is not None and prog.arrays:
            _plan_record(rkey, rops, prog, [result], rec)
        return result
    lo, hi = axes[0], axes[-1] + 1
    shape = child.shape
    outer = int(np.prod(shape[:lo], dtype=np.int64))
    red = int(np.prod(shape[lo:hi], dtype=np.int64))
    inner = int(np.prod(shape[hi:], dtype=np.int64))
    triples = []
    for arr in prog.arrays:
        bst = planner.broadcast_strides(arr, shape)
        t = (_group_strides(bst, shape, 0, lo), _group_strides(bst, shape, lo, hi),
             _group_strides(bst, shape, hi, nd))
        if None in t:
            triples = None
            break
        triples.append(t)
    if triples is None:
        # an operand whose dims do not collapse into (outer, reduced, inner): materialise the
        # producer contiguously once, then reduce that
        dense = NPArray(evaluate_nodes([src])[0][0])
        return _run_reduce_dense(node, dense, outer, red, inner, op, acc_dt, res_dt, post, result)
    _launch_axis_reduce(prog, triples, outer, red, inner, op, acc_dt, res_dt, post, result, dev)
    return result


def _run_reduce_dense(node, dense, outer, red, inner, op, acc_dt, res_dt, post, result):
    prog = planner.build_program([dense])
    item = dense.dtype.itemsize
    triples = [(red * inner * item, inner * item, item)]
    _launch_axis_reduce(prog, triples, outer, red, inner, op, acc_dt, res_dt, post, result,
                        result.dev)
    return result


builtins_all, builtins_any = all, any


def _axis_classes(prog, triples, along, other, extent):
    """Operand classes for the rows / cols kernels and the vector width they allow.
    along = index in the triple of the axis the threads walk contiguously, other = the indices
    of the remaining strides (must keep 16-byte alignment for vector loads)."""
    width = max([a.dtype.itemsize for a in prog.arrays] + [1])
    V = max(1, 16 // width)
    cls = []
    for arr, t in zip(prog.arrays, triples):
        if t == (0, 0, 0):
            cls.append("b")
        elif t[along] == 0:
            cls.append("i")
        elif t[along] == arr.dtype.itemsize and arr.dtype.itemsize * V == 16 and arr.ptr % 16 == 0 \
                and builtins_all(t[k] % 16 == 0 for k in other):
            cls.append("v")
        else:
            cls.append("s")
    if V > 1 and (extent % V != 0 or "v" not in cls or os.environ.get("DR_NO_AXISVEC")):
        V = 1
    if V == 1:
        cls = ["s" if c == "v" else c for c in cls]
    return tuple(cls), V


def _launch_axis_reduce(prog, triples, outer, red, inner, op, acc_dt, res_dt, post, result, dev):
    st = dev_state(dev)
    red_spec = (op, acc_dt, res_dt, None)
    a = Args()
    n_ops = max(len(triples), 1)
    pad = [(0, 0, 0)] * (n_ops - len(triples))
    tr = list(triples) + pad
    def _bytes_where(axis):          # footprint of the operands that are contiguous along `axis`
        return sum(arr.dtype.itemsize * (outer if t[0] else 1) * (red if t[1] else 1)
                   for arr, t in zip(prog.arrays, triples) if t[axis] == arr.dtype.itemsize)
    if inner == 1 and outer > 1 and triples and _bytes_where(0) > _bytes_where(1):
        # the operands are contiguous along the KEPT axis (transposed views, `v @ X`): walk it
        # with the column kernel -- (1, red, outer) with the stride roles swapped -- instead of
        # reading 4-byte elements a row pitch apart
        swapped = [(0, t[1], t[0]) for t in triples]
        return _launch_axis_reduce(prog, swapped, 1, red, outer, op, acc_dt, res_dt, post, result, dev)
    if inner == 1:
        in_class, V = _axis_classes(prog, triples, 1, (0,), red)
        mode = "block" if red >= 2048 or outer < st.sm_count * 8 else "warp"
        key = ("rows", prog.key(), in_class, op, acc_dt.str, res_dt.str, mode, V)
        kern = get_kernel(key, lambda name: codegen.gen_rows(name, prog, in_class, red_spec, mode, V=V))
        geo = [outer, red] + [t[0] for t in tr] + [t[1] for t in tr] + [res_dt.itemsize]
        a.raw(np.asarray(geo, dtype=np.int64).tobytes(), 8)
        for arr in prog.arrays:
            a.ptr(arr.ptr)
        for val, dt in prog.scalars:
            a.scalar(val, dt)
        a.ptr(result.ptr)
        a.f64(post)
        cap = st.sm_count * (kern.blocks_per_sm(dev, 256) if dev >= 0 else 8)
        grid = min(outer, cap) if mode == "block" else min(max(1, -(-outer * 32 // 256)), cap)
        launch(kern, dev, grid, 256, a)
        return
    in_class, V = _axis_classes(prog, triples, 2, (0, 1), inner)
    # enough CTAs to fill the machine: split the reduced axis when (outer x inner) alone is small
    blocks_x = max(1, -(-(outer * (inner // V)) // 256))
    want = st.sm_count * 8
    splits = 1
    if blocks_x < want and red >= 64:
        splits = min(-(-want // blocks_x), red // 16, 1024)
    chunk = -(-red // splits)
    splits = -(-red // chunk)
    partial = splits > 1
    key = ("cols", prog.key(), in_class, op, acc_dt.str, res_dt.str, V, partial)
    kern = get_kernel(key, lambda name: codegen.gen_cols(name, prog, in_class, red_spec, V=V,
                                                         partial=partial))
    geo = [outer, red, inner] + [t[0] for t in tr] + [t[1] for t in tr] + [t[2] for t in tr] + [chunk]
    a.raw(np.asarray(geo, dtype=np.int64).tobytes(), 8)
    for arr in prog.arrays:
        a.ptr(arr.ptr)
    for val, dt in prog.scalars:
        a.scalar(val, dt)
    target = DeviceArray.empty((outer, splits, inner), acc_dt, dev if dev >= 0 else None) if partial else result
    a.ptr(target.ptr)
    a.f64(post)
    gx = min(blocks_x, st.sm_count * 16)
    launch(kern, dev, (gx, splits, 1), 256, a)
    if partial:
        from .delayarray import NPArray
        p2 = planner.build_program([NPArray(target)])
        item = acc_dt.itemsize
        _launch_axis_reduce(p2, [(splits * inner * item, inner * item, item)], outer, splits, inner,
                            op, acc_dt, res_dt, post, result, dev)


# --------------------------------------------------------------------------- contractions
def _run_contraction(node):
    from .delayarray import BinaryNumpyEx, DotEx, MVEx, ReduceEx
    a, b = node.arg1, node.arg2
    if isinstance(node, DotEx):
        red = ReduceEx(np.add, BinaryNumpyEx(np.multiply, a, b), None, False)
        out = red._force()
        node._stamp = red._stamp
        return out if out.dtype == node.dtype else out.astype(node.dtype)
    if isinstance(node, MVEx):
        red = ReduceEx(np.add, BinaryNumpyEx(np.multiply, a, b), 1, False)
        out = red._force()
        node._stamp = red._stamp
        return out if out.dtype == node.dtype else out.astype(node.dtype)
    return _run_matmul(node)


def _run_matmul(node):
    """A(M,K) @ B(K,N) with both operands' elementwise producers fused: iteration space
    (M, K, N), reduced over K.  One thread per output element, B coalesced along N, A
    broadcast within the warp.  (Round-1 functional path; the tcgen05 tile kernel for large
    dense operands is the next row)"""
    a, b = node.arg1, node.arg2
    m, k = a.shape
    n = b.shape[1]
    res_dt = node.dtype
    from .delayarray import as_dtype
    pa = planner.build_program([as_dtype(a, res_dt)])
    skinny = _try_mm_skinny(node, pa, m, k, n, res_dt)      # first: it needs nothing of what follows
    if skinny is not None:
        return skinny
    pb = planner.build_program([as_dtype(b, res_dt)])
    prog = planner.Program()
    triples = []
    remap = {}
    for tag, p, dims in (("A", pa, (m, k)), ("B", pb, (k, n))):
        base_a, base_s, base_t = len(prog.arrays), len(prog.scalars), len(prog.instrs)
        for arr in p.arrays:
            bst = planner.broadcast_strides(arr, dims)
            triples.append((bst[0], bst[1], 0) if tag == "A" else (0, bst[0], bst[1]))
            prog.arrays.append(arr)
        prog.scalars += p.scalars

        def mv(r, base_a=base_a, base_s=base_s, base_t=base_t):
            return (r[0], r[1] + {"a": base_a, "s": base_s, "t": base_t}[r[0]])
        for op, loop, out, args in p.instrs:
            prog.instrs.append((op, loop, out, tuple(mv(r) for r in args)))
        for r, dt in p.dtypes.items():
            prog.dtypes[mv(r)] = dt
        remap[tag] = mv(p.roots[0])
        for buf in p.leaf_bufs:
            if all(buf is not x for x in prog.leaf_bufs):
                prog.leaf_bufs.append(buf)
    if res_dt == np.float32 and min(m, n) >= 8 and max(m, n) >= 128 and k >= 64 \
            and not os.environ.get("DR_NO_TCGEN05"):
        from . import gemm
        node._stamp = _stamp_of(pa) + _stamp_of(pb)
        return gemm.matmul_tf32x3(a, b)
    if res_dt.kind in "fiu" and m * n >= 4096 and k >= 32 and m * n * k >= (1 << 21) \
            and not os.environ.get("DR_NO_TILED_GEMM"):
        # float64 / integer (and float32 shapes without a tensor path): register-tiled kernel
        from . import gemm
        node._stamp = _stamp_of(pa) + _stamp_of(pb)
        return gemm.matmul_tiled(a, b, res_dt)
    prod = ("t", len(prog.instrs))
    prog.instrs.append(("multiply", (res_dt, res_dt), res_dt, (remap["A"], remap["B"])))
    prog.dtypes[prod] = res_dt
    prog.roots = [prod]
    dev = prog.arrays[0].dev if prog.arrays else current_device()
    result = DeviceArray.empty((m, n), res_dt, dev)
    node._stamp = _stamp_of(prog)
    if m * n == 0:
        return result
    if k == 0:
        result.fill(0)
        return result
    _launch_axis_reduce(prog, triples, m, k, n, "sum", codegen.acc_dtype("sum", res_dt), res_dt,
                        1.0, result, dev)
    return result


_SKINNY_FOLD_SRC = r"""
extern "C" __global__ void __launch_bounds__(256) NAME(const TT* __restrict__ partial, TT* __restrict__ out,
    TT* __restrict__ rowsum, i64 m, int n_out, int n, int ksplit) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * n_out) return;
  const i64 row = i / n_out;
  const int c = (int)(i - row * n_out);
  ACC s = (ACC)0;
  for (int q = 0; q < ksplit; ++q) s += (ACC)partial[((i64)q * m + row) * n_out + c];
  if (c < n) out[row * n + c] = (TT)s; else rowsum[row] = (TT)s;
}
"""


def _try_mm_skinny(node, pa, m, k, n, res_dt, with_rowsum=None):
    """A(M,K) @ B(K,n<=7) with a fused all-pairs producer -> gen_mm_skinny.  If a live, not yet
    evaluated ReduceEx(sum, A, axis=1) shares the same producer it is folded in as a virtual
    column of ones and receives its result from this pass too."""
    from .delayarray import NPArray, ReduceEx
    if os.environ.get("DR_NO_SKINNY") or n > 7 or res_dt.kind != "f" or m < 512 or k < 512:
        return None
    a_node, b_node = node.arg1, node.arg2
    roles = []
    for arr in pa.arrays:
        sm, sk = planner.broadcast_strides(arr, (m, k))
        if sm == 0 and sk == 0:
            roles.append("b")
        elif sk == 0:
            roles.append("r")
        elif sm == 0:
            roles.append("c")
        else:
            return None
    if not pa.instrs:
        return None
    b_dev = b_node._force()
    if b_dev.dtype != res_dt:
        b_dev = b_dev.astype(res_dt)
    rowsum = None
    for cons in list(getattr(a_node, "_consumers", ())):
        if isinstance(cons, ReduceEx) and cons.op == "sum" and cons.axes == (1,) and not cons.keepdims \
                and cons.post is None and cons.dtype == res_dt and cons.__dict__.get("array") is None:
            rowsum = cons
            break
    n_out = n + (1 if rowsum is not None else 0)
    dev = b_dev.dev
    st = dev_state(dev)
    ksplit = max(1, min(64, -(-k // 2048)))
    kchunk = -(-k // ksplit)
    kchunk = -(-kchunk // 256) * 256
    ksplit = -(-k // kchunk)
    pow_modes = codegen.relaxed_pow_modes(pa)
    R, TH = int(os.environ.get("DR_SK_R", 8)), int(os.environ.get("DR_SK_TH", 64))
    key = ("mm_skinny", pa.key(), tuple(roles), res_dt.str, n_out, tuple(sorted(pow_modes.items())), R, TH)
    kern = get_kernel(key, lambda name: codegen.gen_mm_skinny(
        name, pa, roles, res_dt, n_out, threads=TH, rows=R, pow_modes=pow_modes))
    partial = DeviceArray.empty((ksplit, m, n_out), res_dt, dev if dev >= 0 else None)
    a = Args()
    n_ops = max(len(pa.arrays), 1)
    sr, sc = [0] * n_ops, [0] * n_ops
    for i, arr in enumerate(pa.arrays):
        sr[i], sc[i] = planner.broadcast_strides(arr, (m, k))
    geo = np.asarray([m, k, kchunk] + sr + sc + [b_dev.strides[0], b_dev.strides[1]],
                     dtype=np.int64).tobytes() + np.asarray([n, 0], dtype=np.int32).tobytes()
    a.raw(geo, 8)
    for arr in pa.arrays:
        a.ptr(arr.ptr)
    for val, dt in pa.scalars:
        a.scalar(val, dt)
    a.ptr(b_dev.ptr)
    a.ptr(partial.ptr)
    launch(kern, dev, (-(-m // (TH * R)), ksplit, 1), TH, a)
    # one small kernel folds the K-split partials in split order (float32 accumulates in double) and
    # writes the product and, if it rode along, the row sum: was a planned axis reduction plus two
    # strided copies -- three launches and most of the host time of an n-body step
    T = codegen.ctype(res_dt)
    ACC = "double" if res_dt == np.float32 else T
    fold = get_kernel(("mm_skinny_fold", res_dt.str), lambda name: _SKINNY_FOLD_SRC.replace("NAME", name)
                      .replace("ACC", ACC).replace("TT", T))
    d = dev if dev >= 0 else None
    out = DeviceArray.empty((m, n), res_dt, d)
    rs = DeviceArray.empty((m,), res_dt, d) if rowsum is not None else None
    a = Args()
    a.ptr(partial.ptr)
    a.ptr(out.ptr)
    a.ptr(rs.ptr if rs is not None else 0)
    a.i64(m)
    for v in (n_out, n, ksplit):
        a.scalar(v, np.int32)
    launch(fold, dev, max(1, -(-m * n_out // 256)), 256, a)
    node._stamp = _stamp_of(pa)
    if rowsum is not None:
        rowsum.array = rs
        rowsum._stamp = _stamp_of(pa)
    return out


# --------------------------------------------------------------------------- assignment / copies
def _overlaps(a, b):
    """Conservative: two views of one buffer whose byte extents intersect."""
    if a.buf is not b.buf:
        return False

    def extent(x):
        lo = hi = x.offset
        for n, s in zip(x.shape, x.strides):
            if n == 0:
                return (0, 0)
            if s >= 0:
                hi += (n - 1) * s
            else:
                lo += (n - 1) * s
        return lo, hi + x.dtype.itemsize
    (alo, ahi), (blo, bhi) = extent(a), extent(b)
    return alo < bhi and blo < ahi


def assign(target, value):
    """target[...] = value with NumPy semantics  (reference delayarray.py:114-121: force the
    RHS into a temporary, then copy).  Here the RHS expression is fused and written straight
    into the target view; a temporary is used only when the RHS reads the target's buffer at
    *other* elements than the one being written (e.g. shifted stencil views)."""
    from .delayarray import DelayArray, NPArray, Scalar, arg_to_numpy_ex
    from numbers import Number
    if isinstance(value, Number):
        node = Scalar(value)
        src = node if node.weak_type is None else Scalar(target.dtype.type(value))
    elif isinstance(value, DelayArray):
        src = value
    else:
        src = arg_to_numpy_ex(value if isinstance(value, (DeviceArray, np.ndarray))
                              else np.asarray(value))
    if src.kind in ("reduce", "matmul"):
        src = NPArray(src._force())
    if src.kind != "scalar" and tuple(src.shape) != target.shape:
        np.broadcast_shapes(src.shape, target.shape)          # raises on mismatch
        if len(src.shape) > target.ndim or np.broadcast_shapes(src.shape, target.shape) != target.shape:
            raise ValueError(f"could not broadcast input array from shape {src.shape} "
                             f"into shape {target.shape}")
    if target.size == 0:
        return
    st_key = st_ops = None
    if _PLAN_CACHE and src.kind == "ewise" and src.__dict__.get("_psig") is not None:
        st_key, st_ops = _stencil_plan_key(src, target)
        if st_key is not None:
            plan = _st_plans.get(st_key)
            if plan is not None and _stencil_plan_launch(plan, st_ops, target):
                plan.keep = src           # keeps the right-hand side hash-consed until the next step
                target.buf.version += 1
                return
    prog = planner.build_program([src]) if src.kind != "scalar" else _scalar_program(src)
    prog.shape = tuple(target.shape)
    tkey = (target.offset, planner.broadcast_strides(target, target.shape))
    hazard, inplace = False, False
    for arr in prog.arrays:
        if arr.buf is target.buf and _overlaps(arr, target):
            if (arr.offset, planner.broadcast_strides(arr, target.shape)) == tkey \
                    and arr.dtype == target.dtype:
                inplace = True
            else:
                hazard = True
    if hazard and _try_stencil(prog, target, st_key, st_ops):
        target.buf.version += 1
        return
    if hazard:
        tmp = DeviceArray.empty(target.shape, target.dtype, target.dev)
        run_program(prog, [tmp])
        p2 = planner.build_program([NPArray(tmp)])
        p2.shape = tuple(target.shape)
        run_program(p2, [target])
    else:
        run_program(prog, [target], inplace=inplace)
    target.buf.version += 1


_TMA_DTYPE = {"float32": 0, "float64": 1, "int32": 2, "int64": 3, "uint8": 4}


def encode_tensormap(dev, dtype_name, gptr, dims, strides_bytes, box, swizzle=0, l2promo=2):
    """cuTensorMapEncodeTiled through the C ABI; returns the 128 descriptor bytes.  The driver
    wants the descriptor 64-byte aligned, so it is built inside an over-allocated buffer."""
    raw = (C.c_uint8 * 256)()
    base = C.addressof(raw)
    off = (-base) % 64
    rank = len(dims)
    if dev >= 0:
        d = (C.c_uint64 * rank)(*dims)
        st = (C.c_uint64 * max(rank - 1, 1))(*(list(strides_bytes) or [0]))
        bx = (C.c_uint32 * rank)(*box)
        check(lib.drc_tensormap_encode(dev, C.c_void_p(base + off), _TMA_DTYPE[dtype_name], rank, gptr,
                                       d, st, bx, swizzle, l2promo))
    return bytes(raw[off:off + 128])


# Prepared stencil launches (the assignment counterpart of _plans): keyed on the RHS signature,
# the target view and, for every leaf that aliases the target's buffer, its exact offset (the
# shift).  A replay allocates the ping-pong buffer, reuses the tensor map encoded for the current
# base address (two addresses alternate) and launches; no planning, no role analysis.
_st_plans = {}


class _StPlan:
    __slots__ = ("kern", "meta", "geo", "grid", "arr_idx", "leaf_sigs", "sc_idx", "sc_dt", "tmaps",
                 "cols", "rows", "pitch", "tiles_x", "max_dy", "keep", "layout")


def _stencil_plan_key(src, target):
    ops = src._pops
    tb = target.buf
    lk = []
    for o in ops:
        if o.kind == "leaf":
            arr = o.array
            if not isinstance(arr, DeviceArray):
                return None, None
            lk.append(arr.offset if arr.buf is tb else -1)
    return (src._psig, src.shape, target.offset, target.shape, target.strides, target.dtype.str,
            tb.nbytes, tuple(lk), target.dev, tb.halo is not None), ops


def _stencil_plan_launch(plan, ops, target):
    from .delayarray import _leaf_sig
    arrays = []
    for i, want in zip(plan.arr_idx, plan.leaf_sigs):
        leaf = ops[i]
        if _leaf_sig(leaf) != want:
            return False
        arrays.append(leaf._force())
    buf, dev, m = target.buf, target.dev, plan.meta
    link = buf.halo
    if link is not None:
        link.before_stencil(plan.max_dy)
        out = link.partner
    else:
        out = DeviceBuffer(buf.nbytes, dev) if dev >= 0 else DeviceBuffer(buf.nbytes)
    tmap = plan.tmaps.get(buf.ptr)
    if tmap is None:
        if len(plan.tmaps) >= 8:
            plan.tmaps.clear()
        tmap = plan.tmaps[buf.ptr] = encode_tensormap(
            dev, target.dtype.name, buf.ptr, (plan.cols, plan.rows), (plan.pitch,), (m["BW"], m["BH"]))
    lay = plan.layout
    if lay is None:
        a = Args()
        a.raw(tmap, 64)
        a.raw(plan.geo, 8)
        a.ptr(buf.ptr)
        a.ptr(out.ptr)
        if link is not None:
            a.raw(link.kernel_args(plan.rows, plan.pitch, m["TH"], plan.tiles_x), 8)
        for arr in arrays:
            a.ptr(arr.ptr)
        for i, dt in zip(plan.sc_idx, plan.sc_dt):
            a.scalar(ops[i].val, dt)
        launch(plan.kern, dev, plan.grid, m["threads"], a, smem=m["smem"])
        if all(np.dtype(dt).kind == "f" for dt in plan.sc_dt):
            a.pack()
            plan.layout = _arg_layouts["".join(a.fmt)]
    else:
        # the argument layout never changes for a plan: pack the values straight into it
        vals = [tmap, plan.geo, buf.ptr, out.ptr]
        if link is not None:
            vals.append(link.kernel_args(plan.rows, plan.pitch, m["TH"], plan.tiles_x))
        for arr in arrays:
            vals.append(arr.ptr)
        for i, dt in zip(plan.sc_idx, plan.sc_dt):
            vals.append(float(dt.type(ops[i].val)))
        _last_kernel[0] = plan.kern.name
        if dev >= 0:
            check(lib.drc_launch_packed(dev, 0, plan.kern.func(dev), plan.grid, 1, 1, m["threads"], 1, 1,
                                        m["smem"], 1, lay[0].pack(*vals), lay[1], lay[2]))
            stats["launches"] += 1
        else:
            dry_log.append((plan.kern, plan.grid, m["threads"]))
    buf.swap_storage(out)
    if link is not None:
        link.after_stencil()
    stats["plan_hits"] = stats.get("plan_hits", 0) + 1
    return True


def _try_stencil(prog, target, plan_key=None, plan_ops=None):
    """Shifted-view self-assignment on a 2-d base array -> TMA-staged stencil kernel writing a
    fresh copy of the base, then the two allocations are swapped (ping-pong).  Returns False
    when the pattern does not apply (the caller falls back to temporary + copy)."""
    if os.environ.get("DR_NO_STENCIL") or target.ndim != 2 or target.dtype.name not in ("float32", "float64"):
        return False
    item = target.dtype.itemsize
    V = 16 // item
    buf = target.buf
    pitch = target.strides[0]
    if target.strides[1] != item or pitch % 16 or pitch <= 0 or buf.ptr % 16:
        return False
    cols = pitch // item
    if cols % V:
        return False
    rows = buf.nbytes // pitch
    if rows * pitch != buf.nbytes or rows >= (1 << 31) or cols >= (1 << 31):
        return False                       # the buffer is not exactly one rows x cols matrix
    r0, rem = divmod(target.offset, pitch)
    c0 = rem // item
    h, w = target.shape
    if c0 + w > cols or r0 + h > rows or 2 * h * w < rows * cols:
        return False
    roles = []
    for arr in prog.arrays:
        if arr.buf is buf:
            if arr.dtype != target.dtype or arr.shape != target.shape or arr.strides != target.strides:
                return False
            ar, rem = divmod(arr.offset, pitch)
            ac = rem // item
            if ac + w > cols:
                return False
            roles.append(("tile", ar - r0, ac - c0))
        elif all(s == 0 for s in planner.broadcast_strides(arr, target.shape)):
            roles.append(("b",))
        else:
            roles.append(("g",))
    tiles = [r for r in roles if r[0] == "tile"]
    if not tiles or max(abs(r[1]) for r in tiles) > 8 or max(abs(r[2]) for r in tiles) > 8:
        return False
    dev = target.dev
    st = dev_state(dev)
    link = buf.halo          # row-sharded block: the kernel also exchanges the halo rows
    max_dy = max(abs(r[1]) for r in tiles)
    if link is not None:
        link.before_stencil(max_dy)
    halo = link is not None
    key = ("stencil", prog.key(), tuple(roles), target.dtype.str) + (("halo",) if halo else ())
    meta_box = {}

    def gen(name):
        src, meta = codegen.gen_stencil(name, prog, roles, target.dtype, halo=halo)
        meta_box.update(meta)
        return src
    kern = get_kernel(key, gen, meta_box)
    if not kern.meta:
        kern.meta.update(codegen.gen_stencil("x", prog, roles, target.dtype, halo=halo)[1])
    m = kern.meta
    if halo:
        out = link.partner
    else:
        out = DeviceBuffer(buf.nbytes, dev) if dev >= 0 else DeviceBuffer(buf.nbytes)
    tiles_x, tiles_y = -(-cols // m["TW"]), -(-rows // m["TH"])
    a = Args()
    a.raw(encode_tensormap(dev, target.dtype.name, buf.ptr, (cols, rows), (pitch,),
                           (m["BW"], m["BH"])), 64)
    n_ops = max(len(prog.arrays), 1)
    gs_row, gs_col = [0] * n_ops, [0] * n_ops
    for i, (arr, role) in enumerate(zip(prog.arrays, roles)):
        if role[0] == "g":
            bst = planner.broadcast_strides(arr, target.shape)
            gs_row[i], gs_col[i] = bst
    geo = np.asarray([rows, cols, r0, c0, h, w, tiles_x, tiles_x * tiles_y], dtype=np.int32).tobytes() \
        + np.asarray([cols] + gs_row + gs_col, dtype=np.int64).tobytes()
    a.raw(geo, 8)
    a.ptr(buf.ptr)
    a.ptr(out.ptr)
    if halo:
        a.raw(link.kernel_args(rows, pitch, m["TH"], tiles_x), 8)
    for arr in prog.arrays:
        a.ptr(arr.ptr)
    for val, dt in prog.scalars:
        a.scalar(val, dt)
    if dev >= 0 and dev not in kern.meta.setdefault("smem_set_devs", set()):
        check(lib.drc_func_set_max_dynamic_smem(dev, kern.func(dev), m["smem"]))
        kern.meta["smem_set_devs"].add(dev)
    per_sm = max(1, min(8, (st.max_smem - 1024) // (m["smem"] + 1024)))
    if dev >= 0:
        per_sm = min(per_sm, kern.blocks_per_sm(dev, m["threads"], m["smem"]))
    grid = min(tiles_x * tiles_y, st.sm_count * per_sm)
    launch(kern, dev, grid, m["threads"], a, smem=m["smem"])
    if plan_key is not None:
        _stencil_plan_record(plan_key, plan_ops, prog, kern, geo, grid, cols, rows, pitch, tiles_x, max_dy)
    buf.swap_storage(out)              # `out` now owns the old allocation and frees it (stream-ordered)
    if halo:
        link.after_stencil()
    return True


def _stencil_plan_record(key, ops, prog, kern, geo, grid, cols, rows, pitch, tiles_x=0, max_dy=0):
    from .delayarray import _leaf_sig
    arr_idx, leaf_sigs, sc_idx = [], [], []
    for arr in prog.arrays:
        for i, o in enumerate(ops):
            if o.kind == "leaf" and o._force() is arr:
                sig = _leaf_sig(o)
                if sig is None:
                    return
                arr_idx.append(i)
                leaf_sigs.append(sig)
                break
        else:
            return
    for node in prog.scalar_nodes:
        for i, o in enumerate(ops):
            if o is node:
                sc_idx.append(i)
                break
        else:
            return
    if len(_st_plans) >= _PLAN_LIMIT:
        _st_plans.clear()
    p = _StPlan()
    p.kern, p.meta, p.geo, p.grid = kern, kern.meta, geo, grid
    p.arr_idx, p.leaf_sigs, p.sc_idx = tuple(arr_idx), tuple(leaf_sigs), tuple(sc_idx)
    p.sc_dt = tuple(dt for _, dt in prog.scalars)
    p.tmaps, p.cols, p.rows, p.pitch = {}, cols, rows, pitch
    p.tiles_x, p.max_dy, p.keep, p.layout = tiles_x, max_dy, None, None
    _st_plans[key] = p


def _scalar_program(node):
    prog = planner.Program()
    prog.scalars.append((node.val, node.dtype))
    prog.dtypes[("s", 0)] = node.dtype
    prog.roots = [("s", 0)]
    return prog


def materialize_view(arr):
    """Contiguous copy of a (possibly strided) DeviceArray.  Complex data (fft results) is moved
    as (re, im) pairs of its component type."""
    from .delayarray import NPArray
    if arr.dtype.kind == "c":
        part = np.dtype(np.float32 if arr.dtype == np.complex64 else np.float64)
        pairs = DeviceArray(arr.buf, arr.shape + (2,), part, arr.strides + (part.itemsize,), arr.offset)
        flat = materialize_view(pairs)
        return DeviceArray(flat.buf, arr.shape, arr.dtype, None, flat.offset)
    from . import extras
    # X.T.copy(): a 2-d transpose is a one-operand region of the tile family (5.3 / 5.6 TB/s for
    # 4- / 8-byte words); batched transposes keep the dedicated kernel (4.7 TB/s)
    fast = extras.transpose_copy(arr) if arr.ndim != 2 or os.environ.get("DR_NO_TILE") else None
    if fast is not None:
        return fast
    outs, _ = evaluate_nodes([NPArray(arr)])
    return outs[0]


def cast_array(arr, dtype):
    from .delayarray import CastEx, NPArray
    outs, _ = evaluate_nodes([CastEx(NPArray(arr), dtype)])
    return outs[0]


def cumsum(arr, axis=None):
    from . import extras
    return extras.cumsum(arr, axis)


def synchronize(dev=None):
    check(lib.drc_device_sync(current_device() if dev is None else dev))
