"""DAG -> fused-region planner (replaces the reference's tree-walking Visitor, visitor.py:4-24,
and the statement-list construction of CupyEmitter, cuda.py:46-88).

One *program* per fused region: an SSA list over typed operands, built by an ITERATIVE
post-order walk of the DAG (each node once -- no Python recursion, no exponential re-walk of
shared sub-expressions), with value numbering on top of the capture layer's
hash-consing.  Operand and temporary names are positional, never the global node counter, so
structurally equal expressions produce byte-identical source and hit the cubin cache
(the reference recompiles every new expression object).

Cut points: reductions, contractions and already materialised nodes become region inputs.
Layout resolution: every array operand is broadcast to the iteration shape and all operands
are collapsed jointly to the fewest dimensions; the result picks the kernel family
("flat" = one contiguous vectorised dimension, "nd" = general strided/broadcast).
"""
import numpy as np

from .device import DeviceArray


_DSTR = {}


def _dstr(dt):
    """dtype -> its '<f4'-style string, cached (np.dtype.str builds a new str on every access)."""
    s = _DSTR.get(dt)
    if s is None:
        s = _DSTR[dt] = dt.str
    return s


class Program:
    """SSA program of one fused region.

    arrays : list of DeviceArray        -- array operands, in first-use order
    scalars: list of (value, np.dtype)  -- typed scalar kernel arguments
    instrs : list of (op, loop_dtypes, out_dtype, args); args are ('a', i) | ('s', j) | ('t', k)
    roots  : list of operand refs, one per requested output
    """

    __slots__ = ("arrays", "scalars", "instrs", "roots", "shape", "leaf_bufs", "dtypes",
                 "scalar_nodes")

    def __init__(self):
        self.arrays, self.scalars, self.instrs, self.roots = [], [], [], []
        self.scalar_nodes = []      # the Scalar node behind each entry of `scalars` (plan cache)
        self.shape = ()
        self.leaf_bufs = []
        self.dtypes = {}            # operand ref -> np.dtype

    def key(self):
        """Structure only: no sizes, pointers or scalar values."""
        ds = _dstr
        return (tuple([ds(a.dtype) for a in self.arrays]),
                tuple([ds(dt) for _, dt in self.scalars]),
                tuple([(op, tuple([ds(d) for d in loop]), ds(out), args)
                       for op, loop, out, args in self.instrs]),
                tuple(self.roots))


def _is_materialised(node):
    # an evaluated node is a snapshot (DelayArray._force): its array is what its consumers read,
    # whether or not its inputs have been written since
    return node.__dict__.get("array") is not None


def build_program(roots):
    """Linearise the elementwise DAG under ``roots`` (list of nodes) into a Program."""
    prog = Program()
    ref = {}                 # id(node) -> operand ref
    array_slot = {}          # layout key -> index into prog.arrays
    scalar_slot = {}         # (id(node), dtype) -> index into prog.scalars
    value_no = {}            # (op, loop, args) -> temp ref

    def array_operand(node, dev_arr):
        lk = dev_arr.layout_key()
        idx = array_slot.get(lk)
        if idx is None:
            idx = array_slot[lk] = len(prog.arrays)
            prog.arrays.append(dev_arr)
        r = ("a", idx)
        prog.dtypes[r] = dev_arr.dtype
        ref[id(node)] = r

    def scalar_operand(node, dtype):
        k = (id(node), dtype.str)
        idx = scalar_slot.get(k)
        if idx is None:
            idx = scalar_slot[k] = len(prog.scalars)
            prog.scalars.append((dtype.type(node.val), dtype))
            prog.scalar_nodes.append(node)
        r = ("s", idx)
        prog.dtypes[r] = dtype
        return r

    def emit(op, loop, out_dt, args):
        vn = (op, loop, out_dt, args)
        hit = value_no.get(vn)
        if hit is None:
            hit = value_no[vn] = ("t", len(prog.instrs))
            prog.instrs.append((op, loop, out_dt, args))
            prog.dtypes[hit] = out_dt
        return hit

    neg = {}                 # id(node) -> True when ref[id(node)] holds MINUS the node's value
    stack = [(r, False) for r in reversed(roots)]
    while stack:
        node, expanded = stack.pop()
        if id(node) in ref:
            continue
        kind = node.kind
        if kind == "scalar":
            continue                       # typed at each use
        if kind == "leaf":
            array_operand(node, node._force())
            continue
        if kind in ("reduce", "matmul") or (not expanded and _is_materialised(node)):
            array_operand(node, node._force())      # cut point: evaluate (or reuse) first
            continue
        if not expanded:
            stack.append((node, True))
            for kid in reversed(node.children):
                if id(kid) not in ref:
                    stack.append((kid, False))
            continue
        args, negs = [], []
        for kid, loop_dt in zip(node.children, node.loop):
            if kid.kind == "scalar":
                args.append(scalar_operand(kid, kid.dtype if kid.weak_type is None else loop_dt))
                negs.append(False)
            else:
                r = ref[id(kid)]
                args.append(r)
                negs.append(neg.get(id(kid), False))
        ref[id(node)], flag = _emit_signed(node, args, negs, emit)
        if flag:
            neg[id(node)] = True

    prog.roots = []
    for r in roots:
        operand = ref[id(r)]
        if neg.get(id(r), False):
            operand = emit("negative", (r.dtype,), r.dtype, (operand,))
        prog.roots.append(operand)
    prog.shape = tuple(np.broadcast_shapes(*[r.shape for r in roots])) if roots else ()
    seen = set()
    for a in prog.arrays:
        if id(a.buf) not in seen:
            seen.add(id(a.buf))
            prog.leaf_bufs.append(a.buf)
    return prog


# --------------------------------------------------------------------------- sign hoisting
_SIGN_PRODUCT = ("multiply", "true_divide", "divide")
_ODD = ("erf",)              # f(-x) == -f(x) bit for bit in NumPy/SciPy and on the device


def _emit_signed(node, args, negs, emit):
    """Emit ``node`` with negations hoisted outwards: -x is never computed, the sign travels
    as a flag through * and / ((-a)*b == -(a*b) exactly, zeros included), odd functions, and
    is absorbed by + and - ((a) + (-b) == a - b is the IEEE definition).  Black-Scholes'
    cnd(-d1), cnd(-d2) thereby reuse erf(d1), erf(d2): two erf evaluations instead of four,
    with bit-identical results.  Cases that would change the sign of a zero sum
    ((-a)+(-b) vs -(a+b)) are not rewritten."""
    op, loop, dt = node.op, node.loop, node.dtype
    plain = dt.kind == "f" and all(d == dt for d in loop)

    def solid(i):
        return emit("negative", (loop[i],), loop[i], (args[i],)) if negs[i] else args[i]

    if plain and any(negs) or (plain and op == "negative"):
        if op == "negative":
            return args[0], not negs[0]
        if op in _SIGN_PRODUCT:
            return emit(op, loop, dt, tuple(args)), negs[0] != negs[1]
        if op in _ODD:
            return emit(op, loop, dt, tuple(args)), negs[0]
        if op in ("absolute", "fabs"):
            return emit(op, loop, dt, tuple(args)), False
        if op == "add":
            if negs == [False, True]:
                return emit("subtract", loop, dt, (args[0], args[1])), False
            if negs == [True, False]:
                return emit("subtract", loop, dt, (args[1], args[0])), False
        if op == "subtract":
            if negs == [False, True]:
                return emit("add", loop, dt, (args[0], args[1])), False
            if negs == [True, True]:
                return emit("subtract", loop, dt, (args[1], args[0])), False
    return emit(op, loop, dt, tuple(solid(i) for i in range(len(args)))), False


# --------------------------------------------------------------------------- layout
def broadcast_strides(arr, shape):
    """Byte strides of ``arr`` viewed at iteration shape ``shape`` (0 on broadcast dims)."""
    nd = len(shape)
    shp = (1,) * (nd - arr.ndim) + tuple(arr.shape)
    st = (0,) * (nd - arr.ndim) + tuple(arr.strides)
    out = []
    for want, have, s in zip(shape, shp, st):
        if have == want and want != 1:
            out.append(s)
        elif have == 1 or want == 1:
            out.append(0)
        else:
            raise ValueError(f"operand of shape {arr.shape} does not broadcast to {shape}")
    return tuple(out)


def collapse(shape, strides_list):
    """Jointly collapse dimensions: drop size-1 dims, merge dims that are contiguous for
    every operand.  Returns (shape, [strides per operand]); at least one dimension."""
    dims = [i for i, n in enumerate(shape) if n != 1]
    if not dims:
        return (1,), [(0,) for _ in strides_list]
    shp = [shape[i] for i in dims]
    sts = [[st[i] for i in dims] for st in strides_list]
    out_shape, out_sts = [shp[0]], [[st[0]] for st in sts]
    for d in range(1, len(shp)):
        n = shp[d]
        if all(o[-1] == st[d] * n for o, st in zip(out_sts, sts)):
            out_shape[-1] *= n
            for o, st in zip(out_sts, sts):
                o[-1] = st[d]
        else:
            out_shape.append(n)
            for o, st in zip(out_sts, sts):
                o.append(st[d])
    return tuple(out_shape), [tuple(o) for o in out_sts]


class Layout:
    """Resolved geometry of one region launch."""

    __slots__ = ("shape", "in_strides", "out_strides", "family", "in_class", "vec_ok", "total")

    def key(self):
        return (self.family, len(self.shape), self.in_class, self.vec_ok)


def resolve_layout(prog, outs, inner_vectors=True):
    """Pick the kernel family for program ``prog`` writing into DeviceArrays ``outs``."""
    shape = prog.shape
    # fast path: every operand is a C-contiguous array of exactly the iteration shape
    if len(shape) >= 1 and all(a.shape == shape and a.is_contiguous for a in prog.arrays) \
            and all(o.shape == shape and o.is_contiguous for o in outs):
        lay = Layout()
        total = 1
        for n in shape:
            total *= n
        lay.shape, lay.total = (total,), total
        lay.in_strides = [(a.dtype.itemsize,) for a in prog.arrays]
        lay.out_strides = [(o.dtype.itemsize,) for o in outs]
        lay.family = "flat"
        lay.in_class = ("c",) * len(prog.arrays)
        lay.vec_ok = all(a.ptr % 16 == 0 for a in prog.arrays) and all(o.ptr % 16 == 0 for o in outs)
        if total > 1 or not prog.arrays:
            return lay
    ins = [broadcast_strides(a, shape) for a in prog.arrays]
    out_st = [broadcast_strides(o, shape) for o in outs]
    cshape, csts = collapse(shape, ins + out_st)
    lay = Layout()
    lay.shape = cshape
    lay.in_strides = csts[:len(ins)]
    lay.out_strides = csts[len(ins):]
    lay.total = 1
    for n in cshape:
        lay.total *= n
    flat = len(cshape) == 1
    cls = []
    if flat:
        for a, st in zip(prog.arrays, lay.in_strides):
            if st[0] == a.dtype.itemsize:
                cls.append("c")                 # contiguous, walks with the index
            elif st[0] == 0:
                cls.append("b")                 # broadcast scalar: one load per thread
            else:
                flat = False
        for o, st in zip(outs, lay.out_strides):
            if st[0] != o.dtype.itemsize:
                flat = False
    if flat:
        aligned = all(a.ptr % 16 == 0 for a, c in zip(prog.arrays, cls) if c == "c") and \
            all(o.ptr % 16 == 0 for o in outs)
        lay.family = "flat"
        lay.in_class = tuple(cls)
        lay.vec_ok = bool(aligned)
    else:
        lay.family = "nd"
        lay.in_class = tuple("b" if all(s == 0 for s in st) else "s" for st in lay.in_strides)
        lay.vec_ok = False
        if inner_vectors and outs:
            _try_inner_vectors(prog, outs, lay)
    return lay


def tile_classes(prog, outs, lay):
    """An `nd` layout over a 2-d space whose outputs are contiguous along the last dimension and
    in which at least one operand is contiguous along the FIRST one (a transposed matrix): the
    operand classes and tile edge for codegen.gen_tile, or None."""
    if lay.family != "nd" or lay.vec_ok or len(lay.shape) != 2 or not outs or lay.total >= (1 << 40):
        return None
    if lay.shape[0] < 32 or lay.shape[1] < 32:
        return None
    for o, st in zip(outs, lay.out_strides):
        if st[1] != o.dtype.itemsize:
            return None
    cls, staged = [], []
    for a, st in zip(prog.arrays, lay.in_strides):
        item = a.dtype.itemsize
        if st[0] == 0 and st[1] == 0:
            cls.append("b")
        elif st[1] == item:
            cls.append("v")
        elif st[0] == item and item in (4, 8) and abs(st[1]) >= item:
            cls.append("t")             # (a column vector, stride 0 along C, is a plain broadcast load)
            staged.append(item)
        else:
            cls.append("s")
    if not staged:
        return None
    T = 64 if max(staged) == 4 else 32
    if sum(T * (T + 1) * i for i in staged) > 40 * 1024:
        T = 32
        if sum(T * (T + 1) * i for i in staged) > 40 * 1024:
            return None
    # two elements per thread (one 2-element vector access) when every staged operand is 4 bytes
    # wide and every vector access would be aligned
    W = 2 if T == 64 and max(staged) == 4 else 1
    if W == 2:
        for a, st, c in zip(prog.arrays, lay.in_strides, cls):
            al = 2 * a.dtype.itemsize
            if (c == "v" and (a.ptr % al or st[0] % al)) or (c == "t" and (a.ptr % al or st[1] % al)):
                W = 1
        for o, st in zip(outs, lay.out_strides):
            al = 2 * o.dtype.itemsize
            if o.ptr % al or st[0] % al:
                W = 1
    return tuple(cls), T, W


def _try_inner_vectors(prog, outs, lay):
    """nd family with 128-bit accesses along the innermost collapsed dimension: every operand is
    either contiguous along it ('v': stride == itemsize, vector load/store) or does not move along
    it ('i': stride 0, one scalar load reused by the vector's lanes; 'b' stays a kernel-wide
    scalar).  Needs equal-width element types, an inner extent that is a multiple of the vector
    length and 16-byte aligned bases and outer strides.  Rewrites lay in place (vec_ok = V)."""
    sizes = {a.dtype.itemsize for a in prog.arrays} | {o.dtype.itemsize for o in outs}
    if len(sizes) != 1:
        return
    item = sizes.pop()
    V = 16 // item
    if V < 2 or lay.shape[-1] % V or lay.total >= (1 << 32):
        return
    cls = []
    for a, st, c in zip(prog.arrays, lay.in_strides, lay.in_class):
        if c == "b":
            cls.append("b")
        elif st[-1] == item:
            if a.ptr % 16 or any(x % 16 for x in st[:-1]):
                return
            cls.append("v")
        elif st[-1] == 0:
            cls.append("i")
        else:
            return
    for o, st in zip(outs, lay.out_strides):
        if st[-1] != item or o.ptr % 16 or any(x % 16 for x in st[:-1]):
            return
    if "v" not in cls and not outs:
        return
    lay.in_class = tuple(cls)
    lay.vec_ok = V
    lay.shape = lay.shape[:-1] + (lay.shape[-1] // V,)
    lay.total //= V
    lay.in_strides = [st[:-1] + (st[-1] * V,) for st in lay.in_strides]
    lay.out_strides = [st[:-1] + (st[-1] * V,) for st in lay.out_strides]
