"""CUDA C++ code generator: one kernel per fused region (replaces CupyEmitter/make_kernel,
reference cuda.py:35-88, which builds a `T name = expr` statement list for
cupy.ElementwiseKernel).

Kernel families
  flat      one contiguous dimension: 128-bit vector loads/stores (ld.global.nc.L1::no_allocate /
            st.global.cs), U independent vectors in flight per thread, grid-stride, scalar tail
  nd        general strided / broadcast operands: index decomposition, coalesced scalar accesses
  *+reduce  either family with a fused full reduction: per-thread accumulators -> warp shuffle
            -> shared memory -> one partial per block -> deterministic last-block finish
  rows      reduction over the trailing axis of a (rows, cols) iteration space (sum(axis=-1),
            matrix @ vector): one warp or one block per row, fused producer
Everything is emitted with positional names so equal structure == equal source text.
"""
import os

import numpy as np

from . import ranges

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "math_tables.cuh")) as _f:
    ERF2_ROWS = int([ln.split()[2] for ln in _f if ln.startswith("#define DR_ERF2_ROWS")][0])
# third-generation erf (tools/gen_erf3.py): table uniform in sqrt(|x|/4), ONE LDS.128 per evaluation
# (4 shared-memory wavefronts instead of 6: the staged kernels are bound by the shared-memory pipe)
# and 12 instead of 14 issue slots.  Its source is appended ONLY to the staged kernels that use it,
# so no other kernel's text (= cubin cache key) changes.
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "erf3.cuh")) as _f:
    ERF3_SRC = _f.read()
ERF3_ROWS = int([ln.split()[2] for ln in ERF3_SRC.splitlines() if ln.startswith("#define DR_ERF3_ROWS")][0])
ERF3 = os.environ.get("DR_ERF_GEN", "3") == "3"

CTYPE = {
    "?": "bool", "b": "signed char", "B": "unsigned char", "h": "short", "H": "unsigned short",
    "i": "int", "I": "unsigned int", "l": "long long", "L": "unsigned long long",
    "q": "long long", "Q": "unsigned long long", "f": "float", "d": "double",
}


_UNSIGNED = {"b": "unsigned int", "h": "unsigned int", "i": "unsigned int", "l": "unsigned long long",
             "q": "unsigned long long"}


def ctype(dt):
    return CTYPE[np.dtype(dt).char]


def _lit(dt, v):
    dt = np.dtype(dt)
    if dt.kind == "f":
        return f"(({ctype(dt)}){float(v)!r})".replace("inf", "DR_INF").replace("nan", "DR_NAN")
    return f"(({ctype(dt)}){int(v)})"


_INFIX = {"add": "+", "subtract": "-", "multiply": "*", "true_divide": "/", "divide": "/",
          "greater": ">", "greater_equal": ">=", "less": "<", "less_equal": "<=",
          "equal": "==", "not_equal": "!=", "bitwise_and": "&", "bitwise_or": "|",
          "bitwise_xor": "^", "left_shift": "<<", "right_shift": ">>"}
_CALL1 = {"sqrt", "exp", "exp2", "expm1", "log", "log2", "log10", "log1p", "sin", "cos", "tan",
          "sinh", "cosh", "tanh", "cbrt", "erf", "erfc", "floor", "ceil", "trunc", "rint",
          "sign", "isnan", "isinf", "isfinite", "signbit"}
_CALL1_RENAMED = {"arcsin": "asin", "arccos": "acos", "arctan": "atan", "arcsinh": "asinh",
                  "arccosh": "acosh", "arctanh": "atanh", "absolute": "abs", "fabs": "abs"}
_CALL2 = {"hypot", "copysign", "fmod", "fmax", "fmin", "remainder", "floor_divide"}
_CALL2_RENAMED = {"arctan2": "atan2", "maximum": "max", "minimum": "min"}


_FAST_F32 = {"true_divide": "dr_div_fast({0}, {1}, bad)", "divide": "dr_div_fast({0}, {1}, bad)",
             "sqrt": "dr_sqrt_fast({0}, bad)", "log": "dr_log_fast({0}, bad)",
             "reciprocal": "dr_div_fast(1.0f, {0}, bad)"}
_HEAVY = {"exp", "log", "erf", "erfc", "sin", "cos", "tan", "power", "tanh", "sinh", "cosh",
          "arctan2", "arctan", "arcsin", "arccos", "exp2", "expm1", "log1p", "log2", "log10",
          "true_divide", "divide", "sqrt", "cbrt", "hypot"}


def has_fast_path(prog):
    """Does the program contain float32 ops whose precise form carries a slow-path branch?"""
    return any(op in _FAST_F32 and loop[0] == np.float32 for op, loop, _, _ in prog.instrs)


def body_weight(prog):
    return sum(1 for op, _, _, _ in prog.instrs if op in _HEAVY)


def emit_expr(op, loop, out_dt, args, arg_dts, fast=False, relaxed=False, pow_mode=None):
    """C expression for one SSA instruction; ``args`` are C expressions of dtype arg_dts.
    ``fast``: use the branch-free flag-raising float32 forms (see prelude, dr_*_fast)."""
    cast_args = []
    for a, have, want in zip(args, arg_dts, loop):
        if have == want:
            cast_args.append(a)
        elif want.kind == "b":
            cast_args.append(f"({a} != 0)")
        else:
            cast_args.append(f"(({ctype(want)})({a}))")
    a = cast_args
    if fast and op in _FAST_F32 and loop[0] == np.float32:
        return _FAST_F32[op].format(*a)
    T = ctype(loop[0])
    k = loop[0].kind
    O = ctype(out_dt)
    if op == "cast":
        return f"({a[0]} != 0)" if out_dt.kind == "b" else f"(({O})({a[0]}))"
    if op == "tf32_hi":          # nearest TF32-representable value (10 explicit mantissa bits)
        return f"__uint_as_float((__float_as_uint({a[0]}) + 0x1000u) & 0xffffe000u)"
    if op == "where":
        return f"({a[0]} ? {a[1]} : {a[2]})"
    if k == "b" and op in ("add", "maximum", "bitwise_or", "logical_or"):
        return f"({a[0]} || {a[1]})"
    if k == "b" and op in ("multiply", "minimum", "bitwise_and", "logical_and"):
        return f"({a[0]} && {a[1]})"
    if k == "b" and op in ("bitwise_xor", "logical_xor", "not_equal"):
        return f"({a[0]} != {a[1]})"
    if op in ("true_divide", "divide") and k in "iu":       # never produced by NumPy's resolver
        raise TypeError("integer true_divide")
    if k == "i" and op in ("add", "subtract", "multiply", "left_shift", "negative"):
        # NumPy's signed integers wrap; in C++ signed overflow is undefined and NVRTC uses that
        # (e.g. it widens `(double)(x * x * x)`): do the arithmetic in the unsigned type
        U = _UNSIGNED[loop[0].char]
        if op == "negative":
            # 0 - x with a zero the assembler cannot see (every launch has gridDim.z == 1): ptxas
            # 12.9 folds a literal negation into the operands of a fused three-input integer
            # min / max (VIMNMX3) and loses it for one of them -- min(-a, -b, -c) came back as
            # min(a', -b, ...) for int32 / int16 (fuzz seed 60525; the SASS shows
            # `VIMNMX3 R4, R3, R4, R15` with only R15 negated).  A real subtraction is one IADD3.
            return f"(({O})(({U})(gridDim.z - 1u) - ({U}){a[0]}))"
        return f"(({O})(({U}){a[0]} {_INFIX[op]} ({U}){a[1]}))"
    if op in _INFIX:
        e = f"({a[0]} {_INFIX[op]} {a[1]})"
        return e if out_dt.kind == "b" else f"(({O}){e})"
    if op == "negative":
        if k == "u":
            return f"(({O})(({O})(gridDim.z - 1u) - {a[0]}))"        # as above
        return f"(({O})(-{a[0]}))"
    if op == "positive":
        return a[0]
    if op == "reciprocal":
        return f"(({O})(({T})1 / {a[0]}))"
    if op == "logical_and":
        return f"(({a[0]} != 0) && ({a[1]} != 0))"
    if op == "logical_or":
        return f"(({a[0]} != 0) || ({a[1]} != 0))"
    if op == "logical_xor":
        return f"(({a[0]} != 0) != ({a[1]} != 0))"
    if op == "logical_not":
        return f"({a[0]} == 0)"
    if op == "invert":
        return f"(!{a[0]})" if k == "b" else f"(({O})(~{a[0]}))"
    if op == "power":
        if pow_mode == "rsqrt3":                 # x ** -1.5 inside a contraction (relaxed_pow_modes)
            return f"dr_rsqrt3_relaxed({a[0]})"
        if pow_mode == "rsqrt":
            return f"dr_rsqrt_relaxed({a[0]})"
        if relaxed and loop[0] == np.float32:
            return f"dr_pow_relaxed({a[0]}, {a[1]})"
        return f"dr_pow({a[0]}, {a[1]})" if k == "f" else f"dr_ipow<{T}>({a[0]}, {a[1]})"
    if op in ("deg2rad", "radians"):
        return f"({a[0]} * (({T})0.017453292519943295))"
    if op in ("rad2deg", "degrees"):
        return f"({a[0]} * (({T})57.29577951308232))"
    if op in _CALL1:
        return f"dr_{op}({a[0]})"
    if op in _CALL1_RENAMED:
        return f"dr_{_CALL1_RENAMED[op]}({a[0]})"
    if op in _CALL2:
        return f"dr_{op}<{T}>({a[0]}, {a[1]})" if k != "f" or op in ("fmax", "fmin") \
            else f"dr_{op}({a[0]}, {a[1]})"
    if op in _CALL2_RENAMED:
        fn = _CALL2_RENAMED[op]
        if fn == "atan2":
            return f"dr_{fn}({a[0]}, {a[1]})"
        # np.maximum / np.minimum return the SECOND operand when the two compare equal (their
        # loops are `a > b ? a : b` with nan fix-ups), which is visible for -0.0 against +0.0;
        # dr_max(x, y) = (x >= y || x != x) ? x : y keeps its first, hence the swapped operands
        return f"dr_{fn}<{T}>({a[1]}, {a[0]})"
    raise KeyError(op)


def _operand_name(ref):
    return {"a": "x", "s": "s", "t": "t"}[ref[0]] + str(ref[1])


def relaxed_pow_modes(prog):
    """{instr index: mode} for float32 `x ** s` with a scalar exponent of -1.5 / -0.5.  Inside a
    contraction only the reduced sum is observable (rtol 1e-5), so these become one MUFU.RSQ
    (relative error 2^-22.4) without the per-element exponent test; the exponent VALUE is part of
    the kernel key of the callers that use this."""
    out = {}
    for k, (op, loop, out_dt, args) in enumerate(prog.instrs):
        if op == "power" and loop[0] == np.float32 and args[1][0] == "s":
            v = float(prog.scalars[args[1][1]][0])
            if v == -1.5:
                out[k] = "rsqrt3"
            elif v == -0.5:
                out[k] = "rsqrt"
    return out


def emit_body(prog, fast=False, relaxed=False, pow_modes=None):
    """The fused scalar body: `const T tK = expr;` lines over x<i> (arrays), s<j> (scalars).
    ``relaxed``: inside a contraction, where only the reduced result is observable (rtol bar),
    float32 pow with a uniform exponent may take its reciprocal-square-root form."""
    lines = []
    for k, (op, loop, out_dt, args) in enumerate(prog.instrs):
        exprs = [_operand_name(r) for r in args]
        dts = [prog.dtypes[r] for r in args]
        lines.append(f"const {ctype(out_dt)} t{k} = "
                     f"{emit_expr(op, loop, out_dt, exprs, dts, fast, relaxed, (pow_modes or {}).get(k))};")
    return lines


# --------------------------------------------------------------------------- reductions
_RED = {"sum": "DrSum", "prod": "DrProd", "max": "DrMax", "min": "DrMin"}


def acc_dtype(op, dt):
    """Accumulator type: float32 sums accumulate in double (error far below rtol 1e-5 at 2^30
    elements); everything else accumulates in its own type."""
    dt = np.dtype(dt)
    if op in ("sum", "prod") and dt == np.float32:
        return np.dtype(np.float64)
    return dt


def _identity(op, dt):
    dt = np.dtype(dt)
    c = ctype(dt)
    if op == "sum":
        return f"(({c})0)"
    if op == "prod":
        return f"(({c})1)"
    if dt.kind == "f":
        return f"(({c})(-DR_INF))" if op == "max" else f"(({c})DR_INF)"
    if dt.kind == "b":
        return "false" if op == "max" else "true"
    info = np.iinfo(dt)
    v = info.min if op == "max" else info.max
    if dt.kind == "i" and dt.itemsize == 8 and op == "max":
        return "((long long)(-9223372036854775807LL - 1))"
    suffix = "ULL" if dt.kind == "u" and dt.itemsize == 8 else ("LL" if dt.itemsize == 8 else "")
    return f"(({c}){v}{suffix})"


# --------------------------------------------------------------------------- flat family
def gen_flat(name, prog, in_class, out_dts, vec_ok, stream, reduce=None, unroll=None,
             threads=None, min_blocks=None, meta=None, sclasses=None):
    """Contiguous 1-d kernel.  ``reduce`` = None or (op, acc np.dtype, result np.dtype, post).

    One copy of the fused body per vector lane (plus one scalar-tail copy): each thread loads
    U vectors of V elements (all loads first), evaluates them, stores them; U = 1 for heavy
    bodies (their own arithmetic hides the memory latency and the loop has to fit the
    instruction cache), up to 4 for streaming bodies (more bytes in flight per thread).
    Float32 '/', sqrt and log run through their branch-free fast forms; a vector that raised
    the exception flag is recomputed once through the precise forms."""
    arrays, scalars = prog.arrays, prog.scalars
    item_sizes = [a.dtype.itemsize for a, c in zip(arrays, in_class) if c == "c"]
    item_sizes += [np.dtype(d).itemsize for d in out_dts] if reduce is None else []
    widest = max(item_sizes) if item_sizes else 4
    V = max(1, 16 // widest) if vec_ok else 1
    n_stream = sum(1 for c in in_class if c == "c") + (len(out_dts) if reduce is None else 0)
    if unroll is None:
        unroll = int(os.environ.get("DR_UNROLL", 0)) or (
            1 if body_weight(prog) >= 2 else (4 if n_stream <= 2 else 2))
    U = unroll
    S = "true" if stream else "false"
    two_tier = has_fast_path(prog)
    lockstep = lockstep_ok(prog, V) and os.environ.get("DR_LOCKSTEP", "1") != "0"
    c_inputs = [(i, a) for i, (a, c) in enumerate(zip(arrays, in_class)) if c == "c"]
    staged = (U == 1 and V > 1 and body_weight(prog) >= 2 and c_inputs
              and all(a.dtype.itemsize * V == 16 for _, a in c_inputs)
              and os.environ.get("DR_STAGED", "1") != "0")
    # table-driven erf in a staged kernel: ONE CTA per SM, so that 16 bank-private replicas of
    # the table (84 KiB) fit beside the per-warp operand rings.  512 threads with four vectors
    # per lane per stage: the kernel no longer depends on occupancy (same time at 512 .. 1024
    # threads) and with 128 registers available the compiler stops re-materialising constants
    # (422 instead of 448 instructions per vector), which matters under the power cap.
    erf_rep = 16 if (staged and lockstep and GEN2 and uses_erf_table(prog)
                     and os.environ.get("DR_ERF_REP", "16") != "1") else 1
    if threads is None:
        threads = int(os.environ.get("DR_THREADS", 0)) or (512 if erf_rep == 16 else 256)
    # vectors per lane per stage: a stage of VPL tiles is filled by ONE bulk copy per operand and
    # consumed by VPL trips of the (not unrolled) inner loop, so the barrier wait, the tile
    # arithmetic and the TMA issue are paid once per VPL vectors.  With the 84 KiB erf table the
    # ring is a single 4-tile stage per warp (the refill overlaps the last vector's arithmetic
    # and the other warps hide the rest of the DRAM latency).
    VPL = int(os.environ.get("DR_VPL", 0)) or (4 if erf_rep == 16 else 2)
    NS = int(os.environ.get("DR_STAGES", 0)) or (1 if erf_rep == 16 else 2)

    params = ["const i64 n"]
    for i, a in enumerate(arrays):
        params.append(f"const {ctype(a.dtype)}* __restrict__ in{i}")
    for j, (_, dt) in enumerate(scalars):
        params.append(f"const {ctype(dt)} s{j}")
    if reduce is None:
        for o, dt in enumerate(out_dts):
            params.append(f"{ctype(dt)}* __restrict__ out{o}")
    else:
        rop, acc_dt, res_dt, post = reduce
        A = ctype(acc_dt)
        params += [f"{A}* __restrict__ partials", "unsigned int* __restrict__ counter",
                   f"{ctype(res_dt)}* __restrict__ result", "const double post_scale"]

    if lockstep:
        two_tier = two_tier or has_lane_fast(prog)
    safe_body = emit_body(prog, fast=False)
    fast_body = emit_body(prog, fast=True) if two_tier else safe_body
    # generation 3 of erf wherever a flat lockstep kernel uses the table: carved out of dynamic
    # shared memory behind the operand rings when staged, a static 32 KiB array otherwise
    erf3 = ERF3 and (erf_rep == 16 or (lockstep and GEN2 and V == 4 and uses_erf_table(prog)
                                       and os.environ.get("DR_ERF3_UNSTAGED", "1") != "0"))
    if lockstep:
        lock_body, lock_uniform = emit_body_lockstep(prog, in_class, V, sclasses, erf_rep, erf3=erf3)
    src = []
    w = src.append
    if erf3:
        w(ERF3_SRC)
    if two_tier:
        fp = [f"const {ctype(a.dtype)} x{i}" for i, a in enumerate(arrays)]
        fp += [f"const {ctype(dt)} s{j}" for j, (_, dt) in enumerate(scalars)]
        res_types = [ctype(dt) for dt in out_dts] if reduce is None else [A]
        w(f"struct {name}_res {{ " + " ".join(f"{t} o{o};" for o, t in enumerate(res_types)) + " };")
        w(f"__device__ __noinline__ {name}_res {name}_safe({', '.join(fp)}) {{")
        for line in safe_body:
            w(f"  {line}")
        w(f"  {name}_res res;")
        if reduce is None:
            for o, (r, dt) in enumerate(zip(prog.roots, out_dts)):
                w(f"  res.o{o} = {_store_expr(prog, r, dt)};")
        else:
            w(f"  res.o0 = ({A}){_operand_name(prog.roots[0])};")
        w("  return res;")
        w("}")
    min_blocks = min_blocks or int(os.environ.get("DR_MINBLOCKS", 0)) or (
        (1 if erf_rep == 16 else 1024 // threads) if (lockstep and body_weight(prog) >= 2) else None)
    lb = f"__launch_bounds__({threads}" + (f", {min_blocks})" if min_blocks else ")")
    w(f'extern "C" __global__ void {lb} {name}({", ".join(params)}) {{')
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "b":
            w(f"  const {ctype(a.dtype)} x{i} = in{i}[0];")
    if reduce is not None:
        w(f"  {A} acc[{U}];")
        w(f"#pragma unroll\n  for (int u = 0; u < {U}; ++u) acc[u] = {_identity(rop, acc_dt)};")
    if lockstep and uses_erf_table(prog):
        if GEN2 and erf_rep == 16:
            pass            # carved out of dynamic shared memory behind the operand rings (below)
        elif erf3:
            w("  __shared__ __align__(16) unsigned char dr_erf_tab[DR_ERF3_SMEM_BYTES];")
            w("  dr_erf3_tab_stage(dr_erf_tab);")
        elif GEN2:
            w("  __shared__ float2 dr_erf_tab[3 * DR_ERF2_ROWS];")
            w("  dr_erf2_tab_stage<1>(dr_erf_tab);")
        else:
            w("  __shared__ float2 dr_erf_tab[DR_ERF_TAB_PAIRS];")
            w("  dr_erf_tab_stage(dr_erf_tab);")
    if lockstep:
        emit_explog_stage(w, prog)
        w(DR_ONE.format("n"))
    w(f"  const i64 nv = n / {V};")
    w("  const i64 stride = (i64)gridDim.x * blockDim.x;")

    def element(body, p, target):
        """One element: bind inputs, run `body`, deliver the root(s) to `target`."""
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            if c == "c":
                w(f"{p}const {ctype(a.dtype)} x{i} = v{i}[u].v[e];")
        for line in body:
            w(f"{p}{line}")
        if reduce is None:
            for o, (r, dt) in enumerate(zip(prog.roots, out_dts)):
                w(f"{p}r{o}.v[e] = {_store_expr(prog, r, dt)};")
        else:
            w(f"{p}val[e] = ({A}){_operand_name(prog.roots[0])};")

    prefetch = (not staged) and U == 1 and body_weight(prog) >= 2 \
        and os.environ.get("DR_PREFETCH", "0") != "0"
    ring_bytes = NS * VPL * len(c_inputs) * threads * 16 if staged else 0
    if meta is not None:
        meta["smem"] = ring_bytes + ((ERF3_ROWS * 8 * 16 if erf3 else 3 * ERF2_ROWS * 16 * 8) if erf_rep == 16 else 0)
        meta["threads"] = threads
    if staged:
        # operands arrive through PER-WARP shared-memory rings filled by 1-d TMA bulk copies: lane
        # 0 of each warp issues the copies of the warp's tile (32 vectors = 512 B per operand) and
        # an mbarrier per (warp, stage) counts the bytes.  No block-wide barrier in the loop: the
        # warps of a CTA drift apart, so their divisions (MUFU), polynomials (FMA pipe) and table
        # look-ups (LSU) overlap instead of arriving at every pipe in phase; NS tiles per warp are
        # in flight while one is evaluated, and no register is held for them.
        nin = len(c_inputs)
        WPB = threads // 32
        TB = 512 * VPL                                   # bytes per operand per stage
        SB = nin * TB                                    # bytes per stage
        w("  extern __shared__ __align__(128) unsigned char dr_smem[];")
        if erf3 and erf_rep == 16:
            w(f"  unsigned char* const dr_erf_tab = dr_smem + {ring_bytes};")
            w("  dr_erf3_tab_stage(dr_erf_tab);")
        elif erf_rep == 16:
            w(f"  float2* const dr_erf_tab = reinterpret_cast<float2*>(dr_smem + {ring_bytes});")
            w("  dr_erf2_tab_stage<16>(dr_erf_tab);")
        w(f"  __shared__ __align__(8) unsigned long long dr_bar[{WPB * NS}];")
        # the shuffle tells the compiler the warp index is warp-uniform: everything derived from
        # it (addresses, byte counts, barrier words) stays on the uniform datapath
        w("  const int dr_warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);")
        w("  const int dr_lane = threadIdx.x & 31;")
        w(f"  const unsigned dr_bar_s = dr_smem_addr(&dr_bar[dr_warp * {NS}]);")
        w(f"  const unsigned dr_ring_s = dr_smem_addr(dr_smem) + dr_warp * {NS * SB};")
        w(f"  if (dr_lane == 0) {{ for (int s = 0; s < {NS}; ++s) dr_mbar_init(&dr_bar[dr_warp * {NS} + s], 1); dr_fence_barrier_init(); }}")
        w("  __syncwarp();")
        # tiles are counted in 32 bits: n / (32 * V) < 2^32 for any array that fits in HBM
        w(f"  const unsigned ntiles = (unsigned)((nv + {32 * VPL - 1}) / {32 * VPL});")
        w(f"  const unsigned dr_last_bytes = (unsigned)(nv - (i64)(ntiles - 1) * {32 * VPL}) * 16u;")
        w(f"  const unsigned dr_gw = blockIdx.x * {WPB}u + dr_warp, dr_nw = gridDim.x * {WPB}u;")
        w("  auto dr_issue = [&](unsigned tile, unsigned stage) {       // one elected lane, tile < ntiles")
        w(f"    const unsigned bytes = tile == ntiles - 1 ? dr_last_bytes : {TB}u;")
        w("    const unsigned bar = dr_bar_s + stage * 8u;")
        w(f"    const unsigned dst = dr_ring_s + stage * {SB}u;")
        w(f"    dr_mbar_expect_tx_s(bar, bytes * {nin}u);")
        for slot, (i, a) in enumerate(c_inputs):
            w(f"    dr_bulk_load_s(dst + {slot * TB}u, in{i} + (i64)tile * {32 * VPL * V}, bytes, bar);")
        w("  };")
        w(f"  if (dr_lane == 0) {{ for (unsigned k = 0; k < {NS}u; ++k) if (dr_gw + k * dr_nw < ntiles) dr_issue(dr_gw + k * dr_nw, k); }}")
        w("  unsigned dr_stage = 0, dr_phase = 0;")
        w("  for (unsigned tile = dr_gw; tile < ntiles; tile += dr_nw) {")
        w("    dr_mbar_wait_s(dr_bar_s + dr_stage * 8u, dr_phase);")
        w("    const unsigned dr_cur = dr_stage;")
        w(f"    if (++dr_stage == {NS}u) {{ dr_stage = 0; dr_phase ^= 1u; }}")
        if VPL > 1:
            w("#pragma unroll 1")
            w(f"    for (unsigned dr_sub = 0; dr_sub < {VPL}u; ++dr_sub) {{")
            w(f"    const i64 i = ((i64)tile * {VPL} + dr_sub) * 32 + dr_lane;")
            w(f"    const unsigned dr_src = dr_ring_s + dr_cur * {SB}u + dr_sub * 512u + dr_lane * 16u;")
        else:
            w("    const i64 i = (i64)tile * 32 + dr_lane;")
            w(f"    const unsigned dr_src = dr_ring_s + dr_cur * {SB}u + dr_lane * 16u;")
        for slot, (i, a) in enumerate(c_inputs):
            w(f"    Vec<{ctype(a.dtype)}, {V}> v{i}[1];")
            w(f"    v{i}[0] = dr_lds16<{ctype(a.dtype)}, {V}>(dr_src + {slot * TB}u);")
        last = f"dr_sub == {VPL - 1}u && " if VPL > 1 else ""
        w(f"    if ({last}true) {{")
        w("      __syncwarp();           // every lane holds its last vector of the stage: refill it")
        w(f"      if (tile + {NS}u * dr_nw < ntiles) {{ if (dr_elect()) dr_issue(tile + {NS}u * dr_nw, dr_cur); }}")
        w("    }")
        w("    if (i < nv) {")
    elif prefetch:
        # software pipelining: the loads of the NEXT vector are in flight while this one is
        # evaluated (a heavy body with one vector per trip would otherwise expose DRAM latency)
        w("  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;")
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            if c == "c":
                w(f"  Vec<{ctype(a.dtype)}, {V}> v{i}[1], nx{i};")
        w("  if (i < nv) {")
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            if c == "c":
                w(f"    v{i}[0] = dr_ld<{S}, {ctype(a.dtype)}, {V}>(in{i} + i * {V});")
        w("  }")
        w("  for (; i < nv; i += stride) {")
        w("    const i64 inx = i + stride < nv ? i + stride : i;")
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            if c == "c":
                w(f"    nx{i} = dr_ld<{S}, {ctype(a.dtype)}, {V}>(in{i} + inx * {V});")
    else:
        w(f"  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += {U} * stride) {{")
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            if c == "c":
                w(f"    Vec<{ctype(a.dtype)}, {V}> v{i}[{U}];")
        w("#pragma unroll")
        w(f"    for (int u = 0; u < {U}; ++u) {{")
        w("      const i64 iu = i + u * stride;")
        w("      const i64 ic = iu < nv ? iu : nv - 1;       // idle slots re-read a valid vector")
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            if c == "c":
                w(f"      v{i}[u] = dr_ld<{S}, {ctype(a.dtype)}, {V}>(in{i} + ic * {V});")
        w("    }")
    w("#pragma unroll")
    w(f"    for (int u = 0; u < {U}; ++u) {{")
    if reduce is None:
        for o, dt in enumerate(out_dts):
            w(f"      Vec<{ctype(dt)}, {V}> r{o};")
    else:
        w(f"      {A} val[{V}];")
    if two_tier:
        w("      bool bad = false;")
    if lockstep:
        for line in lock_body:
            w(f"      {line}")
        w("#pragma unroll")
        w(f"      for (int e = 0; e < {V}; ++e) {{")
        for o, (r, dt) in enumerate(zip(prog.roots, out_dts) if reduce is None else []):
            w(f"        r{o}.v[e] = {_lane_store(prog, r, dt, in_class, lock_uniform)};")
        if reduce is not None:
            w(f"        val[e] = ({A}){_lane_name(prog.roots[0], in_class, lock_uniform)};")
        w("      }")
    else:
        w("#pragma unroll")
        w(f"      for (int e = 0; e < {V}; ++e) {{")
        element(fast_body, "        ", None)
        w("      }")
    if two_tier:
        w("      if (bad) {                                 // rare: precise re-evaluation")
        w("#pragma unroll")
        w(f"        for (int e = 0; e < {V}; ++e) {{")
        call_in = ", ".join([f"v{i}[u].v[e]" if c == "c" else f"x{i}"
                             for i, c in enumerate(in_class)] + [f"s{j}" for j in range(len(scalars))])
        w(f"          const {name}_res res = {name}_safe({call_in});")
        if reduce is None:
            for o in range(len(out_dts)):
                w(f"          r{o}.v[e] = res.o{o};")
        else:
            w("          val[e] = res.o0;")
        w("        }")
        w("      }")
    if reduce is None:
        guard = "" if U == 1 else "if (i + u * stride < nv) "
        for o, dt in enumerate(out_dts):
            w(f"      {guard}dr_st<{S}, {ctype(dt)}, {V}>(out{o} + (i + u * stride) * {V}, r{o});")
    else:
        w(f"      if (i + u * stride < nv) {{")
        w("#pragma unroll")
        w(f"        for (int e = 0; e < {V}; ++e) acc[u] = {_RED[rop]}::op(acc[u], val[e]);")
        w("      }")
    w("    }")
    if prefetch:
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            if c == "c":
                w(f"    v{i}[0] = nx{i};")
    if staged:
        w("    }")
        if VPL > 1:
            w("    }")
    w("  }")
    # scalar tail: the n - nv*V < V trailing elements, first threads of block 0, precise forms
    if V > 1:
        w(f"  if (blockIdx.x == 0 && threadIdx.x < (unsigned)(n - nv * {V})) {{")
        w(f"    const i64 j = nv * {V} + threadIdx.x;")
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            if c == "c":
                w(f"    const {ctype(a.dtype)} x{i} = in{i}[j];")
        for line in safe_body:
            w(f"    {line}")
        if reduce is None:
            for o, (r, dt) in enumerate(zip(prog.roots, out_dts)):
                w(f"    out{o}[j] = {_store_expr(prog, r, dt)};")
        else:
            w(f"    acc[0] = {_RED[rop]}::op(acc[0], ({A}){_operand_name(prog.roots[0])});")
        w("  }")
    if reduce is not None:
        _emit_reduce_finish(w, rop, acc_dt, res_dt, U)
    w("}")
    return "\n".join(src) + "\n"


_PACKED = {"add": "dr_add4", "subtract": "dr_sub4", "multiply": "dr_mul4"}
_NOFUSE = {"add": "dr_add4_nofuse", "subtract": "dr_sub4_nofuse"}
# second generation: accurate-table erf, guards placed by ranges.analyse
GEN2 = os.environ.get("DR_GEN", "2") != "1"
_LANE4_FAST = {"true_divide": "dr_div4_fast", "divide": "dr_div4_fast", "sqrt": "dr_sqrt4_fast",
               "log": "dr_log4_f32", "exp": "dr_exp4_f32", "erf": "dr_erf4_tab"}
_LANE4_R = {"true_divide": "dr_div4_r", "divide": "dr_div4_r", "sqrt": "dr_sqrt4_r",
            "log": "dr_log4_t", "exp": "dr_exp4_t", "erf": "dr_erf4_gal"}
if os.environ.get("DR_EXPLOG_POLY"):         # table-free exp / log (17 / 30 FMA cycles per element)
    _LANE4_R.update({"log": "dr_log4_r", "exp": "dr_exp4_r"})
_TABLE_ARG = {"dr_erf4_gal": ", dr_erf_tab", "dr_exp4_t": ", dr_exp_tab", "dr_log4_t": ", dr_log_tab"}
if os.environ.get("DR_F64_EXPLOG"):          # previous generation: double-precision exp/log/erf
    _LANE4_FAST.update({"log": "dr_log4_fast", "exp": "dr_exp4_fast", "erf": "dr_erf4_fast"})
    GEN2 = False
F32 = np.dtype(np.float32)
if os.environ.get("DR_F32_NATIVE"):
    _LANE4_FAST.update({"log": "dr_log4_native", "exp": "dr_exp4_native", "erf": "dr_erf4_native"})
    GEN2 = False


def explog_tables(prog):
    """(uses exp table, uses log table) for the lane forms of generation 2."""
    if not GEN2:
        return False, False
    ops = {op for op, loop, _, _ in prog.instrs if loop[0] == np.float32}
    return ("exp" in ops and _LANE4_R["exp"] == "dr_exp4_t",
            "log" in ops and _LANE4_R["log"] == "dr_log4_t")


def emit_explog_stage(w, prog, indent="  "):
    ue, ul = explog_tables(prog)
    if ue:
        w(f"{indent}__shared__ float2 dr_exp_tab[DR_EXP2_SMEM_PAIRS];")
    if ul:
        w(f"{indent}__shared__ float4 dr_log_tab[DR_LOG2_SMEM_QUADS];")
    if ue or ul:
        w(f"{indent}dr_explog_tab_stage({'dr_exp_tab' if ue else 'nullptr'}, {'dr_log_tab' if ul else 'nullptr'});")


def uses_erf_table(prog):
    return (GEN2 or _LANE4_FAST["erf"] == "dr_erf4_tab") and any(
        op == "erf" and loop[0] == np.float32 for op, loop, _, _ in prog.instrs)


# 1.0f the compiler cannot see through (derived from a kernel parameter whose sign bit is never
# set; a form based on blockDim.x is folded through __launch_bounds__): fma(a, one, b) == a + b
# exactly and is never contracted with the multiply that produced a
DR_ONE = "  const float dr_one = __int_as_float(0x3f800000 | (int)((unsigned long long){0} >> 63));"


def lockstep_ok(prog, V):
    """The 4-lane lockstep form applies to float32 vectors of four."""
    return V == 4


def has_lane_fast(prog):
    return any(op in _LANE4_FAST and loop[0] == F32 for op, loop, _, _ in prog.instrs)


def emit_body_lockstep(prog, in_class, V=4, sclasses=None, erf_rep=1, erf3=False):
    """Lane-array form of the fused body: every SSA value is `T tK[4]` (or a plain scalar when
    it only depends on scalars / broadcast operands) and each instruction is applied to all
    four lanes at once -- packed f32x2 for float32 + - *, the dr_*4_fast lane functions for
    / sqrt log exp erf, an unrolled lane loop for everything else."""
    lines = []
    uniform = {}
    packed_products = set()          # temps produced by a packed multiply
    an = ranges.analyse(prog, in_class, sclasses, set(_LANE4_R)) if (GEN2 and V == 4) else None
    if an is not None and (an.pos_inputs or an.any_inputs):
        # one min/max tree over every lane of every operand the fast forms rely on
        tests = [f"dr_rg.pos4(v{i}[u].v);" for i in an.pos_inputs]
        tests += [f"dr_rg.any4(v{i}[u].v);" for i in an.any_inputs]
        oks = (["dr_rg.ok_pos()"] if an.pos_inputs else []) + (["dr_rg.ok_any()"] if an.any_inputs else [])
        lines.append("{ DrRange dr_rg; " + " ".join(tests) + f" bad = bad || !({' && '.join(oks)}); }}")

    def is_uniform(r):
        if r[0] == "s":
            return True
        if r[0] == "a":
            return in_class[r[1]] == "b"
        return uniform[r]

    def arr(r):                      # array (or scalar) expression naming the operand
        if r[0] == "a":
            return f"x{r[1]}" if in_class[r[1]] == "b" else f"v{r[1]}[u].v"
        return _operand_name(r)

    def lane(r):
        return arr(r) if is_uniform(r) else f"{arr(r)}[e]"

    for k, (op, loop, out_dt, args) in enumerate(prog.instrs):
        me = ("t", k)
        dts = [prog.dtypes[r] for r in args]
        T = ctype(out_dt)
        if all(is_uniform(r) for r in args):
            uniform[me] = True
            lines.append(f"const {T} t{k} = {emit_expr(op, loop, out_dt, [arr(r) for r in args], dts)};")
            continue
        uniform[me] = False
        lines.append(f"{T} t{k}[{V}];")
        same = all(d == F32 for d in dts) and all(d == F32 for d in loop) and out_dt == F32
        if same and op in ("add", "subtract") and V == 4 and any(r in packed_products for r in args):
            # ptxas fuses mul.rn.f32x2 + add/sub.rn.f32x2 into FFMA2 even though both carry an
            # explicit .rn (observed, CUDA 12.9; the scalar forms are never fused).  An add
            # that consumes a packed product therefore stays scalar: a*b+c must round twice.
            if GEN2:
                lines.append(f"{_NOFUSE[op]}({arr(args[0])}, {arr(args[1])}, t{k}, dr_one);")
            else:
                fn = "__fadd_rn" if op == "add" else "__fsub_rn"
                lines.append(f"_Pragma(\"unroll\") for (int e = 0; e < {V}; ++e) "
                             f"t{k}[e] = {fn}({lane(args[0])}, {lane(args[1])});")
        elif same and op in _PACKED and V == 4:
            lines.append(f"{_PACKED[op]}({arr(args[0])}, {arr(args[1])}, t{k});")
            if op == "multiply":
                packed_products.add(me)
        elif same and an is not None and op in _LANE4_R and k in an.check:
            flags = ", ".join(("true" if c else "false") if isinstance(c, bool) else str(int(c))
                              for c in an.check[k])
            fn = _LANE4_R[op]
            if op == "erf" and erf3:
                fn = "dr_erf4_s"
            elif op == "erf":
                flags += f", {erf_rep}"
            extra = _TABLE_ARG.get(_LANE4_R[op], "")
            lines.append(f"{fn}<{flags}>({', '.join(arr(r) for r in args)}, t{k}, bad{extra});")
        elif same and op in _LANE4_FAST and V == 4:
            extra = ", dr_erf_tab" if _LANE4_FAST[op] == "dr_erf4_tab" else ""
            lines.append(f"{_LANE4_FAST[op]}({', '.join(arr(r) for r in args)}, t{k}, bad{extra});")
        else:
            expr = emit_expr(op, loop, out_dt, [lane(r) for r in args], dts)
            lines.append(f"_Pragma(\"unroll\") for (int e = 0; e < {V}; ++e) t{k}[e] = {expr};")
    return lines, uniform


def _lane_name(ref, in_class, uniform):
    if ref[0] == "s":
        return _operand_name(ref)
    if ref[0] == "a":
        return f"x{ref[1]}" if in_class[ref[1]] == "b" else f"v{ref[1]}[u].v[e]"
    return _operand_name(ref) if uniform[ref] else f"{_operand_name(ref)}[e]"


def _lane_store(prog, ref, out_dt, in_class, uniform):
    have = prog.dtypes[ref]
    nm = _lane_name(ref, in_class, uniform)
    if have == out_dt:
        return nm
    if np.dtype(out_dt).kind == "b":
        return f"({nm} != 0)"
    return f"(({ctype(out_dt)})({nm}))"


def _store_expr(prog, ref, out_dt):
    have = prog.dtypes[ref]
    nm = _operand_name(ref)
    if have == out_dt:
        return nm
    if np.dtype(out_dt).kind == "b":
        return f"({nm} != 0)"
    return f"(({ctype(out_dt)})({nm}))"


def _emit_reduce_finish(w, rop, acc_dt, res_dt, U):
    A = ctype(acc_dt)
    ident = _identity(rop, acc_dt)
    w(f"  {A} total = acc[0];")
    if U > 1:
        w(f"#pragma unroll\n  for (int u = 1; u < {U}; ++u) total = {_RED[rop]}::op(total, acc[u]);")
    w(f"  __shared__ {A} scratch[32];")
    w(f"  total = dr_block_reduce<{_RED[rop]}>(total, {ident}, scratch);")
    w(f"  __shared__ {A} final_value;")
    w(f"  if (dr_grid_reduce<{_RED[rop]}>(total, {ident}, partials, counter, scratch, &final_value)) {{")
    if np.dtype(acc_dt).kind == "f":
        w(f"    *result = ({ctype(res_dt)})(final_value / ({A})post_scale);")
    else:
        w(f"    *result = ({ctype(res_dt)})final_value;")
    w("  }")


# --------------------------------------------------------------------------- nd family
def gen_nd(name, prog, ndim, in_class, out_dts, reduce=None, threads=256, wide_index=False, vec=0,
           sclasses=None):
    """General strided/broadcast kernel.  Geometry arrives in one struct argument:
    shape[ndim], then byte strides per operand per dim (inputs then outputs).
    ``vec`` = V > 1: the innermost dimension is walked in vectors of V elements (planner.
    _try_inner_vectors): operands of class 'v' are read with one 128-bit load, class 'i' (no
    movement along the inner dimension) with one scalar load shared by the V lanes, outputs are
    written with one 128-bit store; the index decomposition is paid once per V elements."""
    if vec and vec > 1 and reduce is None:
        return _gen_nd_vec(name, prog, ndim, in_class, out_dts, vec, threads, wide_index, sclasses)
    arrays, scalars = prog.arrays, prog.scalars
    n_ops = len(arrays) + (len(out_dts) if reduce is None else 0)
    I = "i64" if wide_index else "u32"
    src = []
    w = src.append
    w(f"struct Geo_{name} {{ i64 total; i64 shape[{ndim}]; i64 stride[{max(n_ops, 1)}][{ndim}]; }};")
    params = [f"const Geo_{name} g"]
    for i, a in enumerate(arrays):
        params.append(f"const char* __restrict__ in{i}")
    for j, (_, dt) in enumerate(scalars):
        params.append(f"const {ctype(dt)} s{j}")
    if reduce is None:
        for o, dt in enumerate(out_dts):
            params.append(f"char* __restrict__ out{o}")
    else:
        rop, acc_dt, res_dt, post = reduce
        A = ctype(acc_dt)
        params += [f"{A}* __restrict__ partials", "unsigned int* __restrict__ counter",
                   f"{ctype(res_dt)}* __restrict__ result", "const double post_scale"]
    body = emit_body(prog)
    w(f'extern "C" __global__ void __launch_bounds__({threads}) {name}({", ".join(params)}) {{')
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "b":
            w(f"  const {ctype(a.dtype)} x{i} = *reinterpret_cast<const {ctype(a.dtype)}*>(in{i});")
    if reduce is not None:
        w(f"  {A} acc[1]; acc[0] = {_identity(rop, acc_dt)};")
    w(f"  const {I} n_total = ({I})g.total;")
    w(f"  const {I} step = ({I})gridDim.x * blockDim.x;")
    w(f"  for ({I} idx = ({I})blockIdx.x * blockDim.x + threadIdx.x; idx < n_total; idx += step) {{")
    w(f"    {I} rem = idx;")
    w(f"    i64 off[{max(n_ops, 1)}];")
    w(f"#pragma unroll\n    for (int k = 0; k < {max(n_ops, 1)}; ++k) off[k] = 0;")
    w(f"#pragma unroll\n    for (int d = {ndim - 1}; d >= 0; --d) {{")
    w(f"      const {I} extent = ({I})g.shape[d];")
    w(f"      const {I} q = d ? rem / extent : 0;")
    w(f"      const {I} c = d ? rem - q * extent : rem;")
    w("      rem = q;")
    w(f"#pragma unroll\n      for (int k = 0; k < {max(n_ops, 1)}; ++k) off[k] += (i64)c * g.stride[k][d];")
    w("    }")
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c != "b":
            w(f"    const {ctype(a.dtype)} x{i} = *reinterpret_cast<const {ctype(a.dtype)}*>(in{i} + off[{i}]);")
    for line in body:
        w(f"    {line}")
    if reduce is None:
        for o, (r, dt) in enumerate(zip(prog.roots, out_dts)):
            w(f"    *reinterpret_cast<{ctype(dt)}*>(out{o} + off[{len(arrays) + o}]) = {_store_expr(prog, r, dt)};")
    else:
        w(f"    acc[0] = {_RED[rop]}::op(acc[0], ({A}){_operand_name(prog.roots[0])});")
    w("  }")
    if reduce is not None:
        _emit_reduce_finish(w, rop, acc_dt, res_dt, 1)
    w("}")
    return "\n".join(src) + "\n"


def _gen_nd_vec(name, prog, ndim, in_class, out_dts, V, threads, wide_index, sclasses=None):
    arrays, scalars = prog.arrays, prog.scalars
    n_ops = len(arrays) + len(out_dts)
    I = "i64" if wide_index else "u32"
    src = []
    w = src.append
    w(f"struct Geo_{name} {{ i64 total; i64 shape[{ndim}]; i64 stride[{max(n_ops, 1)}][{ndim}]; }};")
    params = [f"const Geo_{name} g"]
    for i, a in enumerate(arrays):
        params.append(f"const char* __restrict__ in{i}")
    for j, (_, dt) in enumerate(scalars):
        params.append(f"const {ctype(dt)} s{j}")
    for o, dt in enumerate(out_dts):
        params.append(f"char* __restrict__ out{o}")
    body = emit_body(prog)
    # float32 vectors of four: the lockstep body of the flat family (packed f32x2 arithmetic,
    # branch-free guarded / sqrt log exp erf, precise re-evaluation of a flagged vector)
    f32 = all(a.dtype == F32 for a in arrays) and all(np.dtype(d) == F32 for d in out_dts)
    lock = V == 4 and f32 and os.environ.get("DR_LOCKSTEP", "1") != "0"
    if lock:
        lane_class = tuple("c" if c == "v" else "b" for c in in_class)
        lock_body, lock_uniform = emit_body_lockstep(prog, lane_class, V, sclasses, 1)
    w(f'extern "C" __global__ void __launch_bounds__({threads}) {name}({", ".join(params)}) {{')
    if lock:
        emit_explog_stage(w, prog)
        if uses_erf_table(prog):
            if GEN2:
                w("  __shared__ float2 dr_erf_tab[3 * DR_ERF2_ROWS];")
                w("  dr_erf2_tab_stage<1>(dr_erf_tab);")
            else:
                w("  __shared__ float2 dr_erf_tab[DR_ERF_TAB_PAIRS];")
                w("  dr_erf_tab_stage(dr_erf_tab);")
        w(DR_ONE.format("g.total"))
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "b":
            w(f"  const {ctype(a.dtype)} x{i} = *reinterpret_cast<const {ctype(a.dtype)}*>(in{i});")
    w(f"  const {I} n_total = ({I})g.total;")
    w(f"  const {I} step = ({I})gridDim.x * blockDim.x;")
    w(f"  for ({I} idx = ({I})blockIdx.x * blockDim.x + threadIdx.x; idx < n_total; idx += step) {{")
    w(f"    {I} rem = idx;")
    w(f"    i64 off[{max(n_ops, 1)}];")
    w(f"#pragma unroll\n    for (int k = 0; k < {max(n_ops, 1)}; ++k) off[k] = 0;")
    w(f"#pragma unroll\n    for (int d = {ndim - 1}; d >= 0; --d) {{")
    w(f"      const {I} extent = ({I})g.shape[d];")
    w(f"      const {I} q = d ? rem / extent : 0;")
    w(f"      const {I} c = d ? rem - q * extent : rem;")
    w("      rem = q;")
    w(f"#pragma unroll\n      for (int k = 0; k < {max(n_ops, 1)}; ++k) off[k] += (i64)c * g.stride[k][d];")
    w("    }")
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        T = ctype(a.dtype)
        if c == "v":
            w(f"    Vec<{T}, {V}> v{i}[1];")
            w(f"    v{i}[0] = dr_ld<false, {T}, {V}>(reinterpret_cast<const {T}*>(in{i} + off[{i}]));")
        elif c == "i":
            w(f"    const {T} x{i} = *reinterpret_cast<const {T}*>(in{i} + off[{i}]);")
    for o, dt in enumerate(out_dts):
        w(f"    Vec<{ctype(dt)}, {V}> r{o};")
    if lock:
        w("    constexpr int u = 0;")
        w("    bool bad = false;")
        for line in lock_body:
            w(f"    {line}")
        w(f"#pragma unroll\n    for (int e = 0; e < {V}; ++e) {{")
        for o, (r, dt) in enumerate(zip(prog.roots, out_dts)):
            w(f"      r{o}.v[e] = {_lane_store(prog, r, dt, lane_class, lock_uniform)};")
        w("    }")
        w("    if (bad) {                                   // rare: precise re-evaluation")
    w(f"#pragma unroll\n    for (int e = 0; e < {V}; ++e) {{")
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "v":
            w(f"      const {ctype(a.dtype)} x{i} = v{i}[0].v[e];")
    for line in body:
        w(f"      {line}")
    for o, (r, dt) in enumerate(zip(prog.roots, out_dts)):
        w(f"      r{o}.v[e] = {_store_expr(prog, r, dt)};")
    w("    }")
    if lock:
        w("    }")
    for o, dt in enumerate(out_dts):
        w(f"    dr_st<false, {ctype(dt)}, {V}>(reinterpret_cast<{ctype(dt)}*>(out{o} + off[{len(arrays) + o}]), r{o});")
    w("  }")
    w("}")
    return "\n".join(src) + "\n"


# --------------------------------------------------------------------------- tile family
def gen_tile(name, prog, in_class, out_dts, T=64, W=1, threads=256):
    """Elementwise kernel over a 2-d space (R, C) in which some operands are *transposed* — they
    walk contiguously along R, not along C (`X.T + X`, reference `delayarray.py:523-527` hands
    `.T` to the backend array, so such operands arrive as stride-swapped leaves).  The general
    `nd` kernel reads them a row pitch apart (2.1 TB/s).  Here a CTA owns a T x T tile: operands
    of class 't' are loaded with the threads running along R (coalesced) into a padded shared
    tile, then every thread evaluates the fused body with the threads running along C, reading
    't' operands from shared memory and 'v' operands (contiguous along C) and the outputs
    straight from / to global memory, coalesced.  ``W`` = 2: every thread handles two adjacent
    elements of the contiguous direction with one 2-element vector access (4-byte types: 256
    bytes per warp instruction; the planner checks the alignment).  Geometry: R, C, tiles along C,
    tile count, then byte strides (rows, cols) per operand, inputs then outputs.
    Classes: 'b' one value for the whole space, 't' via the shared tile, 'v' / 's' direct."""
    arrays, scalars = prog.arrays, prog.scalars
    n_ops = len(arrays) + len(out_dts)
    lanes = T // W                       # threads along the contiguous direction
    step = threads // lanes
    passes = T // step
    src = []
    w = src.append
    w(f"struct Geo_{name} {{ i64 R; i64 C; i64 tiles_c; i64 ntiles; i64 stride[{n_ops}][2]; }};")
    params = [f"const Geo_{name} g"]
    for i, a in enumerate(arrays):
        params.append(f"const char* __restrict__ in{i}")
    for j, (_, dt) in enumerate(scalars):
        params.append(f"const {ctype(dt)} s{j}")
    for o, dt in enumerate(out_dts):
        params.append(f"char* __restrict__ out{o}")
    body = emit_body(prog)
    w(f'extern "C" __global__ void __launch_bounds__({threads}) {name}({", ".join(params)}) {{')
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "t":
            w(f"  __shared__ {ctype(a.dtype)} tile{i}[{T}][{T + 1}];")
        elif c == "b":
            w(f"  const {ctype(a.dtype)} x{i} = *reinterpret_cast<const {ctype(a.dtype)}*>(in{i});")
    w(f"  const int lx = (threadIdx.x % {lanes}) * {W}, ly = threadIdx.x / {lanes};")
    w("  for (i64 t = blockIdx.x; t < g.ntiles; t += gridDim.x) {")
    w("    const i64 tr = t / g.tiles_c, tc = t - tr * g.tiles_c;")
    w(f"    const i64 r0 = tr * {T}, c0 = tc * {T};")
    staged = [(i, ctype(a.dtype)) for i, (a, c) in enumerate(zip(arrays, in_class)) if c == "t"]
    direct = [(i, ctype(a.dtype)) for i, (a, c) in enumerate(zip(arrays, in_class)) if c == "v"]
    w("    {                                                  // threads along R: the transposed operands")
    w("      const i64 r = r0 + lx;")
    for i, Tn in staged:
        w(f"      const char* p{i} = in{i} + r * g.stride[{i}][0] + (c0 + ly) * g.stride[{i}][1];")
    if W == 1:
        w(f"#pragma unroll\n      for (int k = 0; k < {passes}; ++k) {{")
        w(f"        if (r < g.R && c0 + ly + k * {step} < g.C) {{")
        for i, Tn in staged:
            w(f"          tile{i}[ly + k * {step}][lx] = *reinterpret_cast<const {Tn}*>"
              f"(p{i} + (i64)(k * {step}) * g.stride[{i}][1]);")
        w("        }")
        w("      }")
    else:
        # every load of the tile is issued before the first shared-memory store (the vector
        # accessors are asm statements the compiler keeps in order)
        for i, Tn in staged:
            w(f"      Vec<{Tn}, {W}> h{i}[{passes}];")
        w(f"#pragma unroll\n      for (int k = 0; k < {passes}; ++k) {{")
        w(f"        if (c0 + ly + k * {step} < g.C) {{")
        for i, Tn in staged:
            at = f"reinterpret_cast<const {Tn}*>(p{i} + (i64)(k * {step}) * g.stride[{i}][1])"
            w(f"          if (r + {W - 1} < g.R) h{i}[k] = dr_ld<true, {Tn}, {W}>({at});")
            w(f"          else if (r < g.R) h{i}[k].v[0] = *{at};")
        w("        }")
        w("      }")
        w(f"#pragma unroll\n      for (int k = 0; k < {passes}; ++k) {{")
        for i, Tn in staged:
            for e in range(W):
                w(f"        tile{i}[ly + k * {step}][lx + {e}] = h{i}[k].v[{e}];")
        w("      }")
    w("    }")
    w("    __syncthreads();")
    w("    {                                                  // threads along C: evaluate and store")
    w("      const i64 c = c0 + lx;")
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c in "vs":
            w(f"      const char* p{i} = in{i} + (r0 + ly) * g.stride[{i}][0] + c * g.stride[{i}][1];")
    for o in range(len(out_dts)):
        k = len(arrays) + o
        w(f"      char* q{o} = out{o} + (r0 + ly) * g.stride[{k}][0] + c * g.stride[{k}][1];")
    if W > 1:
        for i, Tn in direct:
            w(f"      Vec<{Tn}, {W}> vx{i}[{passes}];")
        if direct:
            w(f"#pragma unroll\n      for (int k = 0; k < {passes}; ++k) {{")
            w(f"        if (r0 + ly + k * {step} < g.R) {{")
            for i, Tn in direct:
                at = f"reinterpret_cast<const {Tn}*>(p{i} + (i64)(k * {step}) * g.stride[{i}][0])"
                w(f"          if (c + {W - 1} < g.C) vx{i}[k] = dr_ld<true, {Tn}, {W}>({at});")
                w(f"          else if (c < g.C) vx{i}[k].v[0] = *{at};")
            w("        }")
            w("      }")
    w(f"#pragma unroll\n      for (int k = 0; k < {passes}; ++k) {{")
    w(f"        if (r0 + ly + k * {step} < g.R && c < g.C) {{")
    for o, dt in enumerate(out_dts):
        if W > 1:
            w(f"          Vec<{ctype(dt)}, {W}> vr{o};")
    for e in range(W):
        w("          {" if e == 0 else f"          if (c + {e} < g.C) {{")
        for i, (a, c) in enumerate(zip(arrays, in_class)):
            Tn = ctype(a.dtype)
            if c == "t":
                w(f"            const {Tn} x{i} = tile{i}[lx + {e}][ly + k * {step}];")
            elif c == "v" and W > 1:
                w(f"            const {Tn} x{i} = vx{i}[k].v[{e}];")
            elif c in "vs":
                w(f"            const {Tn} x{i} = *reinterpret_cast<const {Tn}*>(p{i} + (i64)(k * {step}) * "
                  f"g.stride[{i}][0] + (i64){e} * g.stride[{i}][1]);")
        for line in body:
            w(f"            {line}")
        for o, (r, dt) in enumerate(zip(prog.roots, out_dts)):
            if W > 1:
                w(f"            vr{o}.v[{e}] = {_store_expr(prog, r, dt)};")
            else:
                w(f"            *reinterpret_cast<{ctype(dt)}*>(q{o} + (i64)(k * {step}) * "
                  f"g.stride[{len(arrays) + o}][0]) = {_store_expr(prog, r, dt)};")
        w("          }")
    if W > 1:
        for o, dt in enumerate(out_dts):
            at = f"reinterpret_cast<{ctype(dt)}*>(q{o} + (i64)(k * {step}) * g.stride[{len(arrays) + o}][0])"
            w(f"          if (c + {W - 1} < g.C) dr_st<true, {ctype(dt)}, {W}>({at}, vr{o});")
            w(f"          else *{at} = vr{o}.v[0];")
    w("        }")
    w("      }")
    w("    }")
    w("    __syncthreads();")
    w("  }")
    w("}")
    return "\n".join(src) + "\n"


# --------------------------------------------------------------------------- rows family
def _emit_lane_operands(w, arrays, in_class, V, indent):
    """inside the per-lane loop `e`: bind x<i> for every non-broadcast operand"""
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "v":
            w(f"{indent}const {ctype(a.dtype)} x{i} = vx{i}.v[e];")
        elif c == "s":
            w(f"{indent}const {ctype(a.dtype)} x{i} = sx{i}[e];")
        elif c == "i":
            w(f"{indent}const {ctype(a.dtype)} x{i} = ix{i};")


def gen_rows(name, prog, in_class, reduce, mode, threads=256, V=1):
    """Reduce the trailing axis of a (rows, cols) space.  Geometry per operand: byte stride
    along rows and along cols (either may be 0 = broadcast).  mode 'warp': one warp per row
    (short rows), 'block': one block per row (long rows).

    in_class per operand: 'b' one value for the whole space, 'v' contiguous along cols and
    16-byte aligned rows (one 128-bit load per V elements), 'i' constant along cols (one scalar
    load per row pass), 's' any other stride (scalar loads).  V > 1 requires cols % V == 0.
    Every thread keeps V independent accumulators and two trips in flight."""
    arrays, scalars = prog.arrays, prog.scalars
    rop, acc_dt, res_dt, post = reduce
    A = ctype(acc_dt)
    n_ops = max(len(arrays), 1)
    src = []
    w = src.append
    w(f"struct Geo_{name} {{ i64 rows; i64 cols; i64 rs[{n_ops}]; i64 cs[{n_ops}]; i64 out_stride; }};")
    params = [f"const Geo_{name} g"]
    for i, a in enumerate(arrays):
        params.append(f"const char* __restrict__ in{i}")
    for j, (_, dt) in enumerate(scalars):
        params.append(f"const {ctype(dt)} s{j}")
    params += ["char* __restrict__ result", "const double post_scale"]
    body = emit_body(prog)
    ident = _identity(rop, acc_dt)
    w(f'extern "C" __global__ void __launch_bounds__({threads}) {name}({", ".join(params)}) {{')
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "b":
            w(f"  const {ctype(a.dtype)} x{i} = *reinterpret_cast<const {ctype(a.dtype)}*>(in{i});")
    if mode == "warp":
        w("  const int lanes = 32;")
        w("  const i64 row0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;")
        w("  const i64 row_step = ((i64)gridDim.x * blockDim.x) >> 5;")
        w("  const int lane = threadIdx.x & 31;")
    else:
        w("  const int lanes = blockDim.x;")
        w("  const i64 row0 = blockIdx.x;")
        w("  const i64 row_step = gridDim.x;")
        w("  const int lane = threadIdx.x;")
        w(f"  __shared__ {A} scratch[32];")
    w("  for (i64 r = row0; r < g.rows; r += row_step) {")
    w(f"    {A} acc[{V}];")
    w(f"#pragma unroll\n    for (int e = 0; e < {V}; ++e) acc[e] = {ident};")
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "i":
            w(f"    const {ctype(a.dtype)} ix{i} = *reinterpret_cast<const {ctype(a.dtype)}*>(in{i} + r * g.rs[{i}]);")
    w(f"#pragma unroll 2\n    for (i64 c = (i64)lane * {V}; c < g.cols; c += (i64)lanes * {V}) {{")
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        T = ctype(a.dtype)
        if c == "v":
            w(f"      const Vec<{T}, {V}> vx{i} = dr_ld<false, {T}, {V}>(reinterpret_cast<const {T}*>"
              f"(in{i} + r * g.rs[{i}]) + c);")
        elif c == "s":
            w(f"      {T} sx{i}[{V}];")
            w(f"#pragma unroll\n      for (int e = 0; e < {V}; ++e) sx{i}[e] = *reinterpret_cast<const {T}*>"
              f"(in{i} + r * g.rs[{i}] + (c + e) * g.cs[{i}]);")
    w(f"#pragma unroll\n      for (int e = 0; e < {V}; ++e) {{")
    _emit_lane_operands(w, arrays, in_class, V, "        ")
    for line in body:
        w(f"        {line}")
    w(f"        acc[e] = {_RED[rop]}::op(acc[e], ({A}){_operand_name(prog.roots[0])});")
    w("      }")
    w("    }")
    w(f"#pragma unroll\n    for (int e = 1; e < {V}; ++e) acc[0] = {_RED[rop]}::op(acc[0], acc[e]);")
    if mode == "warp":
        w(f"    const {A} tot = dr_warp_reduce<{_RED[rop]}>(acc[0]);")
    else:
        w(f"    const {A} tot = dr_block_reduce<{_RED[rop]}>(acc[0], {ident}, scratch);")
    fin = f"({ctype(res_dt)})(tot / ({A})post_scale)" if np.dtype(acc_dt).kind == "f" \
        else f"({ctype(res_dt)})tot"
    w(f"    if (lane == 0) *reinterpret_cast<{ctype(res_dt)}*>(result + r * g.out_stride) = {fin};")
    w("  }")
    w("}")
    return "\n".join(src) + "\n"


# --------------------------------------------------------------------------- cols family
def gen_cols(name, prog, in_class, reduce, threads=256, V=1, partial=False):
    """Reduce the MIDDLE axis of an (outer, red, inner) space with inner > 1: one thread per V
    consecutive (outer, inner) output elements, coalesced along inner, walking a CHUNK of the
    reduced axis (blockIdx.y selects the chunk) with four trips in flight.

    partial=False: one chunk covers the axis, results are finished (post scale, cast) and stored.
    partial=True: each chunk stores its accumulators to out[(o * splits + chunk) * inner + c] in
    the accumulator type; the same kernel family then reduces the (outer, splits, inner) partials
    -- fixed order, deterministic.  Operand classes as in gen_rows ('v' = contiguous along inner)."""
    arrays, scalars = prog.arrays, prog.scalars
    rop, acc_dt, res_dt, post = reduce
    A = ctype(acc_dt)
    OUT = A if partial else ctype(res_dt)
    n_ops = max(len(arrays), 1)
    src = []
    w = src.append
    w(f"struct Geo_{name} {{ i64 outer; i64 red; i64 inner; i64 so[{n_ops}]; i64 sr[{n_ops}]; i64 si[{n_ops}]; i64 chunk; }};")
    params = [f"const Geo_{name} g"]
    for i, a in enumerate(arrays):
        params.append(f"const char* __restrict__ in{i}")
    for j, (_, dt) in enumerate(scalars):
        params.append(f"const {ctype(dt)} s{j}")
    params += [f"{OUT}* __restrict__ result", "const double post_scale"]
    body = emit_body(prog)
    ident = _identity(rop, acc_dt)
    w(f'extern "C" __global__ void __launch_bounds__({threads}) {name}({", ".join(params)}) {{')
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        if c == "b":
            w(f"  const {ctype(a.dtype)} x{i} = *reinterpret_cast<const {ctype(a.dtype)}*>(in{i});")
    w(f"  const i64 nvec = g.inner / {V};")
    w("  const i64 total = g.outer * nvec;")
    w("  const i64 r0 = (i64)blockIdx.y * g.chunk, r1 = r0 + g.chunk < g.red ? r0 + g.chunk : g.red;")
    w("  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {")
    w(f"    const i64 o = idx / nvec, c = (idx - o * nvec) * {V};")
    w(f"    {A} acc[{V}];")
    w(f"#pragma unroll\n    for (int e = 0; e < {V}; ++e) acc[e] = {ident};")
    # 8-byte operands: sixteen trips in flight through the non-coherent path (measured on 16384^2
    # float64: sum(X, axis=0) 5.7 -> 6.1 TB/s, v @ X 4.2 -> 5.4; float32 measured indifferent to both,
    # profiles/r2_cols_family_sweep.txt)
    wide = any(a.dtype.itemsize == 8 and c == "v" for a, c in zip(arrays, in_class))
    unroll, stream = (16, "true") if wide else (4, "false")
    w(f"#pragma unroll {unroll}\n    for (i64 r = r0; r < r1; ++r) {{")
    for i, (a, c) in enumerate(zip(arrays, in_class)):
        T = ctype(a.dtype)
        if c == "v":
            w(f"      const Vec<{T}, {V}> vx{i} = dr_ld<{stream}, {T}, {V}>(reinterpret_cast<const {T}*>"
              f"(in{i} + o * g.so[{i}] + r * g.sr[{i}]) + c);")
        elif c == "i":
            w(f"      const {T} ix{i} = *reinterpret_cast<const {T}*>(in{i} + o * g.so[{i}] + r * g.sr[{i}]);")
        elif c == "s":
            w(f"      {T} sx{i}[{V}];")
            w(f"#pragma unroll\n      for (int e = 0; e < {V}; ++e) sx{i}[e] = *reinterpret_cast<const {T}*>"
              f"(in{i} + o * g.so[{i}] + r * g.sr[{i}] + (c + e) * g.si[{i}]);")
    w(f"#pragma unroll\n      for (int e = 0; e < {V}; ++e) {{")
    _emit_lane_operands(w, arrays, in_class, V, "        ")
    for line in body:
        w(f"        {line}")
    w(f"        acc[e] = {_RED[rop]}::op(acc[e], ({A}){_operand_name(prog.roots[0])});")
    w("      }")
    w("    }")
    if partial:
        w(f"#pragma unroll\n    for (int e = 0; e < {V}; ++e)")
        w("      result[(o * gridDim.y + blockIdx.y) * g.inner + c + e] = acc[e];")
    else:
        fin = f"({ctype(res_dt)})(acc[e] / ({A})post_scale)" if np.dtype(acc_dt).kind == "f" \
            else f"({ctype(res_dt)})acc[e]"
        w(f"#pragma unroll\n    for (int e = 0; e < {V}; ++e) result[o * g.inner + c + e] = {fin};")
    w("  }")
    w("}")
    return "\n".join(src) + "\n"


# --------------------------------------------------------------------------- stencil family
# Device helpers of the halo-pushing stencil variant.  They live here and not in prelude.cuh so
# that the text (and with it the cubin cache key) of every other kernel stays what it was.
HALO_HELPERS = r"""
// Row-sharded stencils (sharding.py): a kernel stores its boundary rows straight into the
// neighbour GPU's block (NVLink peer mapping) and then publishes the step number with a
// system-scope release store; the neighbour's next step acquires it before its TMA unit reads
// the halo rows.  No NCCL, no host round trip, nothing but the fused kernel on the stream.
__device__ __forceinline__ unsigned dr_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void dr_st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long dr_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Spin until *flag has reached step `want` (wrap-safe).  A neighbour that never signals (its
// process died) must not hang the GPU: trap after 120 s (a rank that is
// merely late -- a cold NVRTC compile, a paged-out interpreter -- is waited for).
__device__ __forceinline__ void dr_wait_epoch(const unsigned* flag, unsigned want) {
  unsigned long long t0 = 0;
  while ((int)(dr_ld_acquire_sys(flag) - want) < 0) {
    __nanosleep(100);
    const unsigned long long now = dr_globaltimer();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 120000000000ull) __trap();
  }
  // the halo rows were written by another GPU (generic proxy); the TMA unit reads them through
  // the async proxy
  asm volatile("fence.proxy.async;" ::: "memory");
}
"""


def gen_stencil(name, prog, roles, out_dt, TW=248, TH=32, NS=4, threads=992, halo=False):
    """Shifted-view stencil over ONE 2-d base array, written to a fresh copy of the base
    (ping-pong: Jacobi semantics without the reference's temporary + copy, delayarray.py:114-121).

    roles[i] = ("tile", dy, dx)  operand i is the base array shifted by (dy, dx) relative to the
                                 cell being written -> read from the TMA-staged shared tile
             = ("b",)            broadcast scalar operand (one global load)
             = ("g",)            another array over the target region -> coalesced global load
    Persistent CTAs walk the tiles of the base; an NS-deep ring of (TH + halo) x (TW + halo)
    boxes is filled by cp.async.bulk.tensor.2d + mbarrier so the load of tile t+NS-1 overlaps the
    arithmetic of tile t.  Each thread produces 4 consecutive cells (one 128-bit store) per row
    pass through the lockstep / packed-f32x2 body.  Cells of the base outside the assigned view
    are copied through unchanged.

    ``halo=True`` (row-sharded arrays, sharding.py): the block carries H halo rows on each side
    that belong to the neighbour ranks.  The kernel then (i) walks the tile rows that contain
    its first / last H owned rows FIRST, stores those rows a second time into the neighbour
    GPU's block (peer-mapped memory over NVLink) and, when the last such tile has finished,
    publishes the step number to the neighbour with a system-scope release store; (ii) before
    the TMA load of any tile whose box touches a halo row, acquires the step number the
    neighbour published in the previous step; (iii) never writes its own halo rows (the
    neighbours do).  The exchange is therefore part of the stencil kernel: boundary rows travel
    while the interior is being computed, and nothing else is on the stream."""
    TW = int(os.environ.get("DR_ST_TW", TW))
    TH = int(os.environ.get("DR_ST_TH", TH))
    NS = int(os.environ.get("DR_ST_NS", NS))
    threads = int(os.environ.get("DR_ST_THREADS", threads))
    arrays, scalars = prog.arrays, prog.scalars
    T = ctype(out_dt)
    V = 16 // np.dtype(out_dt).itemsize
    tiles = [r for r in roles if r[0] == "tile"]
    hu = max(0, -min(r[1] for r in tiles))
    hd = max(0, max(r[1] for r in tiles))
    hl = max(0, -min(r[2] for r in tiles))
    hr = max(0, max(r[2] for r in tiles))
    hl_pad = -(-hl // V) * V                       # keep the tile's own columns 16-byte aligned
    BW = hl_pad + TW + (-(-hr // V) * V)
    BH = hu + TH + hd
    assert BW <= 256 and BH <= 256
    stage_bytes = BW * BH * np.dtype(out_dt).itemsize
    stage_bytes_al = -(-stage_bytes // 128) * 128
    # the ring must fit the 227 KB of one SM: float64 boxes are twice as large (3 stages)
    NS = max(2, min(NS, (216 * 1024) // stage_bytes_al))
    cols_per_row = TW // V                          # threads along x
    rows_per_pass = threads // cols_per_row
    in_class = tuple("b" if r[0] == "b" else "c" for r in roles)
    lock_body, lock_uniform = emit_body_lockstep(prog, in_class, V) if V == 4 else (None, None)
    scalar_body = emit_body(prog, fast=False)
    src = []
    w = src.append
    w(f"struct Geo_{name} {{ int rows, cols, r0, c0, h, w, tiles_x, ntiles; i64 pitch_elems; "
      f"i64 gs_row[{max(len(arrays), 1)}]; i64 gs_col[{max(len(arrays), 1)}]; }};")
    params = [f"const __grid_constant__ DrTensorMap tmap", f"const Geo_{name} g",
              f"const {T}* __restrict__ base", f"{T}* __restrict__ out"]
    if halo:
        w(HALO_HELPERS)
        # up = the neighbour that owns the rows above this block, dn = below.  peer_*_rows: where
        # my first / last H owned rows go in THEIR output block; peer_*_flag: their "from below" /
        # "from above" flag; my_flag_*: mine; cnt: two tile counters in my memory.
        w(f"struct Halo_{name} {{ unsigned long long peer_up_rows, peer_dn_rows, peer_up_flag, peer_dn_flag, "
          f"my_flag_up, my_flag_dn, cnt; unsigned epoch; int H, n_up_tiles, n_dn_tiles, nprio; int prio[4]; }};")
        params.append(f"const Halo_{name} hx")
    for i, a in enumerate(arrays):
        params.append(f"const char* __restrict__ in{i}")
    for j, (_, dt) in enumerate(scalars):
        params.append(f"const {ctype(dt)} s{j}")
    w(f'extern "C" __global__ void __launch_bounds__({threads}) {name}({", ".join(params)}) {{')
    w("  extern __shared__ __align__(128) unsigned char dr_smem[];")
    w(f"  __shared__ __align__(8) unsigned long long bar[{NS}];")
    w("  const int tid = threadIdx.x;")
    w(f"  if (tid == 0) {{ for (int s = 0; s < {NS}; ++s) dr_mbar_init(&bar[s], 1); dr_fence_barrier_init(); }}")
    w("  __syncthreads();")
    for i, (a, r) in enumerate(zip(arrays, roles)):
        if r[0] == "b":
            w(f"  const {ctype(a.dtype)} x{i} = *reinterpret_cast<const {ctype(a.dtype)}*>(in{i});")
    if lock_body is not None:
        emit_explog_stage(w, prog)
    if lock_body is not None and uses_erf_table(prog):
        if GEN2:
            w("  __shared__ float2 dr_erf_tab[3 * DR_ERF2_ROWS];")
            w("  dr_erf2_tab_stage<1>(dr_erf_tab);")
        else:
            w("  __shared__ float2 dr_erf_tab[DR_ERF_TAB_PAIRS];")
            w("  dr_erf_tab_stage(dr_erf_tab);")
    w(f"  const int tx = tid % {cols_per_row}, ty = tid / {cols_per_row};")
    if halo:
        w("  const bool has_up = hx.peer_up_rows != 0ull, has_dn = hx.peer_dn_rows != 0ull;")
        w("  bool up_ok = !has_up, dn_ok = !has_dn;")
        w("  const int own_lo = hx.H, own_hi = g.rows - hx.H;          // owned rows [own_lo, own_hi)")
        # boundary tile rows first: hx.prio (sorted) lists them, the rest follow in order
        # (selects on the four kernel-parameter words: indexing the array would put it on the stack)
        w("  const int p0 = hx.prio[0], p1 = hx.prio[1], p2 = hx.prio[2], p3 = hx.prio[3], np_ = hx.nprio;")
        w("  auto tile_row = [&](int trow) {")
        w("    if (trow < np_) return trow == 0 ? p0 : trow == 1 ? p1 : trow == 2 ? p2 : p3;")
        w("    int by = trow - np_;")
        w("    by += (np_ > 0 && by >= p0); by += (np_ > 1 && by >= p1);")
        w("    by += (np_ > 2 && by >= p2); by += (np_ > 3 && by >= p3);")
        w("    return by;")
        w("  };")
    w("  auto issue = [&](int tile, int stage) {")
    w("    if (tile < g.ntiles) {")
    if halo:
        w("      const int trow = tile / g.tiles_x, bx = tile - trow * g.tiles_x, by = tile_row(trow);")
        w(f"      if (!up_ok && by * {TH} - {hu} < own_lo) {{ dr_wait_epoch(reinterpret_cast<const unsigned*>(hx.my_flag_up), hx.epoch - 1u); up_ok = true; }}")
        w(f"      if (!dn_ok && by * {TH} + {TH + hd} > own_hi) {{ dr_wait_epoch(reinterpret_cast<const unsigned*>(hx.my_flag_dn), hx.epoch - 1u); dn_ok = true; }}")
    else:
        w(f"      const int by = tile / g.tiles_x, bx = tile - by * g.tiles_x;")
    w(f"      dr_mbar_expect_tx(&bar[stage], {stage_bytes});")
    w(f"      dr_tma_load_2d(dr_smem + stage * {stage_bytes_al}, &tmap, bx * {TW} - {hl_pad}, by * {TH} - {hu}, &bar[stage]);")
    w("    }")
    w("  };")
    w(f"  if (tid == 0) {{ for (int k = 0; k < {NS - 1}; ++k) issue(blockIdx.x + k * gridDim.x, k); }}")
    # Measured (tools/heat_shard_probe.py, 32768^2 float32): the interior variant executes 16
    # instead of 24 instructions per cell and is SLOWER -- 1.43 ms per step with the row guard kept,
    # 1.59 ms without it, against 1.31 ms for the general code (= 99 % of the measured copy
    # bandwidth).  The kernel is not issue-bound; what it needs is the general variant's load /
    # store interleaving.  The variant stays available for experiments (DR_ST_FAST=1).
    fast_ok = os.environ.get("DR_ST_FAST") == "1"
    fast_guard = os.environ.get("DR_ST_FASTGUARD", "1") == "1"

    def emit_passes(inner, edge=False):
        """The row passes of one tile.  inner=True: the tile lies wholly inside the assigned view
        and inside the array, so every bounds test, the keep/select per cell and the clamping of
        gathered operands are dropped (most tiles; 24 -> 16 instructions per cell for the
        5-point stencil)."""
        w("#pragma unroll")
        w(f"    for (int pass = 0; pass < {TH // rows_per_pass}; ++pass) {{")
        w(f"      const int ly = pass * {rows_per_pass} + ty, gy = by * {TH} + ly;")
        w("      if (gy < g.rows && gx < g.cols) {" if (not inner or fast_guard) else "      {")
        w("        constexpr int u = 0;")
        # gather operands: ONE aligned 128-bit shared load per distinct row offset dy; horizontally
        # shifted views reuse that vector and fetch only the |dx| cells that fall outside it
        # (scalar loads) -- 3 LDS.128 + 2 LDS.32 per 4 cells for the 5-point stencil instead of 11
        row_vec = {}
        for r in roles:
            if r[0] == "tile" and r[1] not in row_vec:
                dy = r[1]
                nm = f"row{'m' if dy < 0 else 'p'}{abs(dy)}"
                row_vec[dy] = nm
                w(f"        const Vec<{T}, {V}> {nm} = *reinterpret_cast<const Vec<{T}, {V}>*>"
                  f"(sm + (ly + {hu + dy}) * {BW} + {hl_pad} + tx * {V});")
        for i, (a, r) in enumerate(zip(arrays, roles)):
            A = ctype(a.dtype)
            if r[0] == "tile":
                dy, dx = r[1], r[2]
                w(f"        Vec<{A}, {V}> v{i}[1];")
                base = f"(ly + {hu + dy}) * {BW} + {hl_pad} + tx * {V}"
                for e in range(V):
                    src_e = e + dx
                    if 0 <= src_e < V:
                        w(f"        v{i}[0].v[{e}] = {row_vec[dy]}.v[{src_e}];")
                    else:
                        w(f"        v{i}[0].v[{e}] = sm[{base} + {src_e}];")
            elif r[0] == "g":
                w(f"        Vec<{A}, {V}> v{i}[1];")
                w(f"#pragma unroll\n        for (int e = 0; e < {V}; ++e) {{")
                if inner:
                    w("          const int yy = gy - g.r0, xx = gx + e - g.c0;")
                else:
                    w(f"          const int yy = min(max(gy - g.r0, 0), g.h - 1), xx = min(max(gx + e - g.c0, 0), g.w - 1);")
                w(f"          v{i}[0].v[e] = *reinterpret_cast<const {A}*>(in{i} + yy * g.gs_row[{i}] + xx * g.gs_col[{i}]);")
                w("        }")
        if 0 not in row_vec:
            w(f"        const Vec<{T}, {V}> rowp0 = *reinterpret_cast<const Vec<{T}, {V}>*>"
              f"(sm + (ly + {hu}) * {BW} + {hl_pad} + tx * {V});")
        if not inner:
            w(f"        const Vec<{T}, {V}> keep = rowp0;")
        w(f"        Vec<{T}, {V}> r0;")
        if lock_body is not None:
            w("      " + DR_ONE.format("(long long)g.h"))
            w("        bool bad = false;")
            for line in lock_body:
                w(f"        {line}")
            w(f"#pragma unroll\n        for (int e = 0; e < {V}; ++e) r0.v[e] = {_lane_store(prog, prog.roots[0], out_dt, in_class, lock_uniform)};")
            if has_lane_fast(prog):
                w("        if (bad) {")
                w(f"#pragma unroll\n          for (int e = 0; e < {V}; ++e) {{")
                for i, (a, r) in enumerate(zip(arrays, roles)):
                    if r[0] != "b":
                        w(f"            const {ctype(a.dtype)} x{i} = v{i}[0].v[e];")
                for line in scalar_body:
                    w(f"            {line}")
                w(f"            r0.v[e] = {_store_expr(prog, prog.roots[0], out_dt)};")
                w("          }")
                w("        }")
        else:
            w(f"#pragma unroll\n        for (int e = 0; e < {V}; ++e) {{")
            for i, (a, r) in enumerate(zip(arrays, roles)):
                if r[0] != "b":
                    w(f"          const {ctype(a.dtype)} x{i} = v{i}[0].v[e];")
            for line in scalar_body:
                w(f"          {line}")
            w(f"          r0.v[e] = {_store_expr(prog, prog.roots[0], out_dt)};")
            w("        }")
        if not inner:
            w("        const bool row_in = gy >= g.r0 && gy < g.r0 + g.h;")
            w(f"#pragma unroll\n            for (int e = 0; e < {V}; ++e) {{")
            w("          const int x = gx + e;")
            w("          if (!(row_in && x >= g.c0 && x < g.c0 + g.w)) r0.v[e] = keep.v[e];")
            w("        }")
        if halo and edge:
            # halo rows belong to the neighbours (they store them); my first / last H owned rows are
            # stored twice: into my block and into the neighbour's
            w("        if (!((has_up && gy < own_lo) || (has_dn && gy >= own_hi)))")
            w(f"          dr_st<true, {T}, {V}>(out + (i64)gy * g.pitch_elems + gx, r0);")
            w("        if (push_up && gy >= own_lo && gy < own_lo + hx.H)")
            w(f"          dr_st<false, {T}, {V}>(reinterpret_cast<{T}*>(hx.peer_up_rows) + (i64)(gy - own_lo) * g.pitch_elems + gx, r0);")
            w("        if (push_dn && gy >= own_hi - hx.H && gy < own_hi)")
            w(f"          dr_st<false, {T}, {V}>(reinterpret_cast<{T}*>(hx.peer_dn_rows) + (i64)(gy - (own_hi - hx.H)) * g.pitch_elems + gx, r0);")
        else:
            w(f"        dr_st<true, {T}, {V}>(out + (i64)gy * g.pitch_elems + gx, r0);")
        w("      }")
        w("    }")

    def emit_tile(edge):
        """One trip of the tile loop.  edge=True (halo kernels only): the tile row holds halo rows
        or rows that are pushed to a neighbour."""
        w(f"    const int stage = it % {NS};")
        w(f"    if (tid == 0) issue(tile + {NS - 1} * gridDim.x, (it + {NS - 1}) % {NS});")
        w(f"    dr_mbar_wait(&bar[stage], (it / {NS}) & 1);")
        w(f"    const {T}* sm = reinterpret_cast<const {T}*>(dr_smem + stage * {stage_bytes_al});")
        if halo:
            w("    const int trow = tile / g.tiles_x, bx = tile - trow * g.tiles_x, by = tile_row(trow);")
        else:
            w("    const int by = tile / g.tiles_x, bx = tile - by * g.tiles_x;")
        if edge:
            w(f"    const bool push_up = has_up && by * {TH} < own_lo + hx.H && by * {TH} + {TH} > own_lo;")
            w(f"    const bool push_dn = has_dn && by * {TH} < own_hi && by * {TH} + {TH} > own_hi - hx.H;")
        w(f"    const int gx = bx * {TW} + tx * {V};")
        if edge:
            emit_passes(False, edge=True)
        elif fast_ok and not halo:
            w(f"    if (by * {TH} >= g.r0 && by * {TH} + {TH} <= g.r0 + g.h && bx * {TW} >= g.c0 && bx * {TW} + {TW} <= g.c0 + g.w) {{")
            emit_passes(True)
            w("    } else {")
            emit_passes(False)
            w("    }")
        else:
            emit_passes(False)
        # (no per-thread system fence here: the block barrier orders every thread's peer stores
        # before thread 0, whose own cumulative fence below then orders them before the flag --
        # 992 MEMBAR.SC.SYS per edge tile cost ~7 us per kernel)
        w("    __syncthreads();")
        if edge:
            # the last boundary tile to finish publishes this step to the neighbour; by then every
            # tile that read the halo rows on that side has consumed them (same tiles), so the
            # neighbour may overwrite them in ITS next step
            w("    if (tid == 0 && (push_up || push_dn)) {")
            w("      unsigned* cnt = reinterpret_cast<unsigned*>(hx.cnt);")
            w("      __threadfence_system();")
            w("      if (push_up && atomicAdd(&cnt[0], 1u) == (unsigned)hx.n_up_tiles - 1u) {")
            w("        cnt[0] = 0u; __threadfence_system();")
            w("        dr_st_release_sys(reinterpret_cast<unsigned*>(hx.peer_up_flag), hx.epoch);")
            w("      }")
            w("      if (push_dn && atomicAdd(&cnt[16], 1u) == (unsigned)hx.n_dn_tiles - 1u) {")
            w("        cnt[16] = 0u; __threadfence_system();")
            w("        dr_st_release_sys(reinterpret_cast<unsigned*>(hx.peer_dn_flag), hx.epoch);")
            w("      }")
            w("    }")

    w("  int it = 0, tile = blockIdx.x;")
    if halo:
        # the boundary tile rows come first in the tile order (tile_row) and have a loop of their
        # own; the main loop below is the unsharded kernel's, instruction for instruction
        w("  for (; tile < hx.nprio * g.tiles_x; tile += gridDim.x, ++it) {")
        emit_tile(True)
        w("  }")
    w("  for (; tile < g.ntiles; tile += gridDim.x, ++it) {")
    emit_tile(False)
    w("  }")
    w("}")
    meta = dict(TW=TW, TH=TH, NS=NS, BW=BW, BH=BH, hl_pad=hl_pad, hu=hu, hd=hd, smem=NS * stage_bytes_al,
                threads=threads)
    return "\n".join(src) + "\n", meta


# --------------------------------------------------------------------------- skinny contraction
def gen_mm_skinny(name, prog, roles, t_dt, n_out, threads=128, tile_k=256, rows=4, pow_modes=None):
    """out[i, 0:n_out] (+)= sum_k A(i, k) * B[k, 0:n_out] where A is a fused elementwise program
    whose array operands each vary along rows only ('r'), along k only ('c') or not at all ('b')
    -- the all-pairs pattern x[None, :] - x[:, None] ... of the n-body workload.  Register tile:
    every thread owns `rows` output rows (rows threads apart, so the final stores coalesce) and
    evaluates the producer for all of them against one k at a time, so the k-varying operands and
    the B row are read from shared memory ONCE per `rows` pairs (broadcast reads: every lane the
    same address).  grid.y splits K, each split writes its own partial block (deterministic: the
    partials are summed by a second fused kernel, no atomics).  B's last column may be a virtual
    column of ones (row sum of A folded into the same pass)."""
    arrays, scalars = prog.arrays, prog.scalars
    T = ctype(t_dt)
    R = rows
    n_ops = max(len(arrays), 1)
    src = []
    w = src.append
    w(f"struct Geo_{name} {{ i64 M, K, kchunk; i64 sr[{n_ops}]; i64 sc[{n_ops}]; i64 b_rs, b_cs; int n_real; }};")
    params = [f"const Geo_{name} g"]
    for i, a in enumerate(arrays):
        params.append(f"const char* __restrict__ in{i}")
    for j, (_, dt) in enumerate(scalars):
        params.append(f"const {ctype(dt)} s{j}")
    params += ["const char* __restrict__ B", f"{T}* __restrict__ partial"]
    body = emit_body(prog, fast=False, relaxed=True, pow_modes=pow_modes)
    w(f'extern "C" __global__ void __launch_bounds__({threads}) {name}({", ".join(params)}) {{')
    cols = [i for i, r in enumerate(roles) if r == "c"]
    rws = [i for i, r in enumerate(roles) if r == "r"]
    for i in cols:
        w(f"  __shared__ {ctype(arrays[i].dtype)} sh{i}[{tile_k}];")
    w(f"  __shared__ __align__(16) {T} shB[{tile_k}][{n_out}];")
    w("  const int tid = threadIdx.x;")
    w(f"  const i64 row0 = (i64)blockIdx.x * {threads * R} + tid;")
    for i, (a, r) in enumerate(zip(arrays, roles)):
        A = ctype(a.dtype)
        if r == "b":
            w(f"  const {A} x{i} = *reinterpret_cast<const {A}*>(in{i});")
    for i in rws:
        A = ctype(arrays[i].dtype)
        w(f"  {A} xr{i}[{R}];")
    w(f"#pragma unroll\n  for (int r = 0; r < {R}; ++r) {{")
    w(f"    const i64 row = row0 + r * {threads};")
    w("    const i64 rr = row < g.M ? row : g.M - 1;")
    for i in rws:
        A = ctype(arrays[i].dtype)
        w(f"    xr{i}[r] = *reinterpret_cast<const {A}*>(in{i} + rr * g.sr[{i}]);")
    w("  }")
    w(f"  {T} acc[{R}][{n_out}];")
    w(f"#pragma unroll\n  for (int r = 0; r < {R}; ++r)")
    w(f"#pragma unroll\n    for (int n = 0; n < {n_out}; ++n) acc[r][n] = ({T})0;")
    w("  const i64 k0 = (i64)blockIdx.y * g.kchunk;")
    w("  const i64 k1 = k0 + g.kchunk < g.K ? k0 + g.kchunk : g.K;")
    w(f"  for (i64 kt = k0; kt < k1; kt += {tile_k}) {{")
    w(f"    for (int t = tid; t < {tile_k}; t += {threads}) {{")
    w("      const i64 kk = kt + t < k1 ? kt + t : k1 - 1;")
    for i in cols:
        A = ctype(arrays[i].dtype)
        w(f"      sh{i}[t] = *reinterpret_cast<const {A}*>(in{i} + kk * g.sc[{i}]);")
    w(f"#pragma unroll\n      for (int n = 0; n < {n_out}; ++n)")
    w(f"        shB[t][n] = n < g.n_real ? *reinterpret_cast<const {T}*>(B + kk * g.b_rs + n * g.b_cs) : ({T})1;")
    w("    }")
    w("    __syncthreads();")
    w(f"    const int lim = (int)(k1 - kt < {tile_k} ? k1 - kt : {tile_k});")
    w("#pragma unroll 2")
    w("    for (int j = 0; j < lim; ++j) {")
    for i in cols:
        w(f"      const {ctype(arrays[i].dtype)} x{i} = sh{i}[j];")
    w(f"      {T} bj[{n_out}];")
    w(f"#pragma unroll\n      for (int n = 0; n < {n_out}; ++n) bj[n] = shB[j][n];")
    w(f"#pragma unroll\n      for (int r = 0; r < {R}; ++r) {{")
    for i in rws:
        w(f"        const {ctype(arrays[i].dtype)} x{i} = xr{i}[r];")
    for line in body:
        w(f"        {line}")
    root = _store_expr(prog, prog.roots[0], t_dt)
    w(f"        const {T} a_ik = {root};")
    if os.environ.get("DR_SK_NOCONTRACT"):
        # EXPERIMENT (wrong results): the producer alone -- one add per pair keeps it alive, the
        # n_out contraction FMAs are gone.  full - this = the most a tensor-core contraction could
        # take off the FP32 pipe (mm_skinny).
        w("        acc[r][0] += a_ik;")
    else:
        w(f"#pragma unroll\n        for (int n = 0; n < {n_out}; ++n) acc[r][n] = fma(a_ik, bj[n], acc[r][n]);")
    w("      }")
    w("    }")
    w("    __syncthreads();")
    w("  }")
    w(f"#pragma unroll\n  for (int r = 0; r < {R}; ++r) {{")
    w(f"    const i64 row = row0 + r * {threads};")
    w("    if (row < g.M) {")
    w(f"#pragma unroll\n      for (int n = 0; n < {n_out}; ++n) partial[((i64)blockIdx.y * g.M + row) * {n_out} + n] = acc[r][n];")
    w("    }")
    w("  }")
    w("}")
    return "\n".join(src) + "\n"
