"""Interval analysis of a fused float32 program: which of the branch-free fast forms
(prelude.cuh: dr_div4_r, dr_sqrt4_r, dr_log4_r, dr_exp4_r, dr_erf4_gal) need a run-time range
test, and which array operands get the one combined test at the top of the vector instead.

The fast forms are exact only on a domain (normal, finite, no underflow of the quotient or of
the exact residual, ...).  The first generation tested every operand of every such operation on
every lane (~16 integer instructions per Black-Scholes option).  Here each value carries a
conservative description

    2^lo <= |x| <= 2^hi when x != 0;  zero / negative zero / negative / positive possible?

(always finite and not nan; ``None`` = nothing known).  Array operands that feed a fast form are
*assumed* to lie in +-[2^-30, 2^30) -- or (0, ...) when something needs a positive argument --
and the kernel proves the assumption with ONE min/max tree over all lanes of all such operands
(DrRange in prelude.cuh); float32 scalars are classified on the host at launch time ('n' =
2^-24 <= |s| < 2^24) and the class is part of the kernel's structural key.  Wherever the
propagated interval does not imply an operation's precondition the per-lane test stays, exactly
as before; a vector that fails any test is recomputed through the precise scalar forms.

No reference counterpart: the reference emits `T a = b / c` and lets CuPy's compiler decide
(cuda.py:55-76); this module only decides where guards go, never what is computed.
"""
import numpy as np

F32 = np.dtype(np.float32)
IN_LO, IN_HI = -30, 30          # DR_IN_LO / DR_IN_HI in prelude.cuh
SC_LO, SC_HI = -24, 24
MIN_SUB, MIN_NORM, MAX_EXP = -149, -126, 127


class R:
    """2^lo <= |x| <= 2^hi for non-zero x; flags say what else is possible."""
    __slots__ = ("lo", "hi", "zero", "nz", "neg", "pos")

    def __init__(self, lo, hi, zero=False, nz=False, neg=True, pos=True):
        self.lo, self.hi, self.zero, self.nz, self.neg, self.pos = lo, hi, zero, nz, neg, pos

    def __repr__(self):
        s = ("-" if self.neg else "") + ("+" if self.pos else "")
        return f"R[{self.lo},{self.hi}){s}{'0' if self.zero else ''}{'z' if self.nz else ''}"


def scalar_class(value, dtype):
    """'n' (nice): float32 scalar with 2^-24 <= |s| < 2^24; anything else 'u' (unknown)."""
    if np.dtype(dtype) != F32:
        return "u"
    a = abs(float(value))
    return "n" if (2.0 ** SC_LO <= a < 2.0 ** SC_HI) else "u"


def scalar_classes(prog):
    return tuple(scalar_class(v, dt) for v, dt in prog.scalars)


def _may_neg_sign(r):
    return r.neg or r.nz


def _may_pos_sign(r):
    return r.pos or r.zero


def _mul(a, b):
    if a is None or b is None:
        return None
    hi = a.hi + b.hi
    if hi > MAX_EXP:
        return None
    lo = a.lo + b.lo - 1                       # rounding of a subnormal product may shrink it
    zero = a.zero or b.zero or lo < MIN_SUB
    lo = max(lo, MIN_SUB)
    neg = (a.neg and b.pos) or (a.pos and b.neg)
    pos = (a.pos and b.pos) or (a.neg and b.neg)
    nz = zero and ((_may_neg_sign(a) and _may_pos_sign(b)) or (_may_pos_sign(a) and _may_neg_sign(b)))
    return R(lo, hi, zero, nz, neg, pos)


def _addsub(a, b, sub):
    if a is None or b is None:
        return None
    hi = max(a.hi, b.hi) + 1
    if hi > MAX_EXP:
        return None
    bneg, bpos = (b.pos, b.neg) if sub else (b.neg, b.pos)
    cancel = (a.neg and bpos) or (a.pos and bneg)
    if cancel:
        # every float of magnitude >= 2^lo is a multiple of 2^(lo-23): so is the sum
        lo = max(min(a.lo, b.lo) - 24, MIN_SUB)
        zero = True
    else:
        lo = min(a.lo, b.lo)
        zero = a.zero and b.zero
    # -0 only from (-0) + (-0) resp. (-0) - (+0); exact cancellation gives +0 (round to nearest)
    nz = a.nz and (b.zero if sub else b.nz)
    return R(lo, hi, zero, nz, a.neg or bneg, a.pos or bpos)


class Analysis:
    """Result of analyse(): per-instruction guard decisions and the operand pre-test lists."""

    def __init__(self):
        self.check = {}          # instr index -> tuple of bools (one per guarded operand)
        self.pos_inputs = []     # array operand indices tested as positive in [2^-30, 2^30)
        self.any_inputs = []     # array operand indices tested as +-[2^-30, 2^30)
        self.ranges = {}         # ref -> R | None   (for tests / debugging)


def _lane_f32(prog, k):
    op, loop, out_dt, args = prog.instrs[k]
    return out_dt == F32 and all(d == F32 for d in loop) and all(prog.dtypes[r] == F32 for r in args)


def analyse(prog, in_class, sclasses, guarded_ops):
    """``guarded_ops``: the op names that have a lane fast form (codegen._LANE4_FAST).
    ``in_class``: per array operand 'c' (contiguous vector operand) or 'b' (broadcast scalar)."""
    # ---- pass 1: which array operands feed a guarded operation, and does anything need them > 0
    demand = {}                                       # array index -> "pos" | "any"

    def want(ref, kind):
        if ref[0] == "a" and in_class[ref[1]] == "c" and prog.dtypes[ref] == F32:
            if demand.get(ref[1]) != "pos":
                demand[ref[1]] = kind

    def want_pos_through(ref):
        """log/sqrt of a quotient or product of array operands: ask for positive operands."""
        want(ref, "pos")
        if ref[0] == "t":
            op, loop, out_dt, args = prog.instrs[ref[1]]
            if op in ("true_divide", "divide", "multiply") and _lane_f32(prog, ref[1]):
                for r in args:
                    want(r, "pos")

    for k, (op, loop, out_dt, args) in enumerate(prog.instrs):
        if op not in guarded_ops or not _lane_f32(prog, k):
            continue
        if op in ("true_divide", "divide"):
            want(args[0], "any")
            want(args[1], "any")
        elif op in ("sqrt", "log"):
            want_pos_through(args[0])
        # exp / erf accept any finite argument: |x| < 87 is tested on the argument itself

    out = Analysis()
    out.pos_inputs = sorted(i for i, kd in demand.items() if kd == "pos")
    out.any_inputs = sorted(i for i, kd in demand.items() if kd == "any")

    # ---- pass 2: forward propagation
    val = {}
    for i in range(len(prog.arrays)):
        if demand.get(i) == "pos":
            val[("a", i)] = R(IN_LO, IN_HI, neg=False)
        elif demand.get(i) == "any":
            val[("a", i)] = R(IN_LO, IN_HI)
        else:
            val[("a", i)] = None
    for j, (v, dt) in enumerate(prog.scalars):
        if sclasses is not None and sclasses[j] == "n":
            val[("s", j)] = R(SC_LO, SC_HI)
        else:
            val[("s", j)] = None

    for k, (op, loop, out_dt, args) in enumerate(prog.instrs):
        me = ("t", k)
        a = [val.get(r) for r in args]
        if not _lane_f32(prog, k):
            val[me] = None
            continue
        if op == "multiply":
            val[me] = _mul(a[0], a[1])
        elif op == "add":
            val[me] = _addsub(a[0], a[1], False)
        elif op == "subtract":
            val[me] = _addsub(a[0], a[1], True)
        elif op == "negative":
            r = a[0]
            val[me] = None if r is None else R(r.lo, r.hi, r.zero, r.zero, r.pos, r.neg)
        elif op in ("absolute", "fabs"):
            r = a[0]
            val[me] = None if r is None else R(r.lo, r.hi, r.zero, False, False, True)
        elif op in ("true_divide", "divide"):
            n, d = a
            checked = lambda r: R(-60, 61, neg=True if r is None else r.neg,          # noqa: E731
                                  pos=True if r is None else r.pos)                  # after dr_tame
            # divisor: normal with a normal reciprocal
            cb = d is None or d.zero or d.lo < -100 or d.hi > 100
            if cb:
                d = checked(d)

            def joint(n, d):     # quotient normal, residual a - b q exactly representable
                return n.hi - d.lo + 1 <= 126 and n.lo - d.hi - 1 >= -125 and n.lo >= -100
            ca = n is None or n.nz or not joint(n, d)       # (+0 is fine, -0 is not)
            if ca:
                n = checked(n)
                if not joint(n, d):
                    cb, d = True, checked(d)
            if op in guarded_ops:
                out.check[k] = (ca, cb)
            hi, lo = n.hi - d.lo + 1, n.lo - d.hi - 1
            neg = (n.neg and d.pos) or (n.pos and d.neg)
            pos = (n.pos and d.pos) or (n.neg and d.neg)
            val[me] = R(lo, hi, n.zero, n.zero and (d.neg or n.nz), neg, pos)
            if op not in guarded_ops:
                val[me] = None
        elif op == "sqrt":
            r = a[0]
            c = r is None or r.neg or r.zero or r.nz or r.lo < MIN_NORM
            if c:
                r = R(-60, 61, neg=False)
            if op in guarded_ops:
                out.check[k] = (c,)
                val[me] = R(r.lo // 2 - 1, -(-r.hi // 2) + 1, neg=False)
            else:
                val[me] = None
        elif op == "log":
            r = a[0]
            c = r is None or r.neg or r.zero or r.nz or r.lo < MIN_NORM
            if op in guarded_ops:
                out.check[k] = (c,)
                # |log x| < 89; the non-zero value closest to 0 is log(1 - 2^-24) ~ -2^-24
                val[me] = R(-25, 7, zero=True)
            else:
                val[me] = None
        elif op == "exp":
            r = a[0]
            # 0: |x| <= 64 proven; 1: nothing known (per-lane test, nan included);
            # 2: finite and not nan, magnitude unknown (one test on the lanes' max |x|)
            c = 1 if r is None else (2 if r.hi > 6 else 0)
            if op in guarded_ops:
                out.check[k] = (c,)
                val[me] = R(MIN_NORM, MAX_EXP, neg=False)
            else:
                val[me] = None
        elif op == "erf":
            r = a[0]
            c = r is None                              # only nan needs the precise path
            if op in guarded_ops:
                out.check[k] = (c,)
                if r is None:
                    val[me] = R(MIN_SUB, 1, zero=True, nz=True)
                else:
                    val[me] = R(max(min(r.lo, -1) - 1, MIN_SUB), 1, r.zero or r.lo - 1 < MIN_SUB,
                                r.nz or (r.neg and r.lo - 1 < MIN_SUB), r.neg, r.pos)
            else:
                val[me] = None
        else:
            val[me] = None
    out.ranges = val
    return out
