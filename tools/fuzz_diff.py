"""Differential fuzzer: random NumPy programs evaluated through the engine and through plain NumPy
(= what the reference's CPU backend computes node by node, cpu.py:13-31) on the same inputs.

usage: python tools/fuzz_diff.py [--n 400] [--seed 0] [--only K] [-v]
Each program is a random expression over 1-3 array leaves (contiguous, sliced, strided, transposed
or broadcast; float32/float64/int32/int64/bool), Python and NumPy scalars, optionally ending in a
reduction.  Comparators follow BASELINE.json's north_star: arithmetic bit-exact, one transcendental
at the root within 2 ulp, reductions within rtol 1e-12 (float64) / 1e-5 (float32), integer and
boolean results exact.  Prints one line per failing program with a self-contained repro string.
"""
import argparse
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ARITH_BIN = ["add", "subtract", "multiply", "true_divide", "maximum", "minimum"]
CMP_BIN = ["greater", "less", "greater_equal", "less_equal", "equal", "not_equal"]
ARITH_UN = ["negative", "absolute", "square", "sqrt_abs", "reciprocal_safe", "floor", "sign"]
TRANS_UN = ["exp_b", "log_abs1", "sin", "cos", "tanh", "erf_", "arctan", "log1p_abs", "expm1_b"]
# extended set (seeds >= 1000, so that the programs of earlier seeds stay what they were)
EXT_BIN = ["fmax", "fmin", "fmod", "remainder", "floor_divide", "copysign"]
EXT_UN = ["ceil", "trunc", "rint", "isnan", "isfinite", "signbit", "logical_not", "positive"]
EXT_TRANS_UN = ["cbrt", "exp2_b", "log2_abs1", "log10_abs1", "arcsinh", "tan_b", "sinh_b", "cosh_b",
                "arcsin_c", "arccos_c", "arctanh_c", "erfc_"]
EXT_TRANS_BIN = ["hypot", "arctan2"]
INT_BIN = ["bitwise_and", "bitwise_or", "bitwise_xor", "floor_divide", "remainder", "add", "multiply",
           "subtract", "maximum", "left_shift_b", "right_shift_b"]
LOGIC_BIN = ["logical_and", "logical_or", "logical_xor"]
SIZES_1D = [0, 1, 2, 3, 4, 5, 7, 31, 32, 33, 127, 128, 129, 1000, 4096, 4097, 65537, 262147]
DTYPES = [np.float32, np.float64, np.float32, np.float64, np.int32, np.int64, np.bool_]


def leaf_values(rng, shape, dtype):
    if dtype == np.bool_:
        return rng.integers(0, 2, shape).astype(np.bool_)
    if np.issubdtype(dtype, np.integer):
        return rng.integers(-50, 50, shape).astype(dtype)
    x = rng.standard_normal(shape) * rng.choice([1e-3, 1.0, 1.0, 30.0])
    return x.astype(dtype)


class Prog:
    """expr: nested tuples; leaves: list of (base array, view recipe string)."""

    def __init__(self, rng, ext=False):
        self.rng = rng
        self.ext = ext
        self.leaves = []
        self.kind = "arith"
        ndim = int(rng.choice([1, 1, 1, 2, 2, 3]))
        if ndim == 1:
            self.shape = (int(rng.choice(SIZES_1D)),)
        elif ndim == 2:
            self.shape = (int(rng.choice([1, 2, 5, 33, 64, 300])), int(rng.choice([1, 3, 4, 17, 128, 257])))
        else:
            self.shape = (int(rng.choice([1, 2, 7])), int(rng.choice([1, 3, 16])), int(rng.choice([1, 4, 9, 32])))
        self.float_only = bool(rng.random() < 0.5)
        depth = int(rng.integers(1, 5))
        self.expr = self.gen(depth)
        r = rng.random()
        self.root = None
        if r < 0.2:
            if ext and rng.random() < 0.25:
                self.expr = ("bin", str(rng.choice(EXT_TRANS_BIN)), self.expr, self.gen(1))
            else:
                k = str(rng.choice(TRANS_UN + EXT_TRANS_UN if ext else TRANS_UN))
                self.expr = ("un", k, self.expr)
            self.kind = "trans"
        elif r < 0.5:
            red = str(rng.choice(["sum", "sum", "max", "min", "mean", "prod"]))
            axis = None
            if len(self.shape) > 1 and rng.random() < 0.6:
                axis = int(rng.integers(0, len(self.shape)))
            self.root = (red, axis)
            self.kind = "reduce"

    def new_leaf(self, ints=False):
        rng = self.rng
        dtype = rng.choice([np.float32, np.float64]) if self.float_only else DTYPES[int(rng.integers(len(DTYPES)))]
        if ints:
            dtype = [np.int32, np.int64, np.uint8, np.int16, np.uint32][int(rng.integers(5))]
        shape = list(self.shape)
        recipe = "c"
        r = rng.random()
        if r < 0.15 and len(shape) >= 1:
            # broadcast: drop leading dims or set some to 1
            k = int(rng.integers(0, len(shape)))
            shape = shape[k:]
            for i in range(len(shape)):
                if rng.random() < 0.4:
                    shape[i] = 1
            base = leaf_values(rng, tuple(shape), dtype)
        elif r < 0.3:
            recipe = "off"                                  # misaligned start: base[1:]
            base = leaf_values(rng, tuple([shape[0] + 1] + shape[1:]), dtype)
        elif r < 0.42:
            recipe = "step"                                 # strided along the last axis
            base = leaf_values(rng, tuple(shape[:-1] + [shape[-1] * 2]), dtype)
        elif r < 0.5 and len(shape) == 2:
            recipe = "T"
            base = leaf_values(rng, (shape[1], shape[0]), dtype)
        elif r < 0.56:
            recipe = "rev"
            base = leaf_values(rng, tuple(shape), dtype)
        else:
            base = leaf_values(rng, tuple(shape), dtype)
        self.leaves.append((base, recipe))
        return ("leaf", len(self.leaves) - 1)

    def scalar(self):
        rng = self.rng
        r = rng.random()
        if r < 0.4:
            return ("sc", float(rng.choice([0.5, 2.0, -1.5, 0.1, 3.0, 1e-3])))
        if r < 0.6:
            return ("sc", int(rng.choice([2, -3, 1, 0, 7])))
        if r < 0.8:
            return ("sc", np.float32(rng.choice([0.25, 1.7, -2.5])))
        return ("sc", np.float64(rng.choice([0.3, -4.0])))

    def gen_int(self, depth):
        """integer / boolean valued sub-expression (operands of bitwise, shift and logical ops)"""
        rng = self.rng
        if depth == 0:
            r = rng.random()
            if r < 0.6:
                return self.new_leaf(ints=True)
            if r < 0.8:
                return ("sc", int(rng.choice([1, 2, 3, 5, -2, 0])))
            return ("cmp", str(rng.choice(CMP_BIN)), self.gen(0), self.scalar())
        r = rng.random()
        if r < 0.7:
            a = self.gen_int(depth - 1)
            b = self.gen_int(depth - 1)
            if a[0] == "sc" and b[0] == "sc":
                a = self.new_leaf(ints=True)
            return ("bin", str(rng.choice(INT_BIN)), a, b)
        if r < 0.85:
            return ("bin", str(rng.choice(LOGIC_BIN)), self.gen_int(depth - 1), self.gen(depth - 1))
        return ("un", str(rng.choice(["invert", "negative", "absolute", "square"])), self.gen_int(depth - 1))

    def gen(self, depth):
        rng = self.rng
        if self.ext and depth > 0 and rng.random() < 0.3:
            r = rng.random()
            if r < 0.4:
                a, b = self.gen(depth - 1), (self.scalar() if rng.random() < 0.3 else self.gen(depth - 1))
                return ("bin", str(rng.choice(EXT_BIN)), a, b)
            if r < 0.7:
                return ("un", str(rng.choice(EXT_UN)), self.gen(depth - 1))
            return self.gen_int(min(depth, 2))
        if depth == 0 or (len(self.leaves) >= 3 and rng.random() < 0.3):
            if self.leaves and (len(self.leaves) >= 3 or rng.random() < 0.4):
                return ("leaf", int(rng.integers(len(self.leaves))))
            return self.new_leaf()
        r = rng.random()
        if r < 0.5:
            op = str(rng.choice(ARITH_BIN))
            a = self.gen(depth - 1)
            b = self.scalar() if rng.random() < 0.3 else self.gen(depth - 1)
            if rng.random() < 0.5:
                a, b = b, a
            if a[0] == "sc" and b[0] == "sc":
                a = self.gen(0)
            return ("bin", op, a, b)
        if r < 0.75:
            return ("un", str(rng.choice(ARITH_UN)), self.gen(depth - 1))
        if r < 0.85:
            return ("pow", self.gen(depth - 1), int(rng.choice([2, 3, 4])))
        if r < 0.95:
            c = ("bin", str(rng.choice(CMP_BIN)), self.gen(depth - 1), self.gen(0) if rng.random() < 0.5 else self.scalar())
            return ("where", c, self.gen(depth - 1), self.scalar() if rng.random() < 0.5 else self.gen(depth - 1))
        return ("cmp", str(rng.choice(CMP_BIN)), self.gen(depth - 1), self.scalar())


def view(a, recipe):
    if recipe == "off":
        return a[1:]
    if recipe == "step":
        return a[..., ::2]
    if recipe == "T":
        return a.T
    if recipe == "rev":
        return a[::-1]
    return a


def ev(e, leaves):
    t = e[0]
    if t == "leaf":
        return leaves[e[1]]
    if t == "sc":
        return e[1]
    if t == "bin" or t == "cmp":
        a, b = ev(e[2], leaves), ev(e[3], leaves)
        if e[1] in ("left_shift_b", "right_shift_b"):           # shift counts kept in [0, 7]
            return getattr(np, e[1][:-2])(a, np.bitwise_and(b, 7))
        return getattr(np, e[1])(a, b)
    if t == "pow":
        x = ev(e[1], leaves)
        if isinstance(x, np.ndarray):
            # the reference expands integer powers into a left-associated multiply chain
            # (delayarray.py:316-324 pow_ex), so that is what its CPU backend computes
            r = x
            for _ in range(e[2] - 1):
                r = np.multiply(r, x)
            return r
        return x ** e[2]
    if t == "where":
        return np.where(ev(e[1], leaves), ev(e[2], leaves), ev(e[3], leaves))
    if t == "un":
        x = ev(e[2], leaves)
        k = e[1]
        if k == "square" and isinstance(x, np.ndarray):
            return np.multiply(x, x)                        # create_ex: square -> multiply(x, x)
        if k == "sqrt_abs":
            return np.sqrt(np.absolute(x))
        if k == "reciprocal_safe":
            return 1.0 / (np.absolute(x) + 0.5)
        if k == "exp_b":
            return np.exp(np.minimum(x, 20))
        if k == "expm1_b":
            return np.expm1(np.minimum(x, 20))
        if k == "log_abs1":
            return np.log(np.absolute(x) + 1)
        if k == "log1p_abs":
            return np.log1p(np.absolute(x))
        if k == "exp2_b":
            return np.exp2(np.minimum(x, 20))
        if k in ("log2_abs1", "log10_abs1"):
            return getattr(np, k[:-5])(np.absolute(x) + 1)
        if k in ("tan_b", "sinh_b", "cosh_b"):
            return getattr(np, k[:-2])(np.minimum(np.maximum(x, -1.5), 1.5))
        if k in ("arcsin_c", "arccos_c", "arctanh_c"):
            return getattr(np, k[:-2])(np.minimum(np.maximum(x, -0.95), 0.95))
        if k == "erfc_":
            import scipy.special
            return scipy.special.erfc(x)
        if k == "erf_":
            import scipy.special
            return scipy.special.erf(x)
        return getattr(np, k)(x)
    raise ValueError(t)


def show(e):
    t = e[0]
    if t == "leaf":
        return f"L{e[1]}"
    if t == "sc":
        return f"{type(e[1]).__name__}({e[1]})"
    if t in ("bin", "cmp"):
        return f"{e[1]}({show(e[2])}, {show(e[3])})"
    if t == "pow":
        return f"({show(e[1])})**{e[2]}"
    if t == "where":
        return f"where({show(e[1])}, {show(e[2])}, {show(e[3])})"
    return f"{e[1]}({show(e[2])})"


def ulps(got, want):
    got = np.asarray(got)
    want = np.asarray(want)
    fin = np.isfinite(want) & np.isfinite(got)
    if not np.array_equal(np.isnan(got), np.isnan(want)):
        return np.inf
    if not np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)]):
        return np.inf
    if not fin.any():
        return 0.0
    sp = np.spacing(np.abs(want[fin]).astype(want.dtype))
    return float(np.max(np.abs(got[fin].astype(np.float64) - want[fin].astype(np.float64)) / sp))


def _signbit_of_nan(e, leaves):
    """True when some `signbit` in the expression is applied to a nan: the sign of a generated nan
    is the platform's (x86 fmod(x, 0) gives -nan, CUDA +nan), so its signbit cannot be compared."""
    if not isinstance(e, tuple) or not e:
        return False
    if e[0] == "un" and e[1] == "signbit":
        with np.errstate(all="ignore"):
            v = np.asarray(ev(e[2], leaves))
        if v.dtype.kind == "f" and np.isnan(v).any():
            return True
    return any(_signbit_of_nan(k, leaves) for k in e[1:] if isinstance(k, tuple))


def _ambiguous_fminmax(e, leaves):
    """True when some fmax / fmin in the expression meets -0.0 against +0.0: NumPy's vector body
    returns the second operand there and its scalar tail the first, so the sign of that zero (and
    of everything computed from it: 1 / 0, arctan2(0, -1)) depends on the element's position."""
    if not isinstance(e, tuple) or not e:
        return False
    if e[0] == "bin" and e[1] in ("fmax", "fmin"):
        with np.errstate(all="ignore"):
            x, y = np.asarray(ev(e[2], leaves)), np.asarray(ev(e[3], leaves))
        if x.dtype.kind == "f" or y.dtype.kind == "f":
            x, y = np.broadcast_arrays(x.astype(np.float64), y.astype(np.float64))
            if ((x == 0) & (y == 0) & (np.signbit(x) != np.signbit(y))).any():
                return True
    return any(_ambiguous_fminmax(k, leaves) for k in e[1:] if isinstance(k, tuple))


def run_one(seed, verbose=False, dry=False):
    import delayrepay_b200 as dr
    rng = np.random.default_rng(seed)
    p = Prog(rng, ext=seed >= 1000)
    desc = f"seed={seed} shape={p.shape} kind={p.kind} root={p.root} expr={show(p.expr)} leaves=" + \
        ",".join(f"{b.dtype}{list(b.shape)}:{r}" for b, r in p.leaves)
    if verbose:
        print(desc, flush=True)
    with np.errstate(all="ignore"):
        try:
            want = ev(p.expr, [view(b, r) for b, r in p.leaves])
            if p.root:
                want = getattr(np, p.root[0])(want, axis=p.root[1])
        except Exception:                                       # noqa: BLE001
            return None                 # not a valid NumPy program (no loop, empty max, ...)
    try:
        got = ev(p.expr, [view(dr.array(b), r) for b, r in p.leaves])
        if p.root:
            got = getattr(np, p.root[0])(got, axis=p.root[1])
        if dry:
            if hasattr(got, "run"):
                got.run()
            return None
        got = got.get() if hasattr(got, "get") else np.asarray(got)
    except Exception as ex:                                     # noqa: BLE001
        if isinstance(ex, TypeError) and "float16" in str(ex):
            return None
        tb = traceback.extract_tb(ex.__traceback__)[-1]
        return f"EXC {type(ex).__name__}: {str(ex)[:200]} @ {os.path.basename(tb.filename)}:{tb.lineno} | {desc}"
    want = np.asarray(want)
    got = np.asarray(got)
    if want.dtype == np.float16:
        return None                                         # float16 loops: not supported, documented
    if "signbit(" in desc and _signbit_of_nan(p.expr, [view(b, r) for b, r in p.leaves]):
        return None
    if ("fmax(" in desc or "fmin(" in desc) and _ambiguous_fminmax(p.expr, [view(b, r) for b, r in p.leaves]):
        return None
    if got.shape != want.shape:
        return f"SHAPE got {got.shape} want {want.shape} | {desc}"
    if got.dtype != want.dtype:
        return f"DTYPE got {got.dtype} want {want.dtype} | {desc}"
    if want.dtype.kind in "biu":
        if p.kind == "trans" or not np.array_equal(got, want):
            if not np.array_equal(got, want):
                return (f"VALUE(int) mismatches={int((got != want).sum())} got={got.ravel()[:3]} "
                        f"want={want.ravel()[:3]} | {desc}")
        return None
    if p.kind == "arith":
        if got.tobytes() != want.tobytes():
            # nan payloads / signs may differ legitimately: compare with nan-equality
            # (the sign and payload of a nan are the platform's: x86 0/0 is -nan, CUDA's +nan)
            ok = ~np.isnan(want)
            if "fmax(" in desc or "fmin(" in desc:
                ok &= want != 0         # NumPy's fmax/fmin: SIMD body and scalar tail disagree on -0 vs +0
            same = np.array_equal(got, want, equal_nan=True) and \
                np.array_equal(np.signbit(got[ok]), np.signbit(want[ok])) if want.dtype.kind == "f" else False
            if not same:
                return f"VALUE(bits) ulps={ulps(got, want):.3g} | {desc}"
        return None
    if p.kind == "trans":
        u = ulps(got, want)
        # two float64 erf implementations (CUDA libm: 2 ulp, SciPy/xsf: ~1 ulp) can be 3 apart
        # (tests/test_parity_gpu.py checks those cases against mpmath)
        lim = 3.0 if (want.dtype == np.float64 and p.expr[1] == "erf_") else 2.0
        if u > lim and p.expr[1] == "erfc_":
            # SciPy's erfc (cephes) is up to ~500 ulp off for large arguments in float64 (measured
            # against mpmath); CUDA's is documented at 5 ulp: judge against the true value
            import mpmath
            with np.errstate(all="ignore"):
                x = np.broadcast_to(np.asarray(ev(p.expr[2], [view(b, r) for b, r in p.leaves])), want.shape)
            x = x.astype(want.dtype).ravel()
            g = got.ravel()
            idx = np.argsort(-np.abs(g.astype(np.float64) - want.ravel().astype(np.float64))
                             / np.spacing(np.abs(want.ravel()).astype(want.dtype)))[:64]
            mpmath.mp.prec = 200
            truth = np.array([float(mpmath.erfc(mpmath.mpf(float(v)))) for v in x[idx]]).astype(want.dtype)
            u, lim = ulps(g[idx], truth), 6.0
        if u > lim and want.dtype == np.float32 and p.expr[0] in ("un", "bin"):
            # NumPy's float32 loops (SVML) are themselves up to ~3 ulp from the truth for a few
            # functions (arcsin near 0.95, arctan2): judge against the float64 evaluation, rounded once
            with np.errstate(all="ignore"):
                lv = [view(b, r) for b, r in p.leaves]
                args = [("sc", np.asarray(ev(e, lv)).astype(np.float64)) for e in p.expr[2:]]
                truth = np.asarray(ev((p.expr[0], p.expr[1]) + tuple(args), []))
            truth = np.broadcast_to(truth, want.shape).astype(np.float32)
            u, lim = ulps(got, truth), 1.0
        if u > lim:
            return f"VALUE(ulp) ulps={u:.3g} | {desc}"
        return None
    rtol = 1e-12 if want.dtype == np.float64 else 1e-5
    if p.root[0] in ("max", "min"):
        ok = np.array_equal(got, want, equal_nan=True)
    else:
        with np.errstate(all="ignore"):
            # scale by the sum of magnitudes (cancellation is not the engine's error)
            mag = ev(("un", "absolute", p.expr), [view(b, r) for b, r in p.leaves])
            mag = np.asarray(getattr(np, "sum" if p.root[0] != "prod" else "prod")(
                mag, axis=p.root[1])) if p.root[0] != "mean" else np.asarray(np.mean(mag, axis=p.root[1]))
        if p.root[0] == "prod":
            with np.errstate(all="ignore"):         # the same, leaving exact zeros out (0 * inf)
                nz = np.asarray(ev(("un", "absolute", p.expr), [view(b, r) for b, r in p.leaves])).astype(np.float64)
                mag_nz = np.prod(np.where(nz == 0, 1.0, nz), axis=p.root[1])
        if p.root[0] == "prod" and np.any((np.abs(want.astype(np.float64)) < np.finfo(want.dtype).tiny) & (want != 0)):
            return None             # a subnormal product: how many bits survive the underflow depends on the order
        if p.root[0] == "prod" and not (np.all(np.isfinite(mag)) and np.all(np.isfinite(mag_nz))):
            # the product of the magnitudes overflows: whether a partial product reaches inf before
            # it meets a zero (0 * inf = nan) depends on the ORDER of the multiplications, which a
            # parallel reduction does not share with NumPy's sequential loop
            return None
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        fin = np.isfinite(want)
        ok = np.array_equal(np.isnan(got), np.isnan(want)) and np.all(err[fin] <= rtol * np.maximum(np.abs(mag[fin]), 1e-300)) \
            and np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)])
    if not ok:
        return f"VALUE(reduce) got={got.ravel()[:3]} want={want.ravel()[:3]} | {desc}"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=400)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--only", type=int, default=None)
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--dry", action="store_true", help="no GPU: plan, generate and compile only")
    a = ap.parse_args()
    import contextlib
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine
    ctx = engine.dry_run() if a.dry else contextlib.nullcontext()
    if not a.dry:
        dr.set_device(0)
    seeds = [a.only] if a.only is not None else range(a.seed, a.seed + a.n)
    bad = 0
    ctx.__enter__()
    for s in seeds:
        try:
            msg = run_one(s, a.v, a.dry)
        except Exception as ex:                                 # noqa: BLE001  (oracle side failed)
            msg = None
            if a.v:
                print(f"skip seed={s}: oracle raised {type(ex).__name__}: {ex}")
        if msg:
            bad += 1
            print(msg, flush=True)
    print(f"fuzz: {bad} failing of {len(list(seeds))}")


if __name__ == "__main__":
    main()
