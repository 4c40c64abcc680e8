"""Parity of the CUDA path against the oracle (oracle/refcpu.py == the reference's CPU backend)
on identical seeded inputs, and against the committed golden fixtures.

Bars (BASELINE.json north_star): elementwise arithmetic BIT-EXACT; transcendentals <= 2 ulp;
reductions rtol 1e-12 (fp64) / 1e-5 (fp32).  Chains that mix both (Black-Scholes, n-body) are
held to a bound derived from those (stated where used) and must be at least as accurate as the
reference against a float64 evaluation.
"""
import os

import numpy as np
import pytest
from scipy.special import erf

import workloads as wl
from oracle import refcpu
from util import (assert_bits_equal, assert_close_to_numpy_or_truth, assert_ulp, erf_exact,
                  ulp_distance)

pytestmark = pytest.mark.gpu
GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "workloads.npz"))
EPS32 = float(np.finfo(np.float32).eps)


# ------------------------------------------------------------------ C1: axpy, bit-exact
@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 2048, (1 << 20) + 3, 1 << 24])
def test_axpy_bit_exact(gpu, n):
    i = wl.make_inputs("axpy", n)
    got = wl.axpy(gpu, i["a"], gpu.array(i["x"]), gpu.array(i["y"])).get()
    want = wl.axpy(refcpu, i["a"], refcpu.leaf(i["x"]), refcpu.leaf(i["y"])).get()
    assert_bits_equal(got, want, f"axpy n={n}")
    if n == 2048:
        assert_bits_equal(got, GOLDEN["axpy"], "axpy golden")


def test_axpy_unaligned_and_strided_views_bit_exact(gpu):
    i = wl.make_inputs("axpy", 10007)
    x, y = gpu.array(i["x"]), gpu.array(i["y"])
    for sl in (slice(1, None), slice(3, -2), slice(None, None, 2), slice(None, None, -1),
               slice(5, 9001, 3)):
        got = (1.5 * x[sl] + y[sl]).get()
        assert_bits_equal(got, 1.5 * i["x"][sl] + i["y"][sl], f"slice {sl}")


# ------------------------------------------------------------------ arithmetic, bit-exact
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_arithmetic_chain_bit_exact(gpu, dt):
    rng = np.random.default_rng(11)
    a, b, c = (rng.standard_normal(100003).astype(dt) for _ in range(3))
    c = np.abs(c) + dt(0.5)

    def f(xp, a, b, c):
        return (a * b + c) / (c - a * 0.25) - xp.sqrt(c) * (a - b) / c + abs(a) * 3
    got = f(gpu, gpu.array(a), gpu.array(b), gpu.array(c)).get()
    want = f(refcpu, refcpu.leaf(a), refcpu.leaf(b), refcpu.leaf(c)).get()
    assert_bits_equal(got, want, "arith chain")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_integer_power_chain_matches_reference_association(gpu, dt):
    rng = np.random.default_rng(6)
    if dt is np.float64:
        rng.standard_normal(2048)
    x = rng.standard_normal(2048).astype(dt)
    name = np.dtype(dt).name
    assert_bits_equal((gpu.array(x) ** 3).get(), GOLDEN[f"pow3_{name}"], "x**3")
    assert_bits_equal((gpu.array(x) ** 5).get(), GOLDEN[f"pow5_{name}"], "x**5")
    assert_bits_equal((gpu.array(x) ** 2).get(), x * x, "x**2")
    assert_bits_equal(np.square(gpu.array(x)).get(), np.square(x), "square")


def test_scalar_typing_follows_nep50(gpu):
    x = np.random.default_rng(3).standard_normal(4099).astype(np.float32)
    assert_bits_equal((gpu.array(x) * 0.1).get(), x * 0.1, "f32 * python float")
    assert_bits_equal((gpu.array(x) * np.float64(0.1)).get(), x * np.float64(0.1), "f32 * f64")
    assert_bits_equal((gpu.array(x) + 7).get(), x + 7, "f32 + int")
    xi = np.arange(-50, 50, dtype=np.int64)
    assert_bits_equal((gpu.array(xi) * 0.5).get(), xi * 0.5, "i64 * float")
    assert_bits_equal((gpu.array(xi) / 3).get(), xi / 3, "i64 / int")
    assert_bits_equal((gpu.array(xi) // 7).get(), xi // 7, "i64 // int")
    assert_bits_equal((gpu.array(xi) % 7).get(), xi % 7, "i64 % int")
    assert_bits_equal((gpu.array(x) > 0.25).get(), x > 0.25, "compare")
    assert_bits_equal(gpu.array(x).astype(np.float64).get(), x.astype(np.float64), "astype")
    inf = gpu.array(x) * float("inf")
    assert_bits_equal(inf.get(), x * float("inf"), "inf scalar (a literal would not compile)")


# ------------------------------------------------------------------ transcendentals, <= 2 ulp
UNARY = [("exp", np.exp, (-20, 20)), ("log", np.log, (1e-3, 1e3)), ("sin", np.sin, (-20, 20)),
         ("cos", np.cos, (-20, 20)), ("tan", np.tan, (-1.5, 1.5)), ("tanh", np.tanh, (-5, 5)),
         ("sinh", np.sinh, (-8, 8)), ("cosh", np.cosh, (-8, 8)), ("erf", erf, (-4, 4)),
         ("sqrt", np.sqrt, (0, 1e6)), ("arctan", np.arctan, (-50, 50)),
         ("log1p", np.log1p, (-0.9, 10)), ("expm1", np.expm1, (-5, 5))]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("name,fn,rng_", UNARY, ids=[u[0] for u in UNARY])
def test_transcendental_within_2ulp(gpu, dt, name, fn, rng_):
    x = np.random.default_rng(17).uniform(rng_[0], rng_[1], 1 << 20).astype(dt)
    got = fn(gpu.array(x)).get()
    want = fn(x)
    assert got.dtype == want.dtype
    limit = 0 if name == "sqrt" else 2
    d = ulp_distance(got, want)
    print(f"{name:6s} {np.dtype(dt).name}: max {int(d.max())} ulp vs NumPy, "
          f"{(d == 0).mean():.4f} bit-identical")
    if dt is np.float64 and name == "erf" and d.max() > limit:
        # two double erf implementations (CUDA libm: 2 ulp; SciPy/xsf: ~1 ulp) can be 3 apart;
        # settle the disagreeing points against a 40-digit evaluation: ours must be <= 2 ulp
        bad = np.flatnonzero(d > limit)[:64]
        from decimal import Decimal
        for i, t in zip(bad, erf_exact(x[bad])):
            err = abs(Decimal(float(got[i])) - t) / Decimal(float(np.spacing(abs(got[i]))))
            assert err <= 2, (x[i], float(err))
        return
    assert_close_to_numpy_or_truth(got, want, fn, (x,), limit, f"{name} {np.dtype(dt).name}")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_binary_functions_within_2ulp(gpu, dt):
    rng = np.random.default_rng(19)
    a = rng.uniform(0.1, 50, 1 << 18).astype(dt)
    b = rng.uniform(-3, 3, 1 << 18).astype(dt)
    assert_close_to_numpy_or_truth(np.arctan2(gpu.array(b), gpu.array(a)).get(), np.arctan2(b, a),
                                   np.arctan2, (b, a), 2, "arctan2")
    assert_close_to_numpy_or_truth(np.power(gpu.array(a), gpu.array(b)).get(), np.power(a, b),
                                   np.power, (a, b), 2, "power")
    assert_ulp((gpu.array(a) ** -1.5).get(), a ** -1.5, 2, "x**-1.5")
    assert_ulp((gpu.array(a) ** 0.5).get(), a ** 0.5, 0, "x**0.5")
    assert_ulp(np.hypot(gpu.array(a), gpu.array(b)).get(), np.hypot(a, b), 2, "hypot")
    assert_bits_equal(np.maximum(gpu.array(a), gpu.array(b)).get(), np.maximum(a, b), "maximum")


def test_special_values(gpu):
    x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, 88.8, -104.0, 1.0],
                 dtype=np.float32)
    with np.errstate(all="ignore"):
        for fn in (np.exp, np.log, np.sqrt, np.tanh, erf, np.abs, np.negative, np.sign):
            got, want = fn(gpu.array(x)).get(), fn(x)
            assert np.array_equal(np.isnan(got), np.isnan(want)), fn.__name__
            assert_ulp(got, want, 2, fn.__name__)
        assert_bits_equal(np.maximum(gpu.array(x), 0.5).get(), np.maximum(x, np.float32(0.5)), "max nan")
        assert_bits_equal(np.isnan(gpu.array(x)).get(), np.isnan(x), "isnan")
        assert_bits_equal((gpu.array(x) / gpu.array(x[::-1].copy())).get(), x / x[::-1], "div")


def test_division_sqrt_log_on_random_bit_patterns(gpu):
    """The branch-free fast paths + flagged precise re-evaluation must stay IEEE-exact on
    every class of operand: random BIT patterns cover subnormals, huge values, inf, nan, 0."""
    rng = np.random.default_rng(41)
    a = rng.integers(0, 1 << 32, 1 << 21, dtype=np.uint32).view(np.float32)
    b = rng.integers(0, 1 << 32, 1 << 21, dtype=np.uint32).view(np.float32)
    with np.errstate(all="ignore"):
        assert_bits_equal((gpu.array(a) / gpu.array(b)).get(), a / b, "div bits")
        assert_bits_equal(np.sqrt(gpu.array(a)).get(), np.sqrt(a), "sqrt bits")
        pos = np.abs(a)
        got, want = np.log(gpu.array(pos)).get(), np.log(pos)
        assert_close_to_numpy_or_truth(got[np.isfinite(want)], want[np.isfinite(want)], np.log,
                                       (pos[np.isfinite(want)],), 2, "log bits")
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
        # moderate magnitudes (the fast path proper), many samples
        c = rng.uniform(-1e3, 1e3, 1 << 22).astype(np.float32)
        d = rng.uniform(1e-3, 1e3, 1 << 22).astype(np.float32)
        assert_bits_equal((gpu.array(c) / gpu.array(d)).get(), c / d, "div moderate")
        assert_bits_equal(np.sqrt(gpu.array(d)).get(), np.sqrt(d), "sqrt moderate")
        assert_bits_equal((1.0 / gpu.array(d)).get(), (1.0 / d).astype(np.float32), "reciprocal")


def _hostile(rng, n, positive=False):
    """Mostly moderate values with every special class sprinkled in."""
    x = rng.uniform(0.01, 100.0, n).astype(np.float32)
    if not positive:
        x *= rng.choice(np.array([-1, 1], np.float32), n)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-42, 3.4e38, -3e38, 1e-38,
                        2.0 ** -31, 2.0 ** 30, 2.0 ** -30, 1.0, -1.0, 5e-324], dtype=np.float32)
    idx = rng.choice(n, n // 16, replace=False)
    x[idx] = rng.choice(special, idx.size)
    return x


def test_guard_placement_keeps_arithmetic_exact_on_hostile_inputs(gpu):
    """The interval analysis (ranges.py) drops per-operation range tests where it can prove the
    fast forms' preconditions from ONE combined test of the inputs; a vector that fails the
    combined test must take the precise path.  Exact operations must stay bit-exact for zeros,
    negatives, subnormals, huge values, inf and nan anywhere in the inputs."""
    rng = np.random.default_rng(77)
    n = (1 << 18) + 3
    a, b, c = _hostile(rng, n), _hostile(rng, n), _hostile(rng, n, positive=True)

    def f(xp, a, b, c):
        q = (a / b + c * 0.5) / (0.3 * xp.sqrt(c))          # the Black-Scholes d1 shape
        return q, q - 0.3 * xp.sqrt(c), xp.sqrt(a / b)
    with np.errstate(all="ignore"):
        got = f(gpu, gpu.array(a), gpu.array(b), gpu.array(c))
        gpu.evaluate(*got)
        want = f(np, a, b, c)
        for g, w, nm in zip(got, want, ("d1-like", "d2-like", "sqrt(a/b)")):
            assert_bits_equal(g.get(), w.astype(np.float32), nm)


def test_generation2_exp_log_erf_over_their_whole_domains(gpu):
    rng = np.random.default_rng(5)
    n = 1 << 20
    with np.errstate(all="ignore"):
        x = np.concatenate([rng.uniform(-104, 89, n), rng.uniform(-1, 1, n // 4)]).astype(np.float32)
        assert_close_to_numpy_or_truth(np.exp(gpu.array(x)).get(), np.exp(x), np.exp, (x,), 2, "exp wide")
        x = np.exp(rng.uniform(-87, 88, n)).astype(np.float32)
        x = np.concatenate([x, rng.uniform(0.5, 2.0, n // 2).astype(np.float32),
                            np.float32(1) + np.arange(-2048, 2048, dtype=np.float32) * np.float32(2.0 ** -24)])
        assert_close_to_numpy_or_truth(np.log(gpu.array(x)).get(), np.log(x), np.log, (x,), 2, "log wide")
        x = (np.exp(rng.uniform(-60, 2.5, n)) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
        x = np.concatenate([x, np.linspace(-6, 6, n // 2, dtype=np.float32),
                            np.array([0.0, -0.0, 1e-45, -1e-45, 4.0, 3.9999998, 1e30, -1e30], np.float32)])
        got, want = erf(gpu.array(x)).get(), erf(x)
        assert_close_to_numpy_or_truth(got, want, erf, (x,), 2, "erf wide")
        assert np.array_equal(np.signbit(got), np.signbit(want)), "erf sign (incl. -0)"
        # inside a fused chain the table forms see arbitrary finite intermediates
        y = rng.uniform(-30, 30, n).astype(np.float32)
        got = erf(np.exp(gpu.array(y) * 0.1) - 2.0).get()
        truth = erf(np.exp(y.astype(np.float64) * np.float64(np.float32(0.1))) - 2.0)
        # the subtraction amplifies exp's last-place error; bound the absolute error instead
        np.testing.assert_allclose(got, truth, rtol=0, atol=2e-6)


@pytest.mark.parametrize("shape", [(300, 200, 150), (64, 1000, 64), (1025, 33, 129), (2048, 512, 64)])
def test_tiled_gemm_float64_and_integers(gpu, shape):
    """float64 / integer A @ B (no tensor-core path): the register-tiled kernel (gemm.matmul_tiled).
    float64 within rtol 1e-12 of the terms' scale, integers exact including wrap-around."""
    from delayrepay_b200 import engine
    m, k, n = shape
    rng = np.random.default_rng(m + k + n)
    a, b = rng.standard_normal((m, k)), rng.standard_normal((k, n))
    got = (gpu.array(a) @ gpu.array(b)).get()
    assert engine.last_kernel_name().startswith("dr_gemm_tiled_")
    scale = np.abs(a) @ np.abs(b)
    assert np.max(np.abs(got - a @ b) / scale) <= 1e-12
    # fused producers and a transposed (strided) operand
    got = ((gpu.array(a) * 2 + 1) @ gpu.array(b.T.copy()).T).get()
    assert np.max(np.abs(got - (a * 2 + 1) @ b) / (np.abs(a * 2 + 1) @ np.abs(b))) <= 1e-12
    for dt in (np.int32, np.int64, np.uint8):
        info = np.iinfo(dt)
        ia = rng.integers(info.min // 2 if dt is not np.uint8 else 0, info.max // 2, (m, k)).astype(dt)
        ib = rng.integers(0, 7, (k, n)).astype(dt)
        with np.errstate(over="ignore"):
            want = ia @ ib                                  # wraps
        assert_bits_equal((gpu.array(ia) @ gpu.array(ib)).get(), want, f"{np.dtype(dt).name} matmul")


def test_generation3_erf_in_staged_kernels(gpu):
    """The staged (TMA-ring) kernels with a heavy body -- Black-Scholes' class -- use the third-
    generation erf (tools/gen_erf3.py: table uniform in sqrt(|x|/4), indexed through MUFU.SQRT and
    a magic-constant add).  `erf(x) / 1` makes such a kernel and returns erf(x) unchanged."""
    from delayrepay_b200 import engine
    rng = np.random.default_rng(55)
    n = 1 << 20
    with np.errstate(all="ignore"):
        x = (np.exp(rng.uniform(-60, 2.5, n)) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
        x = np.concatenate([x, np.linspace(-6, 6, n // 2, dtype=np.float32),
                            rng.uniform(-2.0 ** -7, 2.0 ** -7, n // 4).astype(np.float32),
                            (10.0 ** rng.uniform(-45.5, -30, n // 4) * rng.choice([-1.0, 1.0], n // 4)).astype(np.float32),
                            ((np.arange(0, 260, dtype=np.float64) / 256.0) ** 2 * 4).astype(np.float32),      # row centres
                            (((np.arange(0, 260, dtype=np.float64) + 0.5) / 256.0) ** 2 * 4).astype(np.float32),  # row borders
                            np.array([0.0, -0.0, 1e-45, -1e-45, 4.0, 3.9999998, 4.0000005, 1e30, -1e30, np.inf, -np.inf,
                                      np.nan], np.float32)])
        pad = (-x.size) % 4
        x = np.concatenate([x, np.zeros(pad, np.float32)])
        got = (erf(gpu.array(x)) / gpu.array(np.ones_like(x))).get()
        assert "dr_erf4_s" in [k for k in engine._kernels.values() if k.name == engine.last_kernel_name()][0].source
        want = erf(x)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        worst = assert_close_to_numpy_or_truth(got[ok], want[ok], erf, (x[ok],), 2, "erf generation 3")
        assert np.array_equal(np.signbit(got[ok]), np.signbit(want[ok])), "erf sign (incl. -0)"
        truth = erf(x[ok].astype(np.float64))
        ulp = np.abs(got[ok].astype(np.float64) - truth) / np.spacing(np.abs(truth).astype(np.float32)).astype(np.float64)
        ulp[truth == 0] = 0
        print(f"   erf generation 3: max {ulp.max():.3f} ulp from float64 (|x| >= 2^-9: "
              f"{ulp[np.abs(x[ok]) >= 2.0 ** -9].max():.3f}), {worst} ulp from NumPy")
        assert ulp.max() <= 1.5 and ulp[np.abs(x[ok]) >= 2.0 ** -9].max() <= 0.70


def test_plan_cache_replays_are_exact(gpu):
    """engine._plans: a repeated expression structure launches a prepared plan (no planning);
    values, operand sharing patterns, layouts and scalar classes may change between replays."""
    from delayrepay_b200 import engine
    rng = np.random.default_rng(123)
    n = 4099
    x, y, z = (rng.standard_normal(n) for _ in range(3))
    gx, gy, gz = gpu.array(x), gpu.array(y), gpu.array(z)
    hits0 = engine.stats.get("plan_hits", 0)
    for a, b in ((1.5, 1.5), (2.0, -3.0), (0.25, 0.25), (1e300, 2.0), (-1.0, 7.0)):
        assert_bits_equal((a * gx + b * gy).get(), a * x + b * y, f"a*x+b*y a={a} b={b}")
        assert_bits_equal((a * gx + b * gx).get(), a * x + b * x, "same leaf twice")
        assert_bits_equal((a * gz + b * gy).get(), a * z + b * y, "other leaves, same layout")
    assert engine.stats.get("plan_hits", 0) > hits0 + 6
    # float32 with guarded operations: the scalar's magnitude class is part of the plan
    f = rng.uniform(0.5, 4.0, n).astype(np.float32)
    gf = gpu.array(f)
    for s in (0.3, 0.7, 1e-30, 3.0, 1e30, 0.3):
        with np.errstate(all="ignore"):
            assert_bits_equal((np.sqrt(gf) / (s * gf)).get(), (np.sqrt(f) / (np.float32(s) * f)), f"s={s}")
    # layouts: slices (other offsets / strides), other sizes, in-place leaf astype
    for sl in (slice(0, None), slice(1, None), slice(None, None, 2), slice(3, 1000)):
        assert_bits_equal((1.5 * gx[sl] + gy[sl]).get(), 1.5 * x[sl] + y[sl], f"slice {sl}")
    leaf = gpu.array(x)
    assert_bits_equal((leaf * 2.0 + 1.0).get(), x * 2.0 + 1.0, "f64 leaf")
    leaf.astype(np.float32)
    assert_bits_equal((leaf * 2.0 + 1.0).get(), x.astype(np.float32) * np.float32(2.0) + np.float32(1.0),
                      "same leaf after in-place astype")
    # several roots in one kernel, replayed with a different sharing pattern between the roots
    for a, b in ((2.0, 3.0), (2.0, 2.0), (5.0, 3.0)):
        r1, r2 = a * gx + gy, b * gx - gy
        gpu.evaluate(r1, r2)
        assert_bits_equal(r1.get(), a * x + y, "root 1")
        assert_bits_equal(r2.get(), b * x - y, "root 2")
    # a materialised producer is reused (cut point), and mutation invalidates memoised results
    d = gx * gy
    d.run()
    assert_bits_equal((d + 1.0).get(), x * y + 1.0, "materialised producer")
    gx2 = gpu.array(x.copy())
    r = gx2 * 3.0
    assert_bits_equal(r.get(), x * 3.0, "before mutation")
    gx2[0] = 100.0
    xm = x.copy(); xm[0] = 100.0
    assert_bits_equal((gx2 * 3.0).get(), xm * 3.0, "after mutation")


@pytest.mark.parametrize("n", [5, 63, 64, 65, 257, 4099, (1 << 18) + 7])
def test_staged_kernels_all_dtypes_and_sizes(gpu, n):
    """The per-warp TMA rings (two vectors per lane per stage) with partial stages, partial
    vectors and scalar tails, for float32 (lockstep), float64 (two lanes per vector) and a fused
    full reduction over a heavy body."""
    rng = np.random.default_rng(n)
    x = rng.uniform(0.5, 3.0, n)
    y = rng.uniform(0.5, 3.0, n)
    for dt, tol in ((np.float64, 4), (np.float32, 6)):
        a, b = x.astype(dt), y.astype(dt)
        got = (np.exp(gpu.array(a) * 0.5) * np.log(gpu.array(b)) / np.sqrt(gpu.array(a))).get()
        want = np.exp(a * dt(0.5)) * np.log(b) / np.sqrt(a)
        assert got.dtype == want.dtype
        assert_ulp(got, want, tol, f"heavy chain {np.dtype(dt).name} n={n}")
        s_got = float(np.sum(np.exp(gpu.array(a) * 0.5) * np.log(gpu.array(b))))
        s_want = float(np.sum(np.exp(a.astype(np.float64) * 0.5) * np.log(b.astype(np.float64))))
        rtol = 1e-12 if dt is np.float64 else 1e-5
        assert abs(s_got - s_want) <= rtol * max(abs(s_want), np.abs(np.log(b)).sum() * 1e-3), (dt, n)
    c, d = gpu.array(x.astype(np.float32)), gpu.array(y.astype(np.float32))
    got = erf(c - 1.5) * np.exp(-d)
    np.testing.assert_allclose(got.get(), erf(x.astype(np.float32) - np.float32(1.5)) * np.exp(-y.astype(np.float32)),
                               rtol=1e-6, atol=1e-7)


def test_nd_kernels_with_inner_vectors_bit_exact(gpu):
    """Strided / broadcast regions whose innermost dimension is contiguous are walked in 128-bit
    vectors (planner._try_inner_vectors); everything else stays on the scalar nd kernel.  Both
    must agree with NumPy bit for bit."""
    rng = np.random.default_rng(31)
    for dt in (np.float32, np.float64):
        X = rng.standard_normal((257, 1024)).astype(dt)
        mu = rng.standard_normal(1024).astype(dt)
        sd = rng.uniform(0.5, 2.0, 257).astype(dt)
        gX, gmu, gsd = gpu.array(X), gpu.array(mu), gpu.array(sd)
        assert_bits_equal(((gX - gmu[None, :]) / gsd[:, None]).get(), (X - mu[None, :]) / sd[:, None],
                          f"row and column broadcast {np.dtype(dt).name}")
        assert_bits_equal((gX[1:-1, 4:-4] * 2.0 + gX[2:, 4:-4]).get(), X[1:-1, 4:-4] * 2.0 + X[2:, 4:-4],
                          "aligned strided views")
        assert_bits_equal((gX[1:-1, 1:-3] * 2.0 + gX[2:, 3:-1]).get(), X[1:-1, 1:-3] * 2.0 + X[2:, 3:-1],
                          "misaligned views (scalar nd)")
        assert_bits_equal((gX[:, ::2] + 1.0).get(), X[:, ::2] + 1.0, "strided inner (scalar nd)")
        assert_bits_equal((gX.T * 3.0).get(), X.T * 3.0, "transposed (scalar nd)")
        assert_bits_equal((gX[:, :1022] - gmu[None, :1022]).get(), X[:, :1022] - mu[None, :1022],
                          "inner extent not a multiple of 4 for f32")
        a, b = gX[:200, :512] + gmu[None, :512], gX[:200, 512:] * gsd[:200, None]
        gpu.evaluate(a, b)
        assert_bits_equal(a.get(), X[:200, :512] + mu[None, :512], "two outputs, first")
        assert_bits_equal(b.get(), X[:200, 512:] * sd[:200, None], "two outputs, second")
        T3 = rng.standard_normal((6, 33, 64)).astype(dt)
        g3 = gpu.array(T3)
        assert_bits_equal((g3 * gpu.array(mu[:64])[None, None, :] + g3[:, :1, :]).get(),
                          T3 * mu[:64][None, None, :] + T3[:, :1, :], "3-d with a broadcast middle axis")
    Xi = rng.integers(-100, 100, (64, 256)).astype(np.int32)
    assert_bits_equal((gpu.array(Xi) + gpu.array(Xi[0])[None, :]).get(), Xi + Xi[0][None, :], "int32")
    assert_bits_equal((gpu.array(X.astype(np.float32)) + gpu.array(sd)[:, None]).get(),
                      X.astype(np.float32) + sd[:, None], "mixed widths (scalar nd)")


# ------------------------------------------------------------------ C2: Black-Scholes
def _bs_truth(S, K, T, r=0.02, v=0.30):
    S, K, T = (a.astype(np.float64) for a in (S, K, T))
    return wl.black_scholes(np, S, K, T, r, v)


@pytest.mark.parametrize("n", [2048, (1 << 20) + 1])
def test_black_scholes_parity(gpu, n):
    i = wl.make_inputs("black_scholes", n)
    call, put = wl.black_scholes(gpu, *(gpu.array(i[k]) for k in ("S", "K", "T")))
    gpu.evaluate(call, put)
    rc, rp = wl.black_scholes(refcpu, *(refcpu.leaf(i[k]) for k in ("S", "K", "T")))
    tc, tp = _bs_truth(i["S"], i["K"], i["T"])
    # every op but log/exp/erf is exact; those three are <= 2 ulp from NumPy's, and a 1-ulp
    # difference in any of them moves call/put by at most a few ulp OF THE OPERAND SCALE
    # max(S, K) (the result is a difference of two terms that large), hence this bound:
    scale = 16 * EPS32 * np.maximum(i["S"], i["K"])
    for nm, got, want, truth in (("call", call.get(), rc.get(), tc), ("put", put.get(), rp.get(), tp)):
        assert got.dtype == np.float32
        assert np.all(np.abs(got.astype(np.float64) - want) <= scale), nm
        err_got = np.abs(got - truth).mean()
        err_ref = np.abs(want - truth).mean()
        assert err_got <= 1.05 * err_ref + 1e-9, (nm, err_got, err_ref)
        diff = np.abs(got.astype(np.float64) - want)
        used = float(np.max(diff / (EPS32 * np.maximum(i["S"], i["K"]))))          # the bound is 16
        print(f"{nm}: bit-identical to NumPy {np.mean(got == want):.3f}; mean |err| vs f64 "
              f"ours {err_got:.3e} reference {err_ref:.3e}; max |ours - NumPy| = {used:.2f} eps32*max(S,K)")
        assert used <= 4.0, (nm, used)      # observed 1.1 - 2.4 on the B200 (VERDICT r1, weak point 13)
    if n == 2048:
        assert np.all(np.abs(call.get() - GOLDEN["bs_call"]) <= scale)
        assert np.all(np.abs(put.get() - GOLDEN["bs_put"]) <= scale)


def test_black_scholes_separately_forced_equals_coevaluated(gpu):
    i = wl.make_inputs("black_scholes", 5000)
    c1, p1 = wl.black_scholes(gpu, *(gpu.array(i[k]) for k in ("S", "K", "T")))
    a, b = c1.get(), p1.get()                       # two kernels, as a reference user would
    gpu.reset()
    c2, p2 = wl.black_scholes(gpu, *(gpu.array(i[k]) for k in ("S", "K", "T")))
    gpu.evaluate(c2, p2)                            # one two-output kernel
    assert_bits_equal(a, c2.get(), "call")
    assert_bits_equal(b, p2.get(), "put")


# ------------------------------------------------------------------ C3: fused reductions
@pytest.mark.parametrize("n", [1, 7, 2048, (1 << 22) + 5])
def test_l2_dot_norm_fp64_rtol_1e12(gpu, n):
    i = wl.make_inputs("l2", n)
    a, b = gpu.array(i["a"]), gpu.array(i["b"])
    ra, rb = refcpu.leaf(i["a"]), refcpu.leaf(i["b"])
    for nm, got, want in (
            ("l2", wl.l2_distance(gpu, a, b), wl.l2_distance(refcpu, ra, rb)),
            ("dot", wl.dot(gpu, a, b), wl.dot(refcpu, ra, rb).get()),
            ("norm", wl.norm(gpu, a), np.sqrt(wl.dot(refcpu, ra, ra).get()))):
        got = got.get()
        assert got.dtype == np.float64 and got.shape == ()
        scale = max(abs(float(want)), np.sqrt(n) * 1e-3)      # dot of random signs cancels
        assert abs(float(got) - float(want)) <= 1e-12 * scale, (nm, got, want)
    if n == 2048:
        np.testing.assert_allclose(wl.l2_distance(gpu, a, b).get(), GOLDEN["l2"], rtol=1e-12)
        np.testing.assert_allclose(wl.dot(gpu, a, b).get(), GOLDEN["dot"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(wl.norm(gpu, a).get(), GOLDEN["norm"], rtol=1e-12)


def test_reductions_fp32_rtol_1e5_and_exact_cases(gpu):
    rng = np.random.default_rng(23)
    x = rng.uniform(0, 1, (1 << 21) + 3).astype(np.float32)
    X = gpu.array(x)
    np.testing.assert_allclose(np.sum(X).get(), np.sum(x), rtol=1e-5)
    np.testing.assert_allclose(np.mean(X).get(), np.mean(x), rtol=1e-5)
    assert np.sum(X).dtype == np.float32
    assert_bits_equal(np.max(X).get(), np.max(x), "max")
    assert_bits_equal(np.min(X).get(), np.min(x), "min")
    xi = rng.integers(-1000, 1000, 100001)
    assert_bits_equal(np.sum(gpu.array(xi)).get(), np.sum(xi), "int sum")
    assert_bits_equal(np.sum(gpu.array(xi > 0)).get(), np.sum(xi > 0), "bool sum")
    assert_bits_equal(gpu.sum(gpu.full((64,), 7).astype(np.float32)).get(), np.float32(448.0), "ref test_sum")
    np.testing.assert_allclose(np.var(X).get(), np.var(x), rtol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(X).get(), np.linalg.norm(x), rtol=1e-5)


def test_axis_reductions(gpu):
    rng = np.random.default_rng(29)
    m = rng.standard_normal((37, 53, 11))
    M = gpu.array(m)
    for axis in (0, 1, 2, -1, (1, 2), (0, 1), (0, 2), None):
        np.testing.assert_allclose(np.sum(M, axis=axis).get(), np.sum(m, axis=axis), rtol=1e-12, atol=1e-12)
        assert_bits_equal(np.max(M, axis=axis).get(), np.max(m, axis=axis), f"max axis={axis}")
    np.testing.assert_allclose(np.mean(M, axis=1, keepdims=True).get(), np.mean(m, axis=1, keepdims=True), rtol=1e-12)
    big = rng.standard_normal((3, 50000)).astype(np.float32)
    np.testing.assert_allclose(np.sum(gpu.array(big), axis=1).get(), np.sum(big, axis=1), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(np.sum(gpu.array(big), axis=0).get(), np.sum(big, axis=0), rtol=1e-5, atol=1e-5)
    lazy = np.sum((M * 2.0 + 1.0) ** 2, axis=2)       # producer fused into the row reduction
    np.testing.assert_allclose(lazy.get(), np.sum((m * 2.0 + 1.0) ** 2, axis=2), rtol=1e-12)


# ------------------------------------------------------------------ C4: heat stencil, bit-exact
@pytest.mark.parametrize("shape,steps", [((64, 64), 5), ((130, 257), 7), ((3, 3), 2), ((1000, 1031), 3),
                                         ((2100, 4100), 3), ((4096, 8192), 2)])
def test_heat_bit_exact(gpu, shape, steps):
    rng = np.random.default_rng(4)
    u0 = rng.random(shape, dtype=np.float32) if shape != (64, 64) else wl.make_inputs("heat", 64)["u"]
    u = gpu.array(u0.copy())
    wl.heat(gpu, u, steps)
    want = wl.heat(refcpu, refcpu.leaf(u0.copy()), steps).get()
    assert_bits_equal(u.get(), want, f"heat {shape}")
    if shape == (64, 64):
        assert_bits_equal(u.get(), GOLDEN["heat"], "heat golden")


def test_heat_float64_and_general_shifted_assignments(gpu):
    """The stencil kernel for 8-byte cells (3-stage ring: a float64 box is twice as large) and for
    right-hand sides that are not the heat update (found by tools/fuzz_state.py)."""
    rng = np.random.default_rng(44)
    for shape in ((130, 300), (33, 256), (4, 300), (1000, 1031)):
        u0 = rng.random(shape)
        u = gpu.array(u0.copy())
        wl.heat(gpu, u, 3)
        assert_bits_equal(u.get(), wl.heat(refcpu, refcpu.leaf(u0.copy()), 3).get(), f"heat float64 {shape}")
        a, b = u0.copy(), rng.random(shape)
        A, B = gpu.array(a), gpu.array(b)
        a[:, 2:-2] = np.maximum((a[:, 3:-1] + b[:, 2:-2]) * 0.5, 1.5 * b[:, 1:-3])
        A[:, 2:-2] = np.maximum((A[:, 3:-1] + B[:, 2:-2]) * 0.5, 1.5 * B[:, 1:-3])
        assert_bits_equal(A.get(), a, f"shifted max {shape}")
        a[1:-1, :] = np.where(a[2:, :] > a[:-2, :], a[1:-1, :], b[1:-1, :] * 0.5)
        A[1:-1, :] = np.where(A[2:, :] > A[:-2, :], A[1:-1, :], B[1:-1, :] * 0.5)
        assert_bits_equal(A.get(), a, f"shifted where {shape}")


def test_heat_steps_reuse_one_kernel(gpu):
    from delayrepay_b200 import engine
    u = gpu.array(wl.make_inputs("heat", 96)["u"])
    wl.heat(gpu, u, 2)
    before = engine.stats["compiled"] + engine.stats["disk_hits"]
    wl.heat(gpu, u, 10)
    assert engine.stats["compiled"] + engine.stats["disk_hits"] == before


# ------------------------------------------------------------------ C5: n-body
@pytest.mark.parametrize("n", [128, 517])
def test_nbody_parity(gpu, n):
    i = wl.make_inputs("nbody", n)
    got = wl.nbody_acc(gpu, gpu.array(i["pos"]), gpu.array(i["m"])).get()
    want = np.asarray(wl.nbody_acc(refcpu, refcpu.leaf(i["pos"]), refcpu.leaf(i["m"])))
    truth = wl.nbody_acc(np, i["pos"].astype(np.float64), i["m"].astype(np.float64))
    assert got.shape == (n, 3) and got.dtype == np.float32
    # acc = W@pos - pos*rowsum(W) cancels: tolerance is rtol 1e-5 of the terms' magnitude
    w = i["m"][None, :] * ((i["pos"][:, None, :] - i["pos"][None, :, :]) ** 2).sum(-1).__add__(1e-3) ** -1.5
    scale = (w[:, :, None] * np.abs(i["pos"])[None, :, :]).sum(1) + np.abs(i["pos"]) * w.sum(1)[:, None]
    assert np.all(np.abs(got - want) <= 1e-5 * scale)
    assert np.abs(got - truth).max() <= 1.5 * np.abs(want - truth).max() + 1e-5 * scale.max()
    if n == 128:
        assert np.all(np.abs(got - GOLDEN["nbody"]) <= 1e-5 * scale)


# ------------------------------------------------------------------ semantics around the path
def test_setitem_invalidates_memoised_results(gpu):
    """SURVEY.md section 7 'stale results': the reference returns the OLD x + y after x[0] = 100."""
    x, y = gpu.array(np.arange(8.0)), gpu.array(np.ones(8))
    first = (x + y).get()
    x[0] = 100.0
    second = (x + y).get()
    assert first[0] == 1.0 and second[0] == 101.0


def test_views_alias_and_assign(gpu):
    a0 = np.arange(48, dtype=np.float32).reshape(6, 8)
    a = gpu.array(a0.copy())
    row = a[2]
    a[2, :] = -1.0
    assert_bits_equal(row.get(), np.full(8, -1, np.float32), "view sees write")
    a[1:4, ::2] = a[1:4, 1::2] * 2
    b = a0.copy(); b[2, :] = -1; b[1:4, ::2] = b[1:4, 1::2] * 2
    assert_bits_equal(a.get(), b, "strided assign")
    a[:, 1:] = a[:, :-1]              # overlapping shifted self-assignment: temp semantics
    b[:, 1:] = b[:, :-1].copy()
    assert_bits_equal(a.get(), b, "overlap assign")
    assert_bits_equal(a.T.get(), b.T, "transpose")
    assert_bits_equal(a.reshape(8, 6).get(), b.reshape(8, 6), "reshape")
    assert_bits_equal(a[::-1, ::-2].get(), b[::-1, ::-2], "negative strides")


def test_broadcast_where_and_mixed_ops(gpu):
    rng = np.random.default_rng(31)
    m, v = rng.standard_normal((33, 65)), rng.standard_normal(65)
    M, V = gpu.array(m), gpu.array(v)
    assert_bits_equal((M + V).get(), m + v, "row broadcast")
    assert_bits_equal((M * V[None, :] - M[:, :1]).get(), m * v[None, :] - m[:, :1], "col broadcast")
    assert_bits_equal(np.where(M > 0, M, V).get(), np.where(m > 0, m, v), "where")
    assert_bits_equal(np.where(M > 0, 1.0, -1.0).get(), np.where(m > 0, 1.0, -1.0), "where scalars")
    assert_bits_equal(np.clip(M, -0.5, 0.5).get(), np.clip(m, -0.5, 0.5), "clip")
    assert_bits_equal((V[:, None] - V[None, :]).get(), v[:, None] - v[None, :], "outer difference")
    z = gpu.zeros((0, 5))
    assert (z + 1).get().shape == (0, 5)
    assert float(np.sum(z).get()) == 0.0
    s = gpu.array(np.float64(3.0))
    assert float((s * 2).get()[()]) == 6.0


def test_matvec_gemm_and_dot_variants(gpu):
    """reference tests/test.py:114-140 (TestMatrix) and :87-100 (the vacuous dot tests, made real)."""
    rng = np.random.default_rng(37)
    a, b, v = (rng.standard_normal(s).astype(np.float32) for s in ((64, 48), (48, 40), (48,)))
    A, B, V = gpu.array(a), gpu.array(b), gpu.array(v)
    np.testing.assert_allclose((A @ V).get(), a @ v, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose((A @ B).get(), a @ b, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(np.dot(A, V).get(), a @ v, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(A.dot(B).get(), a @ b, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(((A * 2 + 1) @ (B - 1)).get(), (a * 2 + 1) @ (b - 1), rtol=1e-5, atol=1e-4)
    x = gpu.full((64,), 7).astype(np.float32)
    y = gpu.arange(64).astype(np.float32)
    assert float(x.dot(y)) == float(np.full(64, 7, np.float32).dot(np.arange(64, dtype=np.float32)))
    m = gpu.full((64, 64), 7).astype(np.float32)
    assert_bits_equal((m @ m).get(), np.full((64, 64), 7, np.float32) @ np.full((64, 64), 7, np.float32), "test_gemm")


def test_eager_helpers(gpu):
    a0 = np.arange(24, dtype=np.float64).reshape(4, 6)
    a = gpu.array(a0)
    assert_bits_equal(np.roll(a, 2, axis=1).get(), np.roll(a0, 2, axis=1), "roll axis")
    assert_bits_equal(np.roll(a, -5).get(), np.roll(a0, -5), "roll flat")
    assert_bits_equal(np.repeat(a, 3, axis=0).get(), np.repeat(a0, 3, axis=0), "repeat")
    assert_bits_equal(np.tile(a, (2, 3)).get(), np.tile(a0, (2, 3)), "tile")
    assert_bits_equal(np.diag(a).get(), np.diag(a0), "diag")
    assert_bits_equal(np.diagflat(a[0]).get(), np.diagflat(a0[0]), "diagflat")
    assert_bits_equal(np.transpose(a).get(), a0.T, "transpose")
    assert_bits_equal(gpu.linspace(0, 1, 11).get(), np.linspace(0, 1, 11), "linspace")
    assert_bits_equal(gpu.eye(3).get(), np.eye(3), "eye")
    assert gpu.pi == np.pi and gpu.random.rand(4, 3).shape == (4, 3)


def test_map_chunks_streamed_equals_eager(gpu):
    """Chunked H2D / compute / D2H pipeline gives bit-identical results to the eager path."""
    n = (1 << 20) + 12345
    i = wl.make_inputs("black_scholes", n)
    hin = [gpu.pinned_empty(n, np.float32) for _ in range(3)]
    for h, k in zip(hin, ("S", "K", "T")):
        h[:] = i[k]
    hout = [gpu.pinned_empty(n, np.float32) for _ in range(2)]
    gpu.map_chunks(lambda s, k, t: wl.black_scholes(gpu, s, k, t), hin, hout, chunk=1 << 17)
    call, put = wl.black_scholes(gpu, *(gpu.array(i[k]) for k in ("S", "K", "T")))
    gpu.evaluate(call, put)
    assert_bits_equal(hout[0], call.get(), "streamed call")
    assert_bits_equal(hout[1], put.get(), "streamed put")


def test_cumsum(gpu):
    rng = np.random.default_rng(43)
    xi = rng.integers(-100, 100, 1_000_003)
    assert_bits_equal(np.cumsum(gpu.array(xi)).get(), np.cumsum(xi), "int cumsum 1-d (3-phase)")
    xf = rng.standard_normal(300_007)
    np.testing.assert_allclose(np.cumsum(gpu.array(xf)).get(), np.cumsum(xf), rtol=1e-12, atol=1e-9)
    m = rng.standard_normal((37, 53, 11)).astype(np.float32)
    for axis in (0, 1, 2, None):
        np.testing.assert_allclose(np.cumsum(gpu.array(m), axis=axis).get(), np.cumsum(m, axis=axis),
                                   rtol=1e-5, atol=1e-4)
    assert_bits_equal(np.cumsum(gpu.array(np.ones(10, np.int8))).get(), np.cumsum(np.ones(10, np.int8)), "int8 promotes")
    lazy = np.cumsum(gpu.array(xf) * 2.0 + 1.0)
    np.testing.assert_allclose(lazy.get(), np.cumsum(xf * 2.0 + 1.0), rtol=1e-12, atol=1e-9)


def test_device_random_statistics(gpu):
    """Philox on the device: statistical parity only (different generator from NumPy's)."""
    gpu.random.seed(1234)
    u = gpu.random.rand(1 << 20).get()
    assert u.dtype == np.float64 and 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1 / 12) < 1e-3
    g = gpu.random.randn(1 << 20).get()
    assert abs(g.mean()) < 5e-3 and abs(g.std() - 1.0) < 5e-3 and abs((g ** 3).mean()) < 2e-2
    i = gpu.random.randint(3, 11, (1000, 257)).get()
    assert i.shape == (1000, 257) and i.min() == 3 and i.max() == 10
    counts = np.bincount(i.ravel() - 3, minlength=8) / i.size
    assert np.all(np.abs(counts - 0.125) < 5e-3)
    gpu.random.seed(1234)
    assert_bits_equal(gpu.random.rand(1 << 20).get(), u, "seed reproducibility")
    assert not np.array_equal(gpu.random.rand(1 << 20).get(), u)      # the stream advances
    # histogram uniformity (chi-square, 64 bins)
    h = np.histogram(u, bins=64, range=(0, 1))[0]
    chi2 = ((h - u.size / 64) ** 2 / (u.size / 64)).sum()
    assert chi2 < 130, chi2
    assert gpu.random.rand(4, 3).shape == (4, 3) and gpu.random.random((2, 2)).shape == (2, 2)


def test_fft_matches_numpy(gpu):
    rng = np.random.default_rng(47)
    x = rng.standard_normal(4096)
    got = gpu.fft.fft(gpu.array(x)).get()
    want = np.fft.fft(x)
    assert got.dtype == np.complex128
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-9)
    m = rng.standard_normal((7, 1000)).astype(np.float32)
    got = gpu.fft.fft(gpu.array(m)).get()
    assert got.dtype == np.complex64
    np.testing.assert_allclose(got, np.fft.fft(m), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(gpu.fft.fft(gpu.array(m), axis=0).get(), np.fft.fft(m, axis=0), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(gpu.fft.fft(gpu.array(x), n=1024).get(), np.fft.fft(x, n=1024), rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(gpu.fft.fft(gpu.array(x), n=5000).get(), np.fft.fft(x, n=5000), rtol=1e-10, atol=1e-9)
    back = gpu.fft.ifft(gpu.fft.fft(gpu.array(x))).get()
    np.testing.assert_allclose(back.real, x, rtol=1e-10, atol=1e-10)
    lazy = gpu.fft.fft(gpu.array(x) * 2.0)            # forces the lazy operand, like the reference
    np.testing.assert_allclose(lazy.get(), np.fft.fft(x * 2.0), rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("shape", [(256, 192, 320), (1000, 777, 650), (128, 64, 128), (2048, 2048, 2048),
                                   (1500, 300, 8), (4096, 1024, 64), (9, 200, 700), (130, 64, 17)])
def test_dense_matmul_tcgen05(gpu, shape):
    """A genuine dense float32 `@` runs on the tensor cores (tcgen05, 3xTF32 split) and must still
    meet the fp32 bar: |C - C_exact| <= 1e-5 * (|A| @ |B|), the usual matmul tolerance."""
    from delayrepay_b200 import engine
    m, k, n = shape
    rng = np.random.default_rng(53)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((k, n)).astype(np.float32)
    A, B = gpu.array(a), gpu.array(b)
    before = engine.stats["launches"]
    got = (A @ B).get()
    assert any(name[0] == "tcgen05_gemm" for name in engine._kernels), "tensor-core path not taken"
    exact = a.astype(np.float64) @ b.astype(np.float64)
    scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    assert got.dtype == np.float32 and got.shape == (m, n)
    assert np.all(np.abs(got - exact) <= 1e-5 * scale)
    np.testing.assert_allclose(got, a @ b, rtol=1e-5, atol=1e-5 * float(scale.max()))
    if m <= 1000:       # lazy producers are fused into the operand split pre-pass
        got2 = ((A * 2.0 + 1.0) @ (B - 0.5)).get()
        ex2 = (a.astype(np.float64) * 2.0 + 1.0) @ (b.astype(np.float64) - 0.5)
        sc2 = np.abs(a * 2.0 + 1.0).astype(np.float64) @ np.abs(b - 0.5).astype(np.float64)
        assert np.all(np.abs(got2 - ex2) <= 1e-5 * sc2)


# ------------------------------------------------------------------ found by tools/fuzz_diff.py
def test_full_reduction_over_strided_and_broadcast_operands(gpu):
    rng = np.random.default_rng(31)
    a = rng.standard_normal((301, 130))
    b = rng.standard_normal((130, 301))
    c = rng.standard_normal(130)
    A, B, Cc = gpu.array(a), gpu.array(b), gpu.array(c)
    for got, want in ((np.sum(A[:, ::2]), np.sum(a[:, ::2])),
                      (np.sum(A * B.T), np.sum(a * b.T)),
                      (np.sum(A[1:] + Cc), np.sum(a[1:] + c)),
                      (np.prod(A[::-1][:5, :3] * 0.5 + 1.0), np.prod(a[::-1][:5, :3] * 0.5 + 1.0)),
                      (np.mean(abs(A[::3, 1:])), np.mean(abs(a[::3, 1:])))):
        np.testing.assert_allclose(np.asarray(got.get()), want, rtol=1e-12)
    assert np.max(A[:, ::2]).get() == np.max(a[:, ::2])
    assert np.min(A.T[1:]).get() == np.min(a.T[1:])


def test_floor_ceil_trunc_keep_integer_and_boolean_operands(gpu):
    i = np.arange(-5, 6, dtype=np.int32)
    b = i > 0
    for fn in (np.floor, np.ceil, np.trunc):
        for host in (i, i.astype(np.int64), b):
            got = fn(gpu.array(host)).get()
            want = fn(host)
            assert got.dtype == want.dtype and np.array_equal(got, want), (fn.__name__, host.dtype)
    with pytest.raises(TypeError):                      # float16 loops are not supported
        np.sqrt(gpu.array(b))


def test_where_with_python_scalars_is_weakly_typed(gpu):
    x = np.linspace(-1, 1, 37, dtype=np.float32)
    y = np.linspace(2, 3, 37, dtype=np.float64)
    X, Y = gpu.array(x), gpu.array(y)
    for got, want in ((np.where(X > 0, X, 0.1), np.where(x > 0, x, 0.1)),
                      (np.where(X > 0, 2, X), np.where(x > 0, 2, x)),
                      (np.where(X > 0, X, np.float64(0.1)), np.where(x > 0, x, np.float64(0.1))),
                      (np.where(X > 0, Y, 1), np.where(x > 0, y, 1)),
                      (np.where(X > 0, X, 0.1) - Y, np.where(x > 0, x, 0.1) - y)):
        got = got.get()
        assert got.dtype == want.dtype
        assert_bits_equal(got, want, "where")


def test_reductions_of_empty_arrays_follow_numpy(gpu):
    e = gpu.array(np.zeros(0))
    assert np.sum(e).get() == 0.0 and np.prod(e).get() == 1.0
    assert np.isnan(np.mean(e * 2.0).get())
    with pytest.raises(ValueError):
        np.max(e).get()
    z = gpu.array(np.zeros((0, 5), dtype=np.float32))
    assert np.array_equal(np.sum(z, axis=0).get(), np.zeros(5, np.float32))
    assert np.array_equal(np.prod(z, axis=0).get(), np.ones(5, np.float32))
    assert np.sum(z, axis=1).get().shape == (0,)


# ------------------------------------------------------------------ axis kernels, second generation
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_axis_reductions_vector_split_and_transposed_paths(gpu, dt):
    """rows / cols kernels with 128-bit loads, a split reduced axis (partials folded in fixed
    order), and the transposed dispatch (`v @ X`, reductions of X.T); integer results exact,
    float sums within the reduction bar, max/min exact."""
    rng = np.random.default_rng(71)
    scale = [1.0]       # the bar is relative to the sum of magnitudes (NumPy's own float32 row-by-
    #                     row accumulation is ~1e-5 of that away from the exact sum)

    def check(got, want, what):
        tol = dict(rtol=1e-12, atol=1e-12 * scale[0]) if dt == np.float64 else dict(rtol=1e-5, atol=1e-5 * scale[0])
        got = got.get()
        assert got.shape == want.shape and got.dtype == want.dtype, (what, got.dtype, want.dtype)
        if dt == np.int32 or "max" in what or "min" in what:
            assert np.array_equal(got, want), what
        else:
            np.testing.assert_allclose(got, want, err_msg=what, **tol)

    for shape in ((2048, 512), (4099, 260), (515, 1028), (64, 5000), (7, 100003), (30000, 12)):
        x = (rng.standard_normal(shape) * 20).astype(dt)
        X = gpu.array(x)
        scale[0] = float(np.abs(x).max()) ** 2 * max(shape)
        v0 = (rng.standard_normal(shape[0]) * 3).astype(dt)
        v1 = (rng.standard_normal(shape[1]) * 3).astype(dt)
        check(np.sum(X, axis=0), np.sum(x, axis=0), f"sum0 {shape}")
        check(np.sum(X, axis=1), np.sum(x, axis=1), f"sum1 {shape}")
        check(np.max(X, axis=0), np.max(x, axis=0), f"max0 {shape}")
        check(np.min(X, axis=1), np.min(x, axis=1), f"min1 {shape}")
        check(np.sum(X.T, axis=1), np.sum(x.T, axis=1), f"sumT1 {shape}")
        check(np.max(X.T, axis=0), np.max(x.T, axis=0), f"maxT0 {shape}")
        check(X @ gpu.array(v1), x @ v1, f"X@v {shape}")
        check(gpu.array(v0) @ X, v0 @ x, f"v@X {shape}")
        check(np.sum(X[1:, 1:], axis=0), np.sum(x[1:, 1:], axis=0), f"sum0 offset {shape}")
        check(np.sum(X[:, ::2], axis=1), np.sum(x[:, ::2], axis=1), f"sum1 strided {shape}")
        check(np.sum(X * gpu.array(v1)[None, :], axis=0), np.sum(x * v1[None, :], axis=0), f"sum0 bcast {shape}")
        check(np.sum(X * gpu.array(v0)[:, None], axis=1), np.sum(x * v0[:, None], axis=1), f"sum1 bcast {shape}")
        if dt != np.int32:
            check(np.mean(X * X, axis=0), np.mean(x * x, axis=0), f"mean0 {shape}")
            np.testing.assert_allclose(np.var(X, axis=0).get(), np.var(x, axis=0), rtol=1e-4 if dt == np.float32 else 1e-10)
    c = (rng.standard_normal((6, 4100, 36)) * 5).astype(dt)
    scale[0] = float(np.abs(c).max()) * 4100
    check(np.sum(gpu.array(c), axis=1), np.sum(c, axis=1), "middle axis")
    check(np.max(gpu.array(c), axis=1), np.max(c, axis=1), "middle axis max")


def test_argmax_argmin_pair_reduction(gpu):
    rng = np.random.default_rng(72)
    for dt in (np.float32, np.float64, np.int32, np.int64, np.uint8, np.bool_):
        for shape in ((1,), (100003,), (3_000_001,), (513, 1030), (6, 4100, 36), (70000, 3)):
            x = (rng.standard_normal(shape) * 50).astype(dt)       # small dtypes: many ties
            X = gpu.array(x)
            for ax in (None,) + tuple(range(len(shape))):
                for fn in (np.argmax, np.argmin):
                    got = fn(X, axis=ax).get()
                    want = fn(x, axis=ax)
                    assert got.shape == np.shape(want) and np.array_equal(got, want), (dt, shape, ax, fn.__name__)
    x = rng.standard_normal(5000)
    x[[77, 4000]] = np.nan                                         # first nan wins, like NumPy
    assert int(np.argmax(gpu.array(x))) == 77 and int(np.argmin(gpu.array(x))) == 77
    m = rng.standard_normal((300, 40)); m[200, 7] = np.nan; m[100, 7] = np.nan
    assert np.array_equal(np.argmax(gpu.array(m), axis=0).get(), np.argmax(m, axis=0))
    assert np.array_equal(np.argmin(gpu.array(m), axis=1, keepdims=True).get(), np.argmin(m, axis=1, keepdims=True))
    assert int(np.argmax(gpu.array(x) * 2.0 + 1.0)) == 77        # lazy producer is forced first
    with pytest.raises(ValueError):
        np.argmax(gpu.array(np.zeros(0)))
    # the 128-bit path of the row kernel: constant data (first index wins), infinities, nans in
    # every vector lane and thread position, views that start off a 16-byte boundary
    for dt in (np.float32, np.float64):
        for n in (4096, 1_000_000, 3_000_003):
            for fill in (-np.inf, np.inf, 0.0, np.nan):
                x = np.full(n, fill, dt)
                for fn in (np.argmax, np.argmin):
                    assert int(fn(gpu.array(x))) == int(fn(x)) == 0, (dt, n, fill, fn.__name__)
            x = rng.standard_normal(n).astype(dt)
            for pos in (0, 1, 2, 3, 1023, 1024, 1027, n // 2 + 1, n - 1):
                y = x.copy(); y[pos] = np.nan; y[min(n - 1, pos + 5000)] = np.nan
                assert int(np.argmax(gpu.array(y))) == pos and int(np.argmin(gpu.array(y))) == pos, (dt, n, pos)
                y = x.copy(); y[pos] = np.inf; y[(pos * 7) % n] = -np.inf
                assert int(np.argmax(gpu.array(y))) == int(np.argmax(y)) and int(np.argmin(gpu.array(y))) == int(np.argmin(y))
            X = gpu.array(x)
            for off in (1, 2, 3, 5):
                assert int(np.argmax(X[off:])) == int(np.argmax(x[off:])), "misaligned view: scalar path"
                assert int(np.argmin(X[off:n - 3])) == int(np.argmin(x[off:n - 3]))
    ties = rng.integers(0, 3, 5_000_000).astype(np.int8)
    assert int(np.argmax(gpu.array(ties))) == int(np.argmax(ties)) and int(np.argmin(gpu.array(ties))) == int(np.argmin(ties))


def test_cumsum_row_scan_and_chunked_axis_scan(gpu):
    rng = np.random.default_rng(73)
    for shape in ((300, 5000), (64, 64), (1000, 2049), (5000, 300), (3, 70000, 5), (9, 40, 1000)):
        xi = rng.integers(-100, 100, shape)
        xf = rng.standard_normal(shape)
        for ax in range(len(shape)):
            assert_bits_equal(np.cumsum(gpu.array(xi), axis=ax).get(), np.cumsum(xi, axis=ax), f"int cumsum {shape} {ax}")
            np.testing.assert_allclose(np.cumsum(gpu.array(xf), axis=ax).get(), np.cumsum(xf, axis=ax),
                                       rtol=1e-11, atol=1e-9, err_msg=f"{shape} {ax}")
        x32 = xf.astype(np.float32)
        np.testing.assert_allclose(np.cumsum(gpu.array(x32), axis=-1).get(), np.cumsum(x32, axis=-1), rtol=1e-4, atol=1e-2)
    b = rng.integers(0, 2, (200, 3000)).astype(bool)
    assert_bits_equal(np.cumsum(gpu.array(b), axis=1).get(), np.cumsum(b, axis=1), "bool cumsum rows")


def test_internal_casts_never_convert_the_users_leaf(gpu):
    """`leaf.astype` is in place (reference delayarray.py:401-408), so promotions inside
    reductions, contractions and dtype= arguments must go through a cast NODE: an int32 array is
    still int32 after it has been summed (found by the second-generation axis tests)."""
    xi = np.arange(12, dtype=np.int32).reshape(3, 4)
    X = gpu.array(xi)
    v = gpu.array(np.ones(4, dtype=np.float64))
    for use in (lambda: np.sum(X).get(), lambda: np.sum(X, axis=0).get(), lambda: np.mean(X).get(),
                lambda: np.sum(X, dtype=np.float32).get(), lambda: np.cumsum(X, dtype=np.float64).get(),
                lambda: (X @ v).get(), lambda: (X @ gpu.array(np.ones((4, 2), np.float32))).get(),
                lambda: np.var(X, dtype=np.float64).get(), lambda: np.add(X, 1, dtype=np.float64).get()):
        use()
        assert X.dtype == np.int32 and X.get().dtype == np.int32
    b = gpu.array(xi > 5)
    assert int(np.sum(b)) == 6 and b.dtype == np.bool_
    assert np.max(X, axis=0).get().dtype == np.int32
    f = gpu.array(np.ones((130, 64), np.float64))
    (f @ gpu.array(np.ones((64, 130), np.float32))).get()
    assert f.dtype == np.float64
    assert X.astype(np.float32) is X and X.dtype == np.float32          # the user-facing call IS in place


def test_reduction_plan_cache_replays_are_exact(gpu):
    """Prepared launches for fused full reductions (engine._reduce_plan_key): replays see the
    current values, other layouts / dtypes / scalar classes fall back to the planner."""
    from delayrepay_b200 import engine
    rng = np.random.default_rng(91)
    a, b = rng.standard_normal(100003), rng.standard_normal(100003)
    A, B = gpu.array(a), gpu.array(b)
    hits = engine.stats.get("plan_hits", 0)
    for k in range(4):
        c = float(k) + 0.5
        np.testing.assert_allclose(float(np.sqrt(np.sum((A - B * c) ** 2))), np.sqrt(np.sum((a - b * c) ** 2)), rtol=1e-12)
        np.testing.assert_allclose(float(np.dot(A, B)), np.dot(a, b), rtol=1e-12, atol=1e-9)
        assert float(np.max(A)) == a.max() and float(np.min(A * c)) == (a * c).min()
        np.testing.assert_allclose(float(np.mean(A * B)), np.mean(a * b), rtol=1e-11, atol=1e-12)
        a[k] = 100.0 + k
        A[k] = 100.0 + k                                  # replays must read the new data
    assert engine.stats.get("plan_hits", 0) >= hits + 12
    np.testing.assert_allclose(float(np.sum(A[::2])), a[::2].sum(), rtol=1e-12)          # other layout
    np.testing.assert_allclose(float(np.sum(A[1:])), a[1:].sum(), rtol=1e-12)            # misaligned
    a32 = a.astype(np.float32)
    np.testing.assert_allclose(float(np.sum(gpu.array(a32))), a32.astype(np.float64).sum(), rtol=1e-6)
    assert int(np.sum(gpu.array(np.arange(10)))) == 45 and int(np.sum(gpu.array(np.arange(12)))) == 66


def test_signed_integer_overflow_wraps_like_numpy(gpu):
    """Signed overflow is undefined in C++ and NVRTC uses that (it widened `(double)(x*x*x)`);
    NumPy wraps.  Found by tools/fuzz_diff.py seed 4414."""
    rng = np.random.default_rng(93)
    for dt in (np.int8, np.int16, np.int32, np.int64):
        info = np.iinfo(dt)
        x = rng.integers(info.min, info.max, 5003, dtype=dt, endpoint=True)
        y = rng.integers(info.min, info.max, 5003, dtype=dt, endpoint=True)
        X, Y = gpu.array(x), gpu.array(y)
        with np.errstate(all="ignore"):
            cases = [(X * Y, x * y), (X + Y, x + y), (X - Y, x - y), (-X, -x), (X ** 3, (x * x) * x),
                     (X * Y + X, x * y + x), (X << 3, x << 3),
                     ((X * Y).astype(np.float64), (x * y).astype(np.float64))]
            if dt in (np.int32, np.int64):       # (float functions of int8 / int16 are float16 / float32 loops)
                cases.append((np.arcsinh((X * X) * X), np.arcsinh((x * x) * x)))
            for got, want in cases:
                got = got.get()
                assert got.dtype == want.dtype
                if want.dtype.kind == "f":
                    np.testing.assert_allclose(got, want, rtol=1e-15)
                else:
                    assert np.array_equal(got, want), dt
            if dt in (np.int32, np.int64):
                big = rng.integers(info.max // 4, info.max, 1003, dtype=dt)
                assert int(np.sum(gpu.array(big))) == int(np.sum(big))            # int64 accumulator wraps too
                assert np.prod(gpu.array(big)).get() == np.prod(big)


def test_tile_family_transposed_operands_bit_exact(gpu):
    """`X.T + X` and friends go through the shared-memory tile kernel (codegen.gen_tile): same
    arithmetic as the strided kernel, so everything is bit-exact against NumPy."""
    from delayrepay_b200 import engine
    rng = np.random.default_rng(101)
    for dt in (np.float32, np.float64, np.int32, np.int64):
        for r, c in ((64, 64), (257, 130), (33, 1000), (1000, 33), (640, 448)):
            a = (rng.standard_normal((r, c)) * 100).astype(dt)
            b = (rng.standard_normal((c, r)) * 100).astype(dt)
            v = (rng.standard_normal(c) * 100).astype(dt)
            A, B, V = gpu.array(a), gpu.array(b), gpu.array(v)
            got = (B.T + A).get()
            assert engine.last_kernel_name().startswith("dr_tile_"), engine.last_kernel_name()
            assert_bits_equal(got, b.T + a, f"X.T + X {dt} {r}x{c}")
            assert_bits_equal((B.T * 3 - A * V).get(), b.T * 3 - a * v, f"row vector beside a transposed operand {dt}")
            assert_bits_equal(((B.T > A) & (A > 0)).get(), (b.T > a) & (a > 0), f"bool output {dt}")
            assert_bits_equal((B.T + B.T * B.T).get(), b.T + b.T * b.T, f"only transposed operands {dt}")
            assert_bits_equal(B.T.copy().get(), b.T, "plain transpose")
    # a column block of a wider matrix, transposed; mixed widths (float32 tile beside float64 rows)
    w = rng.standard_normal((300, 500)).astype(np.float32)
    a = rng.standard_normal((200, 300))
    W, A = gpu.array(w), gpu.array(a)
    assert_bits_equal((W[:, 100:300].T + A).get(), w[:, 100:300].T + a, "transposed column block, mixed dtypes")
    assert engine.last_kernel_name().startswith("dr_tile_")
    # two roots over the same transposed operand stay one kernel
    s, d = gpu.evaluate(W.T + 1.0, W.T * 2.0)
    assert_bits_equal(s.get(), w.T + np.float32(1.0), "co-evaluated root 0")
    assert_bits_equal(d.get(), w.T * np.float32(2.0), "co-evaluated root 1")
    # transcendental body and the plan cache (second call replays the prepared launch)
    for _ in range(2):
        got = np.exp(W.T * 0.01).get()
        assert_ulp(got, np.exp((w.T * np.float32(0.01)).astype(np.float64)).astype(np.float32), 2,
                   "exp over a transposed operand")


def test_generation2_scans(gpu):
    """One-pass chained 1-d scan and the vectorised row scan (extras._SCAN2_SRC): integers exact,
    floats within the reduction tolerance, and bit-reproducible from run to run."""
    from delayrepay_b200 import engine
    rng = np.random.default_rng(202)
    for n in ((1 << 20), (1 << 20) + 3, 5_000_001, (1 << 24) + 8191):
        xi = rng.integers(-1000, 1000, n).astype(np.int32)
        assert_bits_equal(np.cumsum(gpu.array(xi)).get(), np.cumsum(xi), f"int32 chained scan {n}")
        assert engine.last_kernel_name().endswith("_chain"), engine.last_kernel_name()
        xl = rng.integers(-10 ** 12, 10 ** 12, n)
        assert_bits_equal(np.cumsum(gpu.array(xl)).get(), np.cumsum(xl), f"int64 chained scan {n}")
        xb = rng.integers(0, 2, n).astype(bool)
        assert_bits_equal(np.cumsum(gpu.array(xb)).get(), np.cumsum(xb), f"bool chained scan {n}")
        xd = rng.standard_normal(n)
        first = np.cumsum(gpu.array(xd)).get()
        np.testing.assert_allclose(first, np.cumsum(xd), rtol=1e-11, atol=1e-8)
        assert_bits_equal(np.cumsum(gpu.array(xd)).get(), first, "float64 scan is reproducible")
        xf = rng.random(n).astype(np.float32)
        first = np.cumsum(gpu.array(xf)).get()
        np.testing.assert_allclose(first, np.cumsum(xf.astype(np.float64)), rtol=1e-5)
        assert_bits_equal(np.cumsum(gpu.array(xf)).get(), first, "float32 scan is reproducible")
    x8 = rng.integers(-100, 100, 3_000_000).astype(np.int8)
    assert_bits_equal(np.cumsum(gpu.array(x8)).get(), np.cumsum(x8), "int8 chained scan")
    xu = rng.integers(0, 60000, 3_000_000).astype(np.uint16)
    assert_bits_equal(np.cumsum(gpu.array(xu)).get(), np.cumsum(xu), "uint16 chained scan")
    # a view that starts 4 bytes into the allocation is not 16-byte aligned: first-generation path
    xi = rng.integers(-1000, 1000, (1 << 21) + 1).astype(np.int32)
    assert_bits_equal(np.cumsum(gpu.array(xi)[1:]).get(), np.cumsum(xi[1:]), "unaligned 1-d scan")
    for shape in ((64, 1024), (300, 4100), (1000, 20480), (70, 65536 + 64)):
        mi = rng.integers(-100, 100, shape).astype(np.int32)
        assert_bits_equal(np.cumsum(gpu.array(mi), axis=1).get(), np.cumsum(mi, axis=1), f"row scan int32 {shape}")
        assert engine.last_kernel_name().endswith("_rowscan2"), engine.last_kernel_name()
        md = rng.standard_normal(shape)
        np.testing.assert_allclose(np.cumsum(gpu.array(md), axis=1).get(), np.cumsum(md, axis=1), rtol=1e-11, atol=1e-9)
        mf = rng.random(shape).astype(np.float32)
        np.testing.assert_allclose(np.cumsum(gpu.array(mf), axis=1).get(), np.cumsum(mf.astype(np.float64), axis=1),
                                   rtol=1e-5)
        mb = rng.integers(0, 2, shape).astype(bool)
        assert_bits_equal(np.cumsum(gpu.array(mb), axis=1).get(), np.cumsum(mb, axis=1), f"row scan bool {shape}")


def test_signed_zeros_of_maximum_minimum_clip_follow_numpy(gpu):
    """np.maximum / np.minimum return their SECOND operand when the two compare equal and np.clip
    depends on the kind of bound: only visible for -0.0 against +0.0 (found by fuzz seeds 7141,
    7884 and 8132).  np.fmax / np.fmin are not pinned: NumPy's vector body returns the second
    operand and its scalar tail the first, so the sign depends on the element's position."""
    for dt in (np.float32, np.float64):
        a = np.array([-0.0, 0.0, -0.0, 0.0, np.nan, 1.0, -0.0, 2.0] * 33, dt)
        b = np.array([0.0, -0.0, -0.0, 0.0, 0.0, np.nan, 1.0, 2.0] * 33, dt)
        A, B = gpu.array(a), gpu.array(b)
        for name in ("maximum", "minimum", "fmax", "fmin"):
            fn = getattr(np, name)
            for x, y, X, Y in ((a, b, A, B), (b, a, B, A)):
                got, want = fn(X, Y).get(), fn(x, y)
                assert np.array_equal(got, want, equal_nan=True), name
                if name[0] == "f":
                    continue        # NumPy's own fmax / fmin differ between their SIMD body and scalar tail
                ok = ~np.isnan(want)
                assert np.array_equal(np.signbit(got[ok]), np.signbit(want[ok])), (name, np.dtype(dt).name)
            if name[0] != "f":
                got, want = fn(A, dt(0.0)).get(), fn(a, dt(0.0))
                assert np.array_equal(np.signbit(got[~np.isnan(want)]), np.signbit(want[~np.isnan(want)])), (name, "scalar")
        one = np.ones_like(a)
        ONE = gpu.array(one)
        for lo, hi, LO, HI in ((dt(0.0), dt(1.0), dt(0.0), dt(1.0)), (dt(-0.0), dt(0.0), dt(-0.0), dt(0.0)),
                               (b, one, B, ONE), (-one, b, -ONE, B),
                               (b, None, B, None), (None, b, None, B), (None, dt(-0.0), None, dt(-0.0)),
                               (None, dt(0.0), None, dt(0.0)), (dt(-0.0), None, dt(-0.0), None),
                               (dt(0.0), None, dt(0.0), None), (dt(0.0), one, dt(0.0), ONE), (b, dt(1.0), B, dt(1.0))):
            got, want = np.clip(A, LO, HI).get(), np.clip(a, lo, hi)
            assert np.array_equal(got, want, equal_nan=True)
            ok = ~np.isnan(want)
            assert np.array_equal(np.signbit(got[ok]), np.signbit(want[ok])), ("clip", np.dtype(dt).name, type(lo), type(hi))


def test_row_gather_in_128_bit_words(gpu):
    """X[rows] with rows of a multiple of 16 bytes takes the vector gather (extras._TAKE16_SRC)."""
    from delayrepay_b200 import engine
    rng = np.random.default_rng(303)
    for dt, shape in ((np.float32, (300, 64)), (np.float64, (50, 12)), (np.int8, (40, 64)), (np.float32, (7, 5, 4, 8)),
                      (np.int64, (1000, 2, 8))):
        h = (rng.standard_normal(shape) * 100).astype(dt)
        d = gpu.array(h)
        for idx in (rng.integers(-shape[0], shape[0], 257), rng.integers(0, shape[0], (3, 40)).astype(np.int32),
                    np.array([shape[0] - 1, 0, -1, -shape[0]])):
            got = d[gpu.array(idx)].get()
            assert engine.last_kernel_name().startswith("dr_take16_"), engine.last_kernel_name()
            assert got.shape == h[idx].shape and np.array_equal(got, h[idx]), (dt, shape)
        assert np.array_equal(d[1:][gpu.array(np.array([0, 5, 2]))].get(), h[1:][[0, 5, 2]])
    odd = gpu.array(np.arange(60, dtype=np.float32).reshape(10, 6))          # 24-byte rows: scalar gather
    assert np.array_equal(odd[gpu.array(np.array([3, 3, 0]))].get(), np.arange(60, dtype=np.float32).reshape(10, 6)[[3, 3, 0]])
    with pytest.raises(IndexError):
        gpu.array(np.zeros((10, 64), np.float32))[gpu.array(np.array([10]))]


def test_integer_min_max_of_negated_values(gpu):
    """ptxas 12.9 loses the negation of one operand when it fuses min(min(p, -a), -b) into a
    three-input VIMNMX3 (int32 / int16; found by fuzz seed 60525: np.min(-x) returned min(x)).
    Integer negation is emitted as a subtraction from a zero the assembler cannot fold."""
    rng = np.random.default_rng(60525)
    for dt in (np.int32, np.int64, np.int16, np.int8, np.uint32, np.uint16):
        for n in (672, 100_003):
            h = rng.integers(0 if np.dtype(dt).kind == "u" else -50, 50, n).astype(dt)
            d = gpu.array(h)
            for red in (np.min, np.max):
                assert int(red(np.negative(d)).get()) == int(red(np.negative(h))), (np.dtype(dt).name, n, red.__name__)
                assert int(red(-d + 3).get()) == int(red(-h + dt(3))), (np.dtype(dt).name, n, red.__name__)
            m = h[:672].reshape(21, 32)
            assert np.array_equal(np.min(np.negative(gpu.array(m)), axis=1).get(), np.min(np.negative(m), axis=1))
            assert np.array_equal(np.max(np.negative(gpu.array(m)), axis=0).get(), np.max(np.negative(m), axis=0))
            a, b, c = d, gpu.array(np.roll(h, 1)), gpu.array(np.roll(h, 2))
            for fn in (np.minimum, np.maximum):
                got = fn(fn(np.negative(a), np.negative(b)), np.negative(c)).get()
                want = fn(fn(np.negative(h), np.negative(np.roll(h, 1))), np.negative(np.roll(h, 2)))
                assert_bits_equal(got, want, f"{fn.__name__} of three negated {np.dtype(dt).name}")
            assert_bits_equal(np.negative(d).get(), np.negative(h), "plain negation")
            lo = 1 if np.dtype(dt).kind == "u" else -7
            assert_bits_equal(np.clip(np.negative(d), lo, 9).get(), np.clip(np.negative(h), lo, 9), "clip of a negation")
