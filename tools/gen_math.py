"""Derive and validate the polynomial coefficients of the fast float32 transcendentals in
delayrepay_b200/csrc/prelude.cuh (evaluated in double, rounded once to float).

Near-minimax fits by Chebyshev interpolation in float64 with long-double targets; accuracy is
checked here by emulating the device evaluation order in NumPy float64 against long-double
truth, in float32 ulps.  Run:  python tools/gen_math.py
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

LD = np.longdouble


def cheb_fit(fn, lo, hi, deg):
    """Monomial coefficients (in x) of the degree-`deg` Chebyshev interpolant of fn on [lo,hi]."""
    k = np.arange(deg + 1)
    nodes = np.cos(np.pi * (k + 0.5) / (deg + 1))
    x = (LD(0.5) * (hi - lo)) * nodes.astype(LD) + LD(0.5) * (hi + lo)
    y = fn(x).astype(np.float64)
    c = C.chebfit(nodes, y, deg)                      # exact interpolation at the nodes
    mono_u = C.cheb2poly(c)                           # in u = (2x - (hi+lo)) / (hi-lo)
    a, b = 2.0 / (hi - lo), -(hi + lo) / (hi - lo)
    out = np.zeros(deg + 1)
    lin = np.array([b, a])
    pw = np.array([1.0])
    for ck in mono_u:
        out[:len(pw)] += ck * pw
        pw = P.polymul(pw, lin)
    return out


def horner(c, x):
    r = np.full_like(x, c[-1])
    for ck in c[-2::-1]:
        r = r * x + ck
    return r


def ulp32(got32, truth_ld):
    g = got32.astype(LD)
    u = np.spacing(np.abs(truth_ld.astype(np.float32))).astype(LD)
    return np.abs(g - truth_ld) / u


def fmt(c):
    return ", ".join(f"{v:.17e}" for v in c)


# ---------------------------------------------------------------- exp: 2^k * (1 + r*P(r))
LN2 = float(np.log(LD(2)))
def expm1_over_r(r):
    r = np.where(r == 0, LD(1e-30), r)
    return np.expm1(r) / r
EXP_DEG = 6
exp_c = cheb_fit(expm1_over_r, -LN2 / 2 * 1.0001, LN2 / 2 * 1.0001, EXP_DEG)

def exp_model(x32):
    x = x32.astype(np.float64)
    k = np.rint(x * (1 / LN2))
    r = x - k * LN2
    pm1 = r * horner(exp_c, r)
    return ((1.0 + pm1) * np.exp2(k)).astype(np.float32)


# ---------------------------------------------------------------- log: e*ln2 + f*L(f)
def log1p_over_f(f):
    f = np.where(f == 0, LD(1e-30), f)
    return np.log1p(f) / f
LOG_DEG = 13
log_c = cheb_fit(log1p_over_f, np.sqrt(0.5) - 1 - 1e-4, np.sqrt(2) - 1 + 1e-4, LOG_DEG)

def log_model(x32):
    m, e = np.frexp(x32.astype(np.float64))          # m in [0.5, 1)
    adj = m < np.sqrt(0.5)
    m = np.where(adj, 2 * m, m)
    e = e - adj
    f = m - 1.0
    return (e * LN2 + f * horner(log_c, f)).astype(np.float32)


# ---------------------------------------------------------------- erf: 1 - exp(-a*Q(a))
def erf_ld(a):
    from scipy.special import erf, erfc
    return erf(a.astype(np.float64)).astype(LD)       # double truth is enough for f32 ulps

def q_target(a):
    from scipy.special import erfc
    a64 = a.astype(np.float64)
    a64 = np.where(a64 == 0, 1e-30, a64)
    # -log(erfc(a))/a, computed stably for small a via log1p(-erf)
    from scipy.special import erf
    small = a64 < 0.5
    t = np.where(small, -np.log1p(-erf(a64)), -np.log(erfc(a64)))
    return (t / a64).astype(LD)
ERF_DEG = 11
ERF_HI = 3.95
erf_c = cheb_fit(q_target, 0.0, ERF_HI, ERF_DEG)

def erf_model(x32):
    x = x32.astype(np.float64)
    a = np.minimum(np.abs(x), ERF_HI)
    t = a * horner(erf_c, a)
    u = -t
    k = np.rint(u * (1 / LN2))
    r = u - k * LN2
    pm1 = r * horner(exp_c, r)
    s = np.exp2(k)
    res = (1.0 - s) - s * pm1
    return np.copysign(res, x).astype(np.float32)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.uniform(-100, 88, 1 << 22).astype(np.float32)
    u = ulp32(exp_model(x), np.exp(x.astype(LD)))
    print(f"exp  deg {EXP_DEG}: max {u.max():.4f} ulp   [{fmt(exp_c)}]")
    x = np.concatenate([rng.uniform(0, 4, 1 << 21), np.exp(rng.uniform(-80, 80, 1 << 21)),
                        rng.uniform(0.9, 1.1, 1 << 20)]).astype(np.float32)
    x = x[x > 0]
    u = ulp32(log_model(x), np.log(x.astype(LD)))
    print(f"log  deg {LOG_DEG}: max {u.max():.4f} ulp   [{fmt(log_c)}]")
    x = np.concatenate([rng.uniform(-4.5, 4.5, 1 << 22), rng.uniform(-1e-3, 1e-3, 1 << 18),
                        rng.uniform(-0.1, 0.1, 1 << 20)]).astype(np.float32)
    from scipy.special import erf
    u = ulp32(erf_model(x), erf(x.astype(np.float64)).astype(LD))
    print(f"erf  deg {ERF_DEG}: max {u.max():.4f} ulp   [{fmt(erf_c)}]")


def sweep():
    import sys
    global log_c, erf_c, exp_c
    rng = np.random.default_rng(1)
    from scipy.special import erf
    xl = np.concatenate([rng.uniform(0, 4, 1 << 20), np.exp(rng.uniform(-80, 80, 1 << 20)),
                         rng.uniform(0.9, 1.1, 1 << 20)]).astype(np.float32)
    xl = xl[xl > 0]
    for d in (9, 10, 11, 12):
        log_c = cheb_fit(log1p_over_f, np.sqrt(0.5) - 1 - 1e-4, np.sqrt(2) - 1 + 1e-4, d)
        print("log deg", d, ulp32(log_model(xl), np.log(xl.astype(LD))).max())
    xe = np.concatenate([rng.uniform(-4.5, 4.5, 1 << 21), rng.uniform(-0.1, 0.1, 1 << 19)]).astype(np.float32)
    te = erf(xe.astype(np.float64)).astype(LD)
    for d in (10, 12, 13, 14, 16):
        erf_c = cheb_fit(q_target, 0.0, ERF_HI, d)
        u = ulp32(erf_model(xe), te)
        i = u.argmax()
        print("erf deg", d, u.max(), "at", xe[i])
    for d in (5,):
        exp_c = cheb_fit(expm1_over_r, -LN2 / 2 * 1.0001, LN2 / 2 * 1.0001, d)
        x = rng.uniform(-100, 88, 1 << 21).astype(np.float32)
        print("exp deg", d, ulp32(exp_model(x), np.exp(x.astype(LD))).max())
