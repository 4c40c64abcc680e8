"""Second-generation float32 erf / exp / log for the lockstep kernels: designed for the B200's
pipe balance (FFMA2 has the FMA pipe's flop rate of FFMA, ALU pipe is half rate, a divergent
LDS.64 costs ~5 SMSP-cycles: tools/microbench2.py), so the aim is FEWER float operations and
FEWER integer operations per element, not only fewer issue slots.

Every operation below is one float32 device instruction, emulated with NumPy float32 arithmetic
(fma = float64 product-sum rounded once; the double rounding affects ~2^-29 of the cases).
Errors are float32 ulps against long-double / mpmath truth.

  python tools/gen_math_v2.py            # accuracy report
  python tools/gen_math_v2.py --emit     # rewrite delayrepay_b200/csrc/math_tables.cuh

erf (dr_erf4_gal)   accurate-table method (Gal): per interval a centre c_j chosen among the
                    floats near the midpoint such that erf(c_j) is a float32 to within 2^-10 ulp,
                    so  erf(a) = C0_j + d (C1_j + d (C2_j + d (C3_j + d C4_j))),  d = a - c_j exact.
                    16 intervals per binade from 2^-12 to 4, row 0 = [0, 2^-12) with c = 0.
                    Row = (c, C0 | C1, C2 | C3, C4): three LDS.64, one FADD, four FFMA.
"""
import os
import sys

import numpy as np
import mpmath as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_math import cheb_fit, LD, fmt          # noqa: E402

f32 = np.float32
mp.mp.prec = 120


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def ulp_err(got, truth):
    u = np.spacing(np.abs(truth.astype(f32))).astype(LD)
    return np.abs(got.astype(LD) - truth) / u


def mp_vec(fn, x):
    """fn evaluated by mpmath at the float32/float64 points x, returned as long double."""
    return np.array([LD(mp.nstr(fn(mp.mpf(float(v))), 25)) for v in x], dtype=LD)


# ----------------------------------------------------------------------------------------- erf
ERF_LOW_EXP = -12                     # rows start at 2^-12 (225 rows: 16 bank-private copies = 84 KiB)
ERF_PER_BINADE = 16
ERF_ROWS = (2 - ERF_LOW_EXP) * ERF_PER_BINADE + 1       # + row 0
ERF_BASE = ((127 + ERF_LOW_EXP) << 4) - 1               # (bits >> 19) - ERF_BASE = row, clamped at 0
ERF_AMAX = f32(3.9999998)


def _erf_ld(x):
    from scipy.special import erf
    return erf(np.asarray(x, np.float64)).astype(LD)


def erf_row(j, search=6000):
    """(c, C0, C1..C4) of row j >= 1."""
    b, m = divmod(j - 1, ERF_PER_BINADE)
    lo = 2.0 ** (ERF_LOW_EXP + b) * (1 + m / ERF_PER_BINADE)
    hi = 2.0 ** (ERF_LOW_EXP + b) * (1 + (m + 1) / ERF_PER_BINADE)
    mid = f32(0.5 * (lo + hi))
    from scipy.special import erf

    def best_of(cand):
        v = erf(cand.astype(np.float64))                    # 53 bits: 29 beyond float32
        frac = np.abs(v - v.astype(f32).astype(np.float64)) / np.spacing(v.astype(f32)).astype(np.float64)
        res = []
        for i in np.argsort(frac)[:4]:                      # confirm the best few with mpmath
            t = mp.erf(mp.mpf(float(cand[i])))
            r32 = f32(float(t))
            res.append((abs(float(t - mp.mpf(float(r32)))) / float(np.spacing(r32)), i))
        miss, i = min(res)
        return miss, cand[i]

    miss, c = best_of((mid.view(np.int32) + np.arange(-search, search + 1, dtype=np.int32)).view(f32))
    if miss > 2e-3:            # erf nearly flat (a > 3.4): search the whole interval, then beyond it
        whole = np.arange(f32(lo).view(np.int32), f32(hi).view(np.int32), dtype=np.int32).view(f32)
        miss, c = best_of(whole)
    if miss > 2e-3:            # no float32 value of erf inside: centre above the interval (d = a - c
        ext = np.arange(f32(hi).view(np.int32), f32(1.25 * hi).view(np.int32), 8, dtype=np.int32).view(f32)
        miss, c = best_of(ext)                              # stays exact, c < 2a)
    c0 = f32(float(mp.erf(mp.mpf(float(c)))))
    # fit (erf(c + d) - C0) / d, degree 3, on the interval (the 2^-10 ulp residual of C0 is dropped)
    def g(d):
        d = np.where(d == 0, LD(1e-12), d)
        return (_erf_ld(np.float64(c) + d.astype(np.float64)) - LD(c0)) / d
    # near d = 0 the quotient loses digits in float64: use the derivative series there instead
    def g_safe(d):
        d64 = d.astype(np.float64)
        cc = np.float64(c)
        e0 = 2 / np.sqrt(np.pi) * np.exp(-cc * cc)
        series = e0 * (1 - cc * d64 + (2 * cc * cc - 1) / 3 * d64 ** 2 - (2 * cc ** 3 - 3 * cc) / 6 * d64 ** 3)
        return np.where(np.abs(d64) < 1e-3 * cc, series.astype(LD), g(d))
    fit = cheb_fit(g_safe, lo - float(c), hi - float(c), 3)
    return [c, c0] + [f32(v) for v in fit], miss


def erf_table():
    rows = np.zeros((ERF_ROWS, 6), dtype=f32)
    two_sqrtpi = 2 / np.sqrt(np.pi)
    # row 0: [0, 2^-12), centre 0: erf(a) = a (C1 + a^2 C3) (C2 = C4 = 0)
    rows[0] = [0, 0, f32(two_sqrtpi), 0, f32(-two_sqrtpi / 3), 0]
    worst = 0.0
    for j in range(1, ERF_ROWS):
        r, miss = erf_row(j)
        rows[j] = r
        worst = max(worst, miss)
    return rows, worst


def erf_gal(x, tab):
    a = np.minimum(np.abs(x), ERF_AMAX)
    bits = a.view(np.int32)
    j = np.maximum((bits >> 19) - ERF_BASE, 0)
    T = tab[j]
    d = a - T[:, 0]
    p = T[:, 5]
    p = fma(p, d, T[:, 4])
    p = fma(p, d, T[:, 3])
    p = fma(p, d, T[:, 2])
    r = fma(p, d, T[:, 1])
    return np.copysign(r, x)


# ----------------------------------------------------------------------------------------- main
def emit(tab_erf, tab_exp, tab_log, log_q):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delayrepay_b200", "csrc",
                        "math_tables.cuh")

    def flt(v):
        return f"{float(v):.9e}f"

    def block(name, flat, per_line=6):
        out = [f"__constant__ float {name}[{len(flat)}] = {{"]
        for i in range(0, len(flat), per_line):
            out.append("  " + ", ".join(flt(v) for v in flat[i:i + per_line]) + ",")
        out[-1] = out[-1].rstrip(",") + "};"
        return out

    out = ["// GENERATED by tools/gen_math_v2.py --emit: do not edit.",
           f"#define DR_ERF2_ROWS {ERF_ROWS}",
           f"#define DR_ERF2_BASE {ERF_BASE}",
           "// coefficient-major float2 pairs: (c, C0) x ROWS, (C1, C2) x ROWS, (C3, C4) x ROWS"]
    out += block("DR_ERF2_TAB", np.concatenate([tab_erf[:, 0:2].ravel(), tab_erf[:, 2:4].ravel(),
                                                tab_erf[:, 4:6].ravel()]))
    out += ["// exp: 2^(j/32) = hi + lo, j = 0..31",
            f"#define DR_EXP2_SCALE {flt(EXP_SCALE)}",
            f"#define DR_EXP2_NCHI {flt(-EXP_C_HI)}",
            f"#define DR_EXP2_NCLO {flt(-EXP_C_LO)}",
            f"#define DR_EXP2_Q0 {flt(EXP2_Q[0])}",
            f"#define DR_EXP2_Q1 {flt(EXP2_Q[1])}",
            f"#define DR_EXP2_Q2 {flt(EXP2_Q[2])}"]
    out += block("DR_EXP2_TAB", tab_exp.ravel())
    out += ["// log: 64 intervals of the offset mantissa: (r, L_hi, L_lo, 0), L = -log r",
            f"#define DR_LOG2_Q0 {flt(log_q[0])}",
            f"#define DR_LOG2_Q1 {flt(log_q[1])}",
            f"#define DR_LOG2_Q2 {flt(log_q[2])}",
            f"#define DR_LOG2_LN2HI {flt(LN2_HI)}",
            f"#define DR_LOG2_LN2LO {flt(LN2_LO)}"]
    out += block("DR_LOG2_TAB", np.concatenate([tab_log, np.zeros((LOG_N, 1), f32)], axis=1).ravel(), 4)
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    tab, worst = erf_table()
    print(f"erf: {ERF_ROWS} rows, worst |erf(c) - C0| = {worst:.2e} ulp")
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-4.5, 4.5, 1 << 22), rng.uniform(-0.13, 0.13, 1 << 21),
                        np.exp(rng.uniform(-40, 1.5, 1 << 21)) * rng.choice([-1, 1], 1 << 21),
                        np.linspace(2.0 ** -13, 4.0, 1 << 21)]).astype(f32)
    e = ulp_err(erf_gal(x, tab), _erf_ld(x))
    i = e.argmax()
    tiny = np.abs(x) < 2.0 ** ERF_LOW_EXP
    print(f"erf f32 (Gal table): max {e[~tiny].max():.4f} ulp for |x| >= 2^{ERF_LOW_EXP}, "
          f"{e[tiny].max():.4f} ulp below (row 0), mean {e.mean():.4f}; worst at {x[i]!r}")


# ----------------------------------------------------------------------------------------- exp
# e^x = 2^(k/32) e^r,  k = rint(32 x / ln 2),  r = x - k ln2/32 (two steps, the first exact).
#   2^(k/32) = 2^(k >> 5) T[k & 31],  T = T_hi + T_lo;   e^r - 1 = p = r + r^2 (c2 + r (c3 + r c4))
#   result = T_hi + (T_hi p + T_lo)        |r| <= ln2/64: every error after the first term is
# scaled by <= 0.011, so the result is within 0.51 ulp.  9 float ops + 1 LDS.64 (was 17).
EXP_N = 32
EXP_MAGIC = f32(12582912.0)
EXP_SCALE = f32(EXP_N / np.log(2.0))
EXP_C_HI = f32(np.float32(np.log(2.0) / EXP_N).view(np.int32) & np.int32(-8192)).view(f32) \
    if False else (np.array([np.log(2.0) / EXP_N], dtype=f32).view(np.int32) & np.int32(-8192)).view(f32)[0]
EXP_C_LO = f32(np.log(LD(2)) / EXP_N - LD(EXP_C_HI))


def exp_table():
    t = np.zeros((EXP_N, 2), dtype=f32)
    for j in range(EXP_N):
        v = mp.power(2, mp.mpf(j) / EXP_N)
        hi = f32(float(v))
        lo = f32(float(v - mp.mpf(float(hi))))
        t[j] = [hi, lo]
    return t


def _expm1_q(r):            # (e^r - 1 - r) / r^2
    small = np.abs(r) < 1e-3
    rr = np.where(small, LD(1), r)
    series = LD(1) / 2 + r / 6 + r * r / 24 + r * r * r / 120 + r ** 4 / 720
    return np.where(small, series, (np.expm1(rr) - rr) / (rr * rr))
EXP2_Q = cheb_fit(_expm1_q, -0.0109, 0.0109, 2).astype(f32)


def exp_tab(x, tab):
    kf = fma(x, EXP_SCALE, EXP_MAGIC)
    kfl = kf - EXP_MAGIC
    r = fma(kfl, -EXP_C_HI, x)
    r = fma(kfl, -EXP_C_LO, r)
    q = fma(r, f32(EXP2_Q[2]), f32(EXP2_Q[1]))
    q = fma(r, q, f32(EXP2_Q[0]))
    t = r * r
    p = fma(t, q, r)
    k = kf.view(np.int32) - EXP_MAGIC.view(np.int32)
    T = tab[k & (EXP_N - 1)]
    s = fma(T[:, 0], p, T[:, 1])
    res = T[:, 0] + s
    return (res.view(np.int32) + ((k >> 5) << 23)).view(f32)


# ----------------------------------------------------------------------------------------- log
# x = 2^e m, m in [sqrt(1/2), sqrt 2) (bit pattern offset 0x3f3504f3); 64 intervals by the top six
# bits of the offset mantissa; r_j ~ 1/m with <= 7 significant bits, so f = m r_j - 1 is EXACT in
# one fma and |f| < 2^-5.6;  L_j = -log r_j = L_hi + L_lo, L_hi a multiple of 2^-16, so that
# B = e LN2_HI + L_hi is exact;  log x = B + f + (f^2 (c2 + f (c3 + f c4)) + L_lo + e LN2_LO) with
# one fast-two-sum for B + f.  14 float ops + LDS.64 + LDS.32 (was 30), 0.55 ulp.
LOG_N = 64
LN2_HI = f32(0.693145751953125)          # 16 significant bits
LN2_LO = f32(np.log(LD(2)) - LD(0.693145751953125))
LOG_OFF = np.int32(0x3f3504f3)


def log_table():
    t = np.zeros((LOG_N, 3), dtype=f32)
    worst_f = 0.0
    for j in range(LOG_N):
        lo_bits = LOG_OFF + np.int32(j << 17)
        hi_bits = LOG_OFF + np.int32(((j + 1) << 17) - 1)
        m = np.arange(lo_bits, hi_bits + 1, dtype=np.int64).astype(np.int32).view(f32).astype(np.float64)
        if m[0] <= 1.0 <= m[-1]:
            r = 1.0
        else:
            mid = 0.5 * (m[0] + m[-1])
            best = None
            for bits in (5, 6, 7):
                scale = 2.0 ** bits if 1.0 / mid >= 1 else 2.0 ** (bits + 1)
                for n in np.round(1.0 / mid * scale) + np.arange(-3, 4):
                    rr = n / scale
                    f = m * rr - 1.0                   # exact in float64 (24 + 8 bits)
                    exact = np.all(f.astype(f32).astype(np.float64) == f)
                    # fast-two-sum of B + f needs |B| >= |f| for e = 0, i.e. |L_hi| >= max |f|
                    l_hi = abs(np.round(-np.log(rr) * 65536.0) / 65536.0)
                    if exact and l_hi >= np.abs(f).max() and (best is None or np.abs(f).max() < best[0]):
                        best = (np.abs(f).max(), rr)
            assert best is not None, j
            r = best[1]
        f = m * r - 1.0
        assert np.all(f.astype(f32).astype(np.float64) == f), j
        worst_f = max(worst_f, np.abs(f).max())
        L = -mp.log(mp.mpf(r))
        l_hi = np.round(float(L) * 65536.0) / 65536.0
        l_lo = f32(float(L - mp.mpf(l_hi)))
        t[j] = [f32(r), f32(l_hi), l_lo]
        assert float(f32(l_hi)) == l_hi
    return t, worst_f


def _log1p_q(f):            # (log1p(f) - f) / f^2
    small = np.abs(f) < 1e-3
    ff = np.where(small, LD(1), f)
    series = -LD(1) / 2 + f / 3 - f * f / 4 + f ** 3 / 5 - f ** 4 / 6
    return np.where(small, series, (np.log1p(ff) - ff) / (ff * ff))


def log_tab(x, tab, Q):
    ix = x.view(np.int32) - LOG_OFF
    e = ix >> 23
    m = ((ix & 0x007fffff) + LOG_OFF).view(f32)
    j = (ix >> 17) & (LOG_N - 1)
    T = tab[j]
    ef = e.astype(f32)
    f = fma(m, T[:, 0], f32(-1))
    B = fma(ef, LN2_HI, T[:, 1])
    q = fma(f, f32(Q[2]), f32(Q[1]))
    q = fma(f, q, f32(Q[0]))
    t = f * f
    C = B + f
    err = (B - C) + f
    w = fma(ef, LN2_LO, T[:, 2])
    w = w + err
    tail = fma(t, q, w)
    return C + tail


def report_exp_log():
    et = exp_table()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-87, 88, 1 << 22), rng.uniform(-1, 1, 1 << 21),
                        rng.uniform(-0.2, 0.0, 1 << 21)]).astype(f32)
    e = ulp_err(exp_tab(x, et), np.exp(x.astype(LD)))
    print(f"exp f32 (table {EXP_N}): max {e.max():.4f} ulp at {x[e.argmax()]!r}, mean {e.mean():.4f}")
    lt, worst_f = log_table()
    Q = cheb_fit(_log1p_q, -worst_f * 1.01, worst_f * 1.01, 2).astype(f32)
    x = np.concatenate([rng.uniform(0, 4, 1 << 21), np.exp(rng.uniform(-80, 80, 1 << 21)),
                        rng.uniform(0.7, 1.45, 1 << 22), rng.uniform(0.96, 1.04, 1 << 21)]).astype(f32)
    x = x[x > 1e-37]
    e = ulp_err(log_tab(x, lt, Q), np.log(x.astype(LD)))
    print(f"log f32 (table {LOG_N}): max |f| {worst_f:.5f}, max {e.max():.4f} ulp at {x[e.argmax()]!r}, "
          f"mean {e.mean():.4f}")
    return et, lt, Q


if __name__ == "__main__":
    _et, _lt, _lq = report_exp_log()
    if "--emit" in sys.argv:
        emit(tab, _et, _lt, _lq)
