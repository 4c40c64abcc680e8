"""Design + accuracy check of the PACKED FLOAT32 exp / log of prelude.cuh (dr_exp4_f32,
dr_log4_f32): every operation below is one float32 instruction on the device (emulated here
with NumPy float32 arithmetic; fma = exact product and sum in float64, rounded once -- double
rounding affects ~2^-29 of the cases and is ignored).  Prints max error in float32 ulps
against long-double truth.  Run: python tools/gen_math_f32.py"""
import numpy as np
from gen_math import cheb_fit, LD, fmt

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def ulp_err(got, truth):
    u = np.spacing(np.abs(truth.astype(f32))).astype(LD)
    return np.abs(got.astype(LD) - truth) / u


LOG2E = f32(1.4426950408889634)
MAGIC = f32(12582912.0)
LN2_HI = f32(0.693145751953125)
LN2_LO = f32(np.log(LD(2)) - LD(0.693145751953125))

# ------------------------------------------------------------------ exp
# e^r = 1 + r + r^2 q(r), q degree 4 on |r| <= ln2/2
def q_target(r):
    small = np.abs(r) < 1e-3
    rr = np.where(small, LD(1), r)
    series = LD(0.5) + r / 6 + r * r / 24 + r * r * r / 120
    return np.where(small, series, (np.expm1(rr) - rr) / (rr * rr))
EXP_Q = cheb_fit(q_target, -0.3466 * 1.0001, 0.3466 * 1.0001, 4).astype(f32)

def exp_f32(x):
    kf = fma(x, LOG2E, MAGIC)
    kfl = kf - MAGIC
    r = fma(kfl, -LN2_HI, x)
    rl = kfl * (-LN2_LO)
    r2 = r + rl                      # only used in the (insensitive) higher-order terms
    q = f32(EXP_Q[4])
    for c in EXP_Q[3::-1]:
        q = fma(q, r2, f32(c))
    t = r2 * r2
    sl = fma(t, q, rl)               # r2^2 q + rl
    a = f32(1) + r                   # 1 + r + (rl + r2^2 q), the leading terms kept exact
    e1 = (f32(1) - a) + r
    c = e1 + sl
    res = a + c
    k = (kf.view(np.int32) - MAGIC.view(np.int32)).astype(np.int32)
    return (res.view(np.int32) + (k << 23)).view(f32)

# ------------------------------------------------------------------ log
# log1p(f) = f - f^2/2 + f^3 P(f), P degree 8 on [sqrt(.5)-1, sqrt(2)-1]
def p_target(f):
    f = np.where(f == 0, LD(1e-30), f)
    return (np.log1p(f) - f + f * f / 2) / (f * f * f)
LOG_DEG = 8
LOG_P = cheb_fit(p_target, np.sqrt(0.5) - 1 - 1e-4, np.sqrt(2) - 1 + 1e-4, LOG_DEG).astype(f32)

def log_f32(x):
    u = x.view(np.int32)
    ix = u - np.int32(0x3f3504f3)
    e = ix >> 23
    m = ((ix & 0x007fffff) + np.int32(0x3f3504f3)).view(f32)
    f = m - f32(1)
    ef = e.astype(f32)
    p = f32(LOG_P[LOG_DEG])
    for c in LOG_P[LOG_DEG - 1::-1]:
        p = fma(p, f, f32(c))
    th = f * f
    tl = fma(f, f, -th)                 # exact low part of f^2
    h = f32(-0.5) * th                  # exact
    g = th * (f * p)                    # f^3 P
    g = fma(f32(-0.5), tl, g)
    g = fma(ef, LN2_LO, g)
    yh = ef * LN2_HI                    # exact
    a1 = f + h                          # |f| >= |h|
    e1 = (f - a1) + h
    a2 = yh + a1                        # |yh| >= |a1| or yh == 0
    e2 = (yh - a2) + a1
    c = (e1 + e2) + g
    return a2 + c


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-87, 88, 1 << 22), rng.uniform(-1, 1, 1 << 20)]).astype(f32)
    e = ulp_err(exp_f32(x), np.exp(x.astype(LD)))
    print(f"exp f32: max {e.max():.4f} ulp, mean {e.mean():.4f}   q = [{fmt(EXP_Q)}]")
    x = np.concatenate([rng.uniform(0, 4, 1 << 21), np.exp(rng.uniform(-80, 80, 1 << 21)),
                        rng.uniform(0.7, 1.45, 1 << 21)]).astype(f32)
    x = x[x > 1e-37]
    e = ulp_err(log_f32(x), np.log(x.astype(LD)))
    i = e.argmax()
    print(f"log f32: max {e.max():.4f} ulp at {x[i]!r}, mean {e.mean():.4f}   p = [{fmt(LOG_P)}]")


# ------------------------------------------------------------------ erf: table-driven, float32 only
# a = |x|.  a < 1/8: a*c0h + a*(c0l + s P(s)), s = a^2.  1/8 <= a < 4: 40 intervals (8 per binade,
# index from the float's exponent + top 3 mantissa bits), centre c_j, d = a - c_j (exact):
#   erf(a) = C0h_j + (C0l_j + d (C1_j + d (C2_j + d (C3_j + d (C4_j + d C5_j)))))
from scipy.special import erf as _erf
C0 = 2 / np.sqrt(np.pi)
ERF_C0H = f32(C0)
ERF_C0L = f32(C0 - float(ERF_C0H))


def _small_target(s):        # (erf(a)/a - c0) / s  with s = a^2
    a = np.sqrt(s.astype(np.float64))
    a = np.where(a == 0, 1e-6, a)
    return ((_erf(a) / a - C0) / (a * a)).astype(LD)

# Taylor: erf(a)/a = c0 (1 - s/3 + s^2/10 - s^3/42 ...); fit P(s) = -c0/3 + c0 s/10 - ...
def _small_series(s):
    s = s.astype(LD)
    return LD(C0) * (-LD(1) / 3 + s / 10 - s * s / 42 + s * s * s / 216)
ERF_SMALL = cheb_fit(_small_series, 0.0, 1.0 / 64, 2).astype(f32)


def erf_table():
    rows = []
    for j in range(40):
        e, m = divmod(j, 8)
        lo = 2.0 ** (e - 3) * (1 + m / 8)
        hi = 2.0 ** (e - 3) * (1 + (m + 1) / 8)
        c = 0.5 * (lo + hi)
        fit = cheb_fit(lambda d, c=c: _erf((c + d).astype(np.float64)).astype(LD),
                       lo - c, hi - c, 5)
        c0h = f32(fit[0])
        c0l = f32(fit[0] - float(c0h))
        rows.append([c0h, c0l] + [f32(v) for v in fit[1:]] + [f32(0)])
    return np.array(rows, dtype=f32)            # (40, 8): C0h C0l C1 C2 C3 C4 C5 pad
ERF_TAB = erf_table()


def erf_f32(x):
    a = np.minimum(np.abs(x), f32(3.9999998))
    bits = a.view(np.int32)
    j = np.clip((bits >> 20) - 0x3e0, 0, 39)
    c = ((bits & np.int32(-1048576)) | np.int32(0x00080000)).view(f32)
    d = a - c
    T = ERF_TAB[j]
    p = T[:, 6]
    for k in (5, 4, 3, 2):
        p = fma(p, d, T[:, k])
    t = fma(p, d, T[:, 1])
    big = T[:, 0] + t
    s = a * a
    q = f32(ERF_SMALL[2])
    q = fma(q, s, f32(ERF_SMALL[1]))
    q = fma(q, s, f32(ERF_SMALL[0]))
    q = fma(q, s, ERF_C0L)
    small = fma(a, ERF_C0H, a * q)
    return np.copysign(np.where(a < f32(0.125), small, big), x)


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-4.5, 4.5, 1 << 22), rng.uniform(-0.13, 0.13, 1 << 21),
                        np.exp(rng.uniform(-40, 1.5, 1 << 21)) * rng.choice([-1, 1], 1 << 21),
                        np.linspace(0.1249, 4.0, 1 << 21)]).astype(f32)
    truth = _erf(x.astype(np.float64)).astype(LD)
    e = ulp_err(erf_f32(x), truth)
    i = e.argmax()
    print(f"erf f32 (table): max {e.max():.4f} ulp at {x[i]!r}, mean {e.mean():.4f}")
    ok = np.abs(x) < 0.125
    print("   small branch max", e[ok].max(), " table branch max", e[~ok].max())
