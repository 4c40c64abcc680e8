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
