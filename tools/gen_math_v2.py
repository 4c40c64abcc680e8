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
def emit(tab_erf):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "delayrepay_b200", "csrc",
                        "math_tables.cuh")
    out = ["// GENERATED by tools/gen_math_v2.py --emit: do not edit.",
           f"#define DR_ERF2_ROWS {ERF_ROWS}",
           f"#define DR_ERF2_BASE {ERF_BASE}",
           "// coefficient-major float2 pairs: (c, C0) x ROWS, (C1, C2) x ROWS, (C3, C4) x ROWS",
           f"__constant__ float DR_ERF2_TAB[{ERF_ROWS * 6}] = {{"]
    flat = np.concatenate([tab_erf[:, 0:2].ravel(), tab_erf[:, 2:4].ravel(), tab_erf[:, 4:6].ravel()])
    for i in range(0, len(flat), 6):
        out.append("  " + ", ".join(f"{float(v):.9e}f" for v in flat[i:i + 6]) + ",")
    out[-1] = out[-1].rstrip(",") + "};"
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
    if "--emit" in sys.argv:
        emit(tab)
