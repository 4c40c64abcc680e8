"""Comparators of the parity suite: bit equality, ulp distance, relative tolerance."""
import numpy as np


def bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def assert_bits_equal(got, want, what=""):
    """Bitwise equality; NaNs compare equal to NaNs (x86 and CUDA differ in the sign/payload of
    a generated NaN -- 0xFFC00000 vs 0x7FFFFFFF -- which NumPy semantics do not define)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    if got.dtype.kind == "f":
        gn, wn = np.isnan(got), np.isnan(want)
        assert np.array_equal(gn, wn), f"{what}: NaN positions differ"
        if gn.any():
            got, want = np.where(gn, 0, got), np.where(wn, 0, want)
    if got.tobytes() != want.tobytes():
        bad = np.flatnonzero(got.ravel().view(_uint(got.dtype)) != want.ravel().view(_uint(want.dtype)))
        i = bad[0]
        raise AssertionError(f"{what}: {bad.size}/{got.size} elements differ bitwise; first at "
                             f"{i}: got {got.ravel()[i]!r} want {want.ravel()[i]!r}")


def _uint(dt):
    return {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[np.dtype(dt).itemsize]


def ulp_distance(a, b):
    """Units-in-the-last-place distance between same-dtype float arrays (NaN==NaN -> 0)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype and a.dtype.kind == "f"
    it = np.int32 if a.dtype == np.float32 else np.int64
    ia, ib = a.view(it).astype(np.int64), b.view(it).astype(np.int64)
    sign = np.int64(np.iinfo(it).min)
    ia = np.where(ia < 0, sign - ia, ia)
    ib = np.where(ib < 0, sign - ib, ib)
    d = np.abs(ia - ib)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0, d)


def assert_ulp(got, want, max_ulp, what=""):
    d = ulp_distance(got, want)
    worst = int(d.max()) if d.size else 0
    assert worst <= max_ulp, f"{what}: max ulp distance {worst} > {max_ulp} " \
                             f"({(d > max_ulp).sum()} of {d.size} elements)"
    return worst


def erf_exact(x):
    """erf of float64 points to ~40 digits (Taylor series in decimal arithmetic): the
    higher-precision truth for the few points where two double implementations disagree."""
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    out = []
    two_over_sqrt_pi = Decimal(2) / Decimal("3.14159265358979323846264338327950288419716939937510582097494").sqrt()
    for v in np.asarray(x, dtype=np.float64).ravel():
        d = Decimal(float(v))
        term, total, n = d, d, 0
        while abs(term) > Decimal(10) ** -55:
            n += 1
            term = -term * d * d / n
            total += term / (2 * n + 1)
        out.append(total * two_over_sqrt_pi)
    return out


def assert_close_to_numpy_or_truth(got, want, truth_fn, x_args, limit, what):
    """<= `limit` ulp from NumPy; where NumPy itself is > 1 ulp from the true value (its SIMD
    float32 kernels are documented up to ~4 ulp), we must instead be within 1 ulp of the truth.
    ``truth_fn(*x_args)`` evaluates in float64 (for float32 data)."""
    d = ulp_distance(got, want)
    if d.size == 0 or d.max() <= limit:
        return int(d.max()) if d.size else 0
    assert got.dtype == np.float32, f"{what}: {int(d.max())} ulp from NumPy"
    bad = d > limit
    truth = truth_fn(*[a.astype(np.float64) for a in x_args])[bad]
    ulp = np.spacing(np.abs(truth).astype(np.float32)).astype(np.float64)
    ours = np.abs(got[bad].astype(np.float64) - truth) / ulp
    theirs = np.abs(want[bad].astype(np.float64) - truth) / ulp
    assert np.all(ours <= 1.0) and np.all(theirs > 1.0), (what, ours.max(), theirs.min())
    print(f"   {what}: {bad.sum()} element(s) > {limit} ulp from NumPy; there NumPy is up to "
          f"{theirs.max():.2f} ulp from the float64 truth, ours {ours.max():.2f}")
    return int(d.max())
