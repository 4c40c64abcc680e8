"""Comparators of the parity suite: bit equality, ulp distance, relative tolerance."""
import numpy as np


def bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def assert_bits_equal(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
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
