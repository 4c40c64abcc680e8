"""Round-2 regressions: stale memo hits after host / self writes, np.take
along a non-zero axis, axis validation, ufunc.reduce keywords, np.average argument checks and
complex parts of fft results.  Shapes are checked in dry-run mode on the CPU; values on the GPU."""
import numpy as np
import pytest

import delayrepay_b200 as dr
from delayrepay_b200 import engine


def test_take_shapes_match_numpy():
    x = np.zeros((3, 4, 5), np.float32)
    cases = [(2, 1), (np.array([[0, 1], [2, 3]]), 1), ([1, 2], 2), (1, -1), ([0, 2], 0), (3, None)]
    with engine.dry_run():
        d = dr.array(x)
        for idx, ax in cases:
            assert np.take(d, idx, axis=ax).shape == np.take(x, idx, axis=ax).shape, (idx, ax)
        y = dr.array(np.zeros((3, 4)))
        assert np.take(y, 1, axis=1).shape == (3,)


def test_axes_are_validated_like_numpy():
    with engine.dry_run():
        y = dr.array(np.zeros((3, 4)))
        for bad in (5, -3, 2):
            with pytest.raises(np.exceptions.AxisError):
                np.sum(y, axis=bad)
        with pytest.raises(ValueError):
            np.sum(y, axis=(0, 0))
        with pytest.raises(TypeError):
            np.average(y, weights=np.ones(4))            # shapes differ, no axis
        with pytest.raises(ValueError):
            np.average(y, weights=np.ones(3), axis=1)    # wrong length
        assert np.average(y, weights=np.ones(4), axis=1).shape == (3,)
        assert np.add.reduce(y, initial=5).shape == (4,)
        assert np.add.reduce(y, axis=1, keepdims=True).shape == (3, 1)


def test_memo_is_not_hit_after_self_write_or_for_host_operands():
    with engine.dry_run():
        a = dr.array(np.arange(8.0))
        c = a * 2
        c._force()
        c[1:3] = 0
        assert (a * 2) is not c                 # c was written in place: a fresh capture recomputes
        assert np.sin(a) is np.sin(a)           # hash-consing of untouched nodes is kept
        h = np.ones(8)
        r = a + h
        assert (a + h) is r                     # reference tests/test.py:145-149: host leaves memoise
        r._force()                              # ... until they have been uploaded
        assert (a + h) is not r


@pytest.mark.gpu
def test_advice_values(gpu):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 4, 5)).astype(np.float32)
    d = dr.array(x)
    for idx, ax in [(2, 1), (np.array([[0, 1], [2, 3]]), 1), ([1, 2], 2), ([0, 2], 0), (7, None)]:
        np.testing.assert_array_equal(np.take(d, idx, axis=ax).get(), np.take(x, idx, axis=ax))
    # self-write invalidates the memoised node, holders of the old node keep its (mutated) storage
    a = dr.array(np.arange(8.0))
    c = a * 2
    c._force()
    c[1:3] = 0
    np.testing.assert_array_equal((a * 2).get(), np.arange(8.0) * 2)
    np.testing.assert_array_equal(c.get(), np.where((np.arange(8) >= 1) & (np.arange(8) < 3), 0, np.arange(8.0) * 2))
    # host operand mutated between captures
    h = np.ones(8)
    r = a + h
    r._force()
    h[:] = 7
    np.testing.assert_array_equal((a + h).get(), np.arange(8.0) + 7)
    y = rng.standard_normal((3, 4))
    dy = dr.array(y)
    np.testing.assert_allclose(np.add.reduce(dy, initial=5).get(), np.add.reduce(y, initial=5), rtol=1e-12)
    np.testing.assert_allclose(np.average(dy, weights=np.arange(1.0, 5.0), axis=1).get(),
                               np.average(y, weights=np.arange(1.0, 5.0), axis=1), rtol=1e-12)


@pytest.mark.gpu
def test_complex_parts_of_fft_results(gpu):
    rng = np.random.default_rng(8)
    x = rng.standard_normal(256)
    f = dr.fft.fft(dr.array(x))
    want = np.fft.fft(x)
    np.testing.assert_allclose(f.real.get(), want.real, atol=1e-9)
    np.testing.assert_allclose(f.imag.get(), want.imag, atol=1e-9)
    np.testing.assert_allclose(f.conj().get(), want.conj(), atol=1e-9)


def test_negative_zero_scalar_is_not_hash_consed_with_positive_zero():
    """-0.0 == 0.0 and hash(-0.0) == hash(0.0): the memo table used to hand back the Scalar node of
    whichever zero was captured first (np.clip(x, -0.0, 0.0) then clipped to [+0.0, +0.0])."""
    import numpy as np
    import delayrepay_b200 as dr
    for ty in (float, np.float32, np.float64):
        p, n = dr.Scalar(ty(0.0)), dr.Scalar(ty(-0.0))
        assert p is not n and not np.signbit(p.val) and np.signbit(n.val)
        assert dr.Scalar(ty(-0.0)) is n and dr.Scalar(ty(0.0)) is p


def test_numpy_bool_scalar_operands_are_captured():
    """np.True_ is not a numbers.Number: `np.where(np.less_equal(-3, -3), x, y)` raised
    NotImplementedError (fuzz seed 20358)."""
    import numpy as np
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine
    with engine.dry_run():
        x = dr.array(np.ones(8, np.float32))
        r = np.where(np.less_equal(-3, -3), x, x * 2)
        assert isinstance(r, dr.DelayArray) and r.dtype == np.float32 and r.shape == (8,)
        r.run()
        assert (x * np.True_).dtype == np.float32 and (x + np.bool_(False)).shape == (8,)


@pytest.mark.gpu
def test_gather_and_mask_results_are_not_cached_as_views(gpu):
    """The per-leaf view cache (round 2) also kept the RESULT of `a[idx]` / `a[mask]` under the
    index object: the second `a[idx]` after a write to `a` returned the stale copy."""
    import numpy as np
    h = np.arange(40, dtype=np.float32).reshape(10, 4)
    a = gpu.array(h.copy())
    idx = gpu.array(np.array([1, 3, 3, 0]))
    mask = gpu.array(h[:, 0] > 10)
    first, firstm = a[idx].get(), a[mask].get()
    a[...] = a * 2.0
    assert np.array_equal(first, h[[1, 3, 3, 0]]) and np.array_equal(a[idx].get(), 2 * h[[1, 3, 3, 0]])
    assert np.array_equal(firstm, h[h[:, 0] > 10]) and np.array_equal(a[mask].get(), 2 * h[h[:, 0] > 10])
    assert a[2:5] is a[2:5] and a[1, None] is a[1, None]          # basic indices still share one view node


def test_where_wraps_python_ints_that_do_not_fit_like_numpy():
    """np.where(c, uint8_array, -3) holds 253 (fuzz seed 71979 raised OverflowError in the planner)."""
    import numpy as np
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine
    with engine.dry_run():
        x = dr.array(np.ones(8, np.uint8))
        r = np.where(x > 0, x, -3)
        assert r.dtype == np.uint8 and r.children[2].val == 253 and r.children[2].dtype == np.uint8
        r.run()
        y = np.where(dr.array(np.ones(8, np.int8)) > 0, 300, dr.array(np.ones(8, np.int8)))
        assert y.dtype == np.int8 and y.children[1].val == 44
        y.run()
