"""Keyword arguments of the reduction handlers that used to be swallowed (`initial=`, `where=`,
1-d `weights=` along an axis, tuple shifts of `np.roll`, `np.clip(min=, max=)`): implemented as
compositions of fused primitives, everything else raises NotImplementedError -- never a silently
different result.  (Written after this round's GPU budget was spent: shapes and dtypes are
checked on the CPU by tools/fuzz_shapes.py-style dry runs, values here; the file sorts last so
that it cannot mask the rest of the suite under `pytest -x`.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_initial_where_weights_and_roll_keywords(gpu):
    rng = np.random.default_rng(17)
    x = rng.standard_normal((37, 53))
    i = rng.integers(-20, 20, (37, 53)).astype(np.int32)
    X, I = gpu.array(x), gpu.array(i)
    m = x > 0.3
    cases = [
        (np.sum(X, where=X > 0.3), np.sum(x, where=m)),
        (np.sum(X, axis=0, where=X > 0.3), np.sum(x, axis=0, where=m)),
        (np.prod(X[:4, :5], where=X[:4, :5] > 0.3), np.prod(x[:4, :5], where=m[:4, :5])),
        (np.max(X, axis=1, where=X > 0.3, initial=-1.0), np.max(x, axis=1, where=m, initial=-1.0)),
        (np.min(X, where=gpu.array(m), initial=9.0), np.min(x, where=m, initial=9.0)),
        (np.sum(X, initial=5.0), np.sum(x, initial=5.0)),
        (np.max(X, initial=100.0), np.max(x, initial=100.0)),
        (np.min(X, axis=0, initial=-0.5), np.min(x, axis=0, initial=-0.5)),
        (np.sum(I, initial=7), np.sum(i, initial=7)),
        (np.max(I, axis=0, initial=2.5), np.max(i, axis=0, initial=2.5)),
        (np.prod(I[:3, :3], initial=2), np.prod(i[:3, :3], initial=2)),
        (X.sum(axis=1, initial=1.0), x.sum(axis=1, initial=1.0)),
        (np.average(X, axis=0, weights=np.arange(1.0, 38.0)), np.average(x, axis=0, weights=np.arange(1.0, 38.0))),
        (np.average(X, axis=1, weights=gpu.array(np.arange(1.0, 54.0))), np.average(x, axis=1, weights=np.arange(1.0, 54.0))),
        (np.average(X, weights=gpu.array(np.abs(x))), np.average(x, weights=np.abs(x))),
        (np.roll(X, (1, 2), axis=(0, 1)), np.roll(x, (1, 2), axis=(0, 1))),
        (np.roll(X, 3, axis=(0, 1)), np.roll(x, 3, axis=(0, 1))),
        (np.roll(X, (2, -5)), np.roll(x, (2, -5))),
        (np.clip(X, min=-0.5, max=0.25), np.clip(x, min=-0.5, max=0.25)),
        (np.sum(I, where=I > 3), np.sum(i, where=i > 3)),
    ]
    for k, (got, want) in enumerate(cases):
        got = got.get()
        want = np.asarray(want)
        assert got.shape == want.shape and got.dtype == want.dtype, (k, got.shape, got.dtype, want.shape, want.dtype)
        if want.dtype.kind == "f":
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, err_msg=str(k))
        else:
            assert np.array_equal(got, want), k


def test_unsupported_keywords_raise(gpu):
    X = gpu.array(np.arange(12.0).reshape(3, 4))
    with pytest.raises(ValueError):                     # NumPy's own rule: max + where needs initial
        np.max(X, where=X > 3)
    for call in (lambda: np.mean(X, where=X > 3), lambda: np.var(X, where=X > 3), lambda: np.any(X > 3, where=X > 5),
                 lambda: np.sum(X, out=np.zeros(())), lambda: np.repeat(X, [1, 2, 3], axis=0),
                 lambda: np.linalg.norm(X.ravel(), ord=1)):
        with pytest.raises(NotImplementedError):
            call()
