"""The NumPy surface around the hot path (callers): every case is the
same NumPy-level program run on host arrays by NumPy (what the reference's CPU backend would
compute after forcing, delayarray.py:511-568) and on DelayArrays by the engine; results must be
identical (arithmetic only, so bit-exact), shapes and dtypes included.  The reference raises
KeyError for most of these functions; here they are device implementations (no host fallback).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A = np.random.default_rng(5).standard_normal((6, 8))
B = np.random.default_rng(6).standard_normal((6, 8))
V = np.random.default_rng(7).standard_normal(48)
I = np.random.default_rng(8).integers(-9, 9, (6, 8))

CASES = {
    "any_all": lambda a, b, v, i: (np.any(a > 2.5), np.all(a > -9.0), np.any(a > 0, axis=0), np.all(i, axis=1),
                                   (a > 0).any(), (a > 0).all(axis=1)),
    "count_nonzero": lambda a, b, v, i: (np.count_nonzero(i), np.count_nonzero(i, axis=0)),
    "argmax_argmin": lambda a, b, v, i: (np.argmax(v), np.argmin(v), np.argmax(a, axis=0), np.argmin(a, axis=1),
                                         np.argmax(i), np.argmax(a * b), a.argmax(axis=1), np.argmax(i, axis=1)),
    "ptp_trace": lambda a, b, v, i: (np.ptp(a), np.ptp(a, axis=0), np.trace(a[:, :6]), np.trace(a, 1)),
    "reshape_family": lambda a, b, v, i: (np.reshape(v, (6, 8)) + a, np.ravel(a) * v, a.ravel(), a.T.flatten(),
                                          np.squeeze(a[:, None, :]) - b, np.expand_dims(v, 0), np.expand_dims(a, (0, 2)),
                                          np.swapaxes(a, 0, 1), np.moveaxis(a[None], 0, 2), a.swapaxes(1, 0) * 2.0,
                                          np.squeeze(a[None], axis=0)),
    "broadcast_to": lambda a, b, v, i: (np.broadcast_to(a[0], (3, 8)) + 1.0, np.broadcast_to(v[:8], (6, 8)) * a),
    "concatenate": lambda a, b, v, i: (np.concatenate([a, b]), np.concatenate([a, b * 2, a - b], axis=1),
                                       np.concatenate([v, v[:5] + 1.0]), np.concatenate([a, i]),
                                       np.concatenate([a, b], axis=None), np.stack([a, b]), np.stack([v, v * 2], axis=1),
                                       np.vstack([v, v]), np.hstack([a, b]), np.hstack([v, v])),
    "outer_inner_vdot": lambda a, b, v, i: (np.outer(v[:5], v[5:9]), np.outer(a, b[0])),
    "diff": lambda a, b, v, i: (np.diff(v), np.diff(a, axis=0), np.diff(a, n=2), np.diff(i), np.diff(a > 0)),
    "round": lambda a, b, v, i: (np.round(a), np.round(a * 100, 1), np.around(a, 2), np.round(a * 1000, -2), a.round(3),
                                 np.round(i)),
    "isclose": lambda a, b, v, i: (np.isclose(a, a + 1e-9), np.isclose(a, b), np.isclose(a / (i * 1.0), a / (i * 1.0)),
                                   np.isclose(a / (i * 1.0), a / (i * 1.0), equal_nan=True)),
    "shape_size_ndim": lambda a, b, v, i: (np.shape(a + b), np.size(a), np.size(a, 1), np.ndim(a * 2), (a + b).nbytes,
                                           (a + b).itemsize, len(a + b)),
    "like": lambda a, b, v, i: (np.zeros_like(a), np.ones_like(a + b), np.full_like(i, 7), np.full_like(a, 2.5),
                                np.zeros_like(a, dtype=np.float32), np.empty_like(a).shape, np.ones_like(i, shape=(3,))),
    "divmod": lambda a, b, v, i: divmod(a, b) + divmod(i, 4) + (np.divmod(a, 0.75)[1],),
    "out_kw": lambda a, b, v, i: (np.add(a, b, out=np.zeros_like(a)), np.multiply(a, 2.0, out=np.empty_like(a)),
                                  np.sqrt(abs(a), out=np.ones_like(a))),
    "copy_real": lambda a, b, v, i: (np.copy(a), a.real, a.imag, a.conj(), (a + b).copy()),
    "iteration": lambda a, b, v, i: tuple(row * 2.0 for row in a) + ([float(x) for x in v[:4]],),
    "tolist_item": lambda a, b, v, i: ((a + b).tolist(), np.sum(i).item(), (a * 2)[1, 2].item(), v[3].item()),
    "methods": lambda a, b, v, i: (a.clip(-0.5, 0.5), v.cumsum(), (a * b).squeeze(), a.any(), i.all()),
}


RTOL = {"ptp_trace": 1e-12}        # trace is a float64 sum: the reduction bar, not bit-exactness


def _same(got, want, where):
    if isinstance(want, (tuple, list)) and not isinstance(got, np.ndarray):
        assert len(got) == len(want), where
        for k, (g, w) in enumerate(zip(got, want)):
            _same(g, w, f"{where}[{k}]")
        return
    if hasattr(got, "get"):
        got = got.get()
    if isinstance(want, np.ndarray) or isinstance(want, np.generic):
        got, want = np.asarray(got), np.asarray(want)
        assert got.shape == want.shape, f"{where}: shape {got.shape} != {want.shape}"
        assert got.dtype == want.dtype, f"{where}: dtype {got.dtype} != {want.dtype}"
        rtol = RTOL.get(where.split("[")[0])
        if rtol:
            np.testing.assert_allclose(got, want, rtol=rtol, atol=0, err_msg=where)
        else:
            assert np.array_equal(got, want, equal_nan=want.dtype.kind == "f"), f"{where}: values differ"
    else:
        assert got == want, f"{where}: {got!r} != {want!r}"


@pytest.mark.parametrize("name", sorted(CASES))
def test_numpy_surface_matches_numpy(gpu, name):
    with np.errstate(all="ignore"):
        want = CASES[name](A.copy(), B.copy(), V.copy(), I.copy())
    got = CASES[name](gpu.array(A), gpu.array(B), gpu.array(V), gpu.array(I))
    _same(got, want, name)


def test_inplace_operators_write_through_views(gpu):
    a = A.copy()
    d = gpu.array(A)
    row_h, row_d = a[2], d[2]                    # views taken BEFORE the updates
    for f in (lambda x, y: x.__iadd__(y), lambda x, y: x.__imul__(y), lambda x, y: x.__isub__(y * 0.5),
              lambda x, y: x.__itruediv__(y * y + 1.0)):
        a = f(a, B)
        d = f(d, gpu.array(B))
    a **= 2
    d **= 2
    assert np.array_equal(d.get(), a) and np.array_equal(row_d.get(), row_h)
    lazy = d + 1.0                               # a lazy expression has no storage: rebinding
    lazy += d
    assert np.array_equal(lazy.get(), (a + 1.0) + a)
    i_h, i_d = I.copy(), gpu.array(I)
    i_h += 3
    i_d += 3
    assert np.array_equal(i_d.get(), i_h) and i_d.dtype == i_h.dtype
    with pytest.raises(TypeError):               # same_kind casting, like NumPy
        i_d += 1.5
    host_out = np.zeros_like(A)
    np.add(gpu.array(A), gpu.array(B), out=host_out)
    assert np.array_equal(host_out, A + B)


def test_masked_assignment(gpu):
    a, d = A.copy(), gpu.array(A)
    a[a < 0] = 0.0
    d[d < 0] = 0.0
    assert np.array_equal(d.get(), a)
    a[a > 1] = (B * 3.0)[a > 1]
    d[d > 1] = (gpu.array(B) * 3.0)[d > 1]       # one value per selected position
    with pytest.raises(ValueError):
        d[d > 1] = gpu.array(B) * 3.0            # NumPy: cannot assign 48 values to the selected ones
    assert np.array_equal(d.get(), a)
    i_h, i_d = I.copy(), gpu.array(I)
    i_h[i_h % 2 == 0] = -1
    i_d[i_d % 2 == 0] = -1
    assert np.array_equal(i_d.get(), i_h)
    d[gpu.array(np.array([1, 2]))] = 0.0         # integer-array assignment (rows 1 and 2)
    a[np.array([1, 2])] = 0.0
    assert np.array_equal(d.get(), a)


def test_scalar_conversions_and_truthiness(gpu):
    d = gpu.array(V)
    s = np.sum(d * d)
    assert float(s) == float(np.sum(V * V)) or abs(float(s) - float(np.sum(V * V))) < 1e-12
    assert bool(np.max(d) > 0) is True and bool(np.max(d) > 1e9) is False
    steps = 0
    x = gpu.array(np.full(8, 100.0))
    while np.max(x) > 1.0:                       # a convergence loop terminates (truthiness forces)
        x = x * 0.5
        steps += 1
    assert steps == 7
    assert int(np.sum(gpu.array(I))) == int(I.sum())
    assert [int(t) for t in gpu.array(np.arange(3))] == [0, 1, 2]
    assert np.arange(5)[gpu.array(np.array(3))] == 3


def test_module_namespace_passes_numpy_names_through(gpu):
    import delayrepay as dnp
    x = dnp.array(A)
    assert np.array_equal(dnp.floor(x).get(), np.floor(A))
    assert np.array_equal(dnp.concatenate([x, x]).get(), np.concatenate([A, A]))
    assert dnp.inf == np.inf and dnp.int8 is np.int8 and dnp.e == np.e
    assert abs(float(dnp.linalg.norm(x)) - np.linalg.norm(A)) < 1e-12
    with pytest.raises(KeyError):                # no device implementation: loud, never on the host
        dnp.sort(x)
    with pytest.raises(AttributeError):
        dnp.no_such_name


def test_integer_array_and_boolean_mask_indexing(gpu):
    """a[idx], a[mask], a[idx] = v, a[mask] = values, np.take/compress/extract/nonzero/argwhere:
    gather, scan-based compaction and scatter kernels (extras.py) -- data movement, so exact."""
    rng = np.random.default_rng(12)
    for dt in (np.float64, np.float32, np.int32, np.int64, np.uint8, np.bool_):
        h = (rng.standard_normal((50, 7)) * 40).astype(dt)
        v = (rng.standard_normal(100003) * 40).astype(dt)
        d, dv = gpu.array(h), gpu.array(v)
        idx = rng.integers(-50, 50, 33)
        big = rng.integers(-100003, 100003, (4, 1000))
        cases = [(d[gpu.array(idx)], h[idx]), (d[idx], h[idx]), (d[list(idx[:5])], h[list(idx[:5])]),
                 (dv[gpu.array(big)], v[big]), (dv[dv > 3], v[v > 3]), (d[d > 3], h[h > 3]),
                 (d[gpu.array(h[:, 0] > 0)], h[h[:, 0] > 0]), (d[h[:, 1] > 0], h[h[:, 1] > 0]),
                 (dv[dv > 1e9] if dt != np.bool_ else dv[~dv & dv], v[v > 1e9] if dt != np.bool_ else v[~v & v]),
                 (np.take(d, [0, 2, -1], axis=1), np.take(h, [0, 2, -1], axis=1)),
                 (np.take(d, idx[:4]), np.take(h, idx[:4])), (np.take(dv, 7), np.take(v, 7)),
                 (np.compress(h[0] > 0, d, axis=1), np.compress(h[0] > 0, h, axis=1)),
                 (np.compress([True, False, True], d, axis=0), np.compress([True, False, True], h, axis=0)),
                 (np.extract(d > 0, d), np.extract(h > 0, h)), (np.flatnonzero(dv), np.flatnonzero(v)),
                 (np.argwhere(d > 0), np.argwhere(h > 0)), (d[(gpu.array(idx),)], h[(idx,)])]
        cases += list(zip(np.nonzero(d > 0), np.nonzero(h > 0))) + list(zip(np.where(dv > 0), np.where(v > 0)))
        for k, (got, want) in enumerate(cases):
            got = got.get()
            assert got.shape == want.shape and got.dtype == want.dtype, (dt, k, got.shape, want.shape, got.dtype)
            assert np.array_equal(got, want, equal_nan=dt in (np.float32, np.float64)), (dt, k)
        # assignment (distinct indices: NumPy's order for duplicates is "last wins", ours unspecified)
        uniq = rng.permutation(50)[:20]
        h2, d2 = h.copy(), gpu.array(h)
        h2[uniq] = 1
        d2[gpu.array(uniq)] = 1
        assert np.array_equal(d2.get(), h2)
        rows = (rng.standard_normal((20, 7)) * 9).astype(dt)
        h2[uniq] = rows
        d2[uniq] = gpu.array(rows)
        assert np.array_equal(d2.get(), h2)
        h2[uniq[:3]] = rows[0]
        d2[list(uniq[:3])] = rows[0]
        assert np.array_equal(d2.get(), h2)
        m = h2 > 0
        h2[m] = (h2 * 2)[m]
        d2[d2 > 0] = (d2 * 2)[d2 > 0]
        assert np.array_equal(d2.get(), h2)
        rm = h2[:, 0] > 0
        fill = (rng.standard_normal((int(rm.sum()), 7)) * 5).astype(dt)
        h2[rm] = fill
        d2[gpu.array(rm)] = fill
        assert np.array_equal(d2.get(), h2)
    col = np.arange(7.0).reshape(7, 1)                  # trailing unit dimensions survive a row mask
    keep = np.array([True, False, True, True, False, False, True])
    got = gpu.array(col)[gpu.array(keep)].get()
    assert got.shape == (4, 1) and np.array_equal(got, col[keep])
    assert np.array_equal(np.compress(keep[:2], gpu.array(np.ones((1, 2))), axis=1).get(),
                          np.compress(keep[:2], np.ones((1, 2)), axis=1))
    assert np.array_equal(np.compress(keep, gpu.array(col.reshape(1, 7, 1)), axis=-2).get(),
                          np.compress(keep, col.reshape(1, 7, 1), axis=-2))
    with pytest.raises(IndexError):
        gpu.array(np.arange(5.0))[gpu.array(np.array([1, 5]))]
    with pytest.raises(IndexError):
        gpu.array(np.arange(5.0))[np.array([0.5])]
    with pytest.raises(IndexError):
        gpu.array(np.arange(6.0))[gpu.array(np.array([True, False]))]
    assert gpu.array(np.arange(5.0))[gpu.array(np.zeros(0, dtype=np.int64))].get().shape == (0,)


def test_transposed_copies_use_the_tiled_kernel_and_are_exact(gpu):
    from delayrepay_b200 import engine
    rng = np.random.default_rng(21)
    for dt in (np.float32, np.float64, np.int32, np.int64):
        for shape in ((33, 65), (256, 512), (1000, 37), (3, 64, 96), (2, 3, 40, 50)):
            x = (rng.standard_normal(shape) * 100).astype(dt)
            X = gpu.array(x)
            t = np.swapaxes(X, -1, -2)
            assert np.array_equal(t.copy().get(), np.swapaxes(x, -1, -2).copy())
            assert np.array_equal(t.reshape(-1).get(), np.swapaxes(x, -1, -2).reshape(-1))
        m = (rng.standard_normal((300, 400)) * 100).astype(dt)
        M = gpu.array(m)
        assert np.array_equal(M[:, 17:250].T.copy().get(), m[:, 17:250].T.copy())      # pitch > width
        assert np.array_equal(M[5:, :].T.copy().get(), m[5:, :].T.copy())
        assert np.array_equal(M[::2, :].T.copy().get(), m[::2, :].T.copy())            # every second row: pitch
        assert np.array_equal(M[:, ::2].T.copy().get(), m[:, ::2].T.copy())            # strided columns: nd path
    assert any(k[0] == "transpose" for k in engine._kernels), "tiled transpose not taken"
