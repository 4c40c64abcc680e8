"""The reference's own acceptance tests (reference tests/test.py:14-162), re-stated against the
drop-in alias ``import delayrepay`` so they read like the originals; same constants, same
assertions (the three dot tests, vacuous upstream, assert here)."""
import numpy as np
import numpy.testing as npt
import pytest

pytestmark = pytest.mark.gpu
SIZE = 64


@pytest.fixture()
def mod(gpu):
    import delayrepay
    return delayrepay


def test_elementwise(mod):                                      # TestElwise, test.py:14-67
    arr, np_arr = mod.ones(SIZE), np.ones(SIZE)
    npt.assert_array_almost_equal((arr + 1).get(), np_arr + 1)
    npt.assert_array_almost_equal((arr * 3).get(), np_arr * 3)
    npt.assert_array_almost_equal((7 * arr).get(), np_arr * 7)
    npt.assert_array_almost_equal((8 * arr + 9).get(), 8 * np_arr + 9)
    npt.assert_array_almost_equal((arr + arr * 3 + 9).get(), np_arr + np_arr * 3 + 9)
    assert arr + 3
    npt.assert_array_almost_equal(np.cos(arr).get(), np.cos(np_arr))
    npt.assert_array_almost_equal((arr ** 2).get(), np_arr ** 2)
    arr32 = mod.ones(SIZE).astype(np.float32)
    npt.assert_array_almost_equal((arr32 ** 2).get(), np_arr.astype(np.float32) ** 2)
    res = np.sin(arr) ** 2 + np.cos(arr) ** 2
    npt.assert_array_almost_equal(res.get(), np.sin(np_arr) ** 2 + np.cos(np_arr) ** 2)


def test_vector(mod):                                           # TestVector, test.py:70-111
    arr = mod.full((SIZE,), 7).astype(np.float32)
    arr2 = mod.full((SIZE,), 3).astype(np.float32)
    np_arr = np.full((SIZE,), 7).astype(np.float32)
    np_arr2 = np.full((SIZE,), 3).astype(np.float32)
    npt.assert_array_almost_equal((arr + arr2).get(), np_arr + np_arr2)
    npt.assert_array_almost_equal((arr * arr2).get(), np_arr * np_arr2)
    assert abs(float(arr.dot(arr2)) - np_arr.dot(np_arr2)) < 0.001
    assert abs(float(np.dot(arr, arr2)) - np_arr.dot(np_arr2)) < 0.001
    assert abs(float(arr @ arr2) - np_arr @ np_arr2) < 0.001
    assert mod.sum(arr) == np.sum(np_arr)
    npt.assert_array_almost_equal(np.arctan2(arr, arr2).get(), np.arctan2(np_arr, np_arr2))


def test_matrix(mod):                                           # TestMatrix, test.py:114-140
    mat = mod.full((SIZE, SIZE), 7).astype(np.float32)
    vec = mod.full((SIZE,), 3).astype(np.float32)
    np_mat = np.full((SIZE, SIZE), 7).astype(np.float32)
    np_vec = np.full((SIZE,), 3).astype(np.float32)
    npt.assert_array_almost_equal((mat * 3).get(), np_mat * 3)
    npt.assert_array_almost_equal((mat @ vec).get(), np_mat @ np_vec)
    a = mod.full((64, 64), 10.0, dtype=np.float32)
    b = mod.full((64,), 2.0, dtype=np.float32)
    npt.assert_array_almost_equal((a @ b).get(), np.full((64, 64), 10.0, np.float32) @ np.full((64,), 2.0, np.float32))
    npt.assert_array_almost_equal((mat @ mat).get(), np_mat @ np_mat)


def test_meta(mod):                                             # TestMeta, test.py:143-162
    arr = np.array([1, 2, 3])
    assert mod.NPArray(arr) is mod.NPArray(arr)
    assert mod.full((3,), 5).astype(np.float32) is not mod.full((3,), 3).astype(np.float32)
    x = mod.array([1, 2, 3])
    assert np.sin(x) is np.sin(x)
