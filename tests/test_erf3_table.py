"""The generation-3 erf table as shipped (delayrepay_b200/csrc/erf3.cuh), checked without a GPU:
the words are parsed from the header and the device sequence of dr_erf4_s is emulated operation by
operation in NumPy float32 (fma = float64 product-sum rounded once) against scipy's float64 erf.
The GPU counterpart is tests/test_parity_gpu.py::test_generation3_erf_in_staged_kernels."""
import os
import re

import numpy as np
from scipy.special import erf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def _table():
    text = open(os.path.join(ROOT, "delayrepay_b200", "csrc", "erf3.cuh")).read()
    rows = int(re.search(r"#define DR_ERF3_ROWS (\d+)", text).group(1))
    body = re.search(r"DR_ERF3_TAB\[\d+\] = \{(.*?)\};", text, re.S).group(1)
    words = np.array([int(w.strip().rstrip("u"), 16) for w in body.split(",")], dtype=np.uint32).reshape(rows, 4)
    mask = int(re.search(r"pk & (0x[0-9a-f]+)u", text).group(1), 16)
    shift = int(re.search(r"pk << (\d+)\)", text).group(1))
    return words, mask, shift


def _fma(a, b, c):
    return (np.float64(a) * np.float64(b) + np.float64(c)).astype(f32)


def _erf3(x, words, mask, shift, sqrt_err_ulp=0.0):
    magic = f32(98304.0)
    ap = np.minimum(np.abs(x), f32(4.0)).astype(f32)                              # FMNMX (exact, also for subnormals)
    s = (np.sqrt(ap.astype(np.float64)).astype(f32) * f32(1 + sqrt_err_ulp * 2.0 ** -23)).astype(f32)   # MUFU.SQRT
    row = (s + magic).astype(f32).view(np.int32) - magic.view(np.int32)           # FADD, LEA
    w = words[row]
    c, c0, c1 = (w[:, i].copy().view(f32) for i in range(3))
    c2 = (w[:, 3] & np.uint32(mask)).view(f32)
    c3 = ((w[:, 3].astype(np.uint64) << shift) & 0xFFFFFFFF).astype(np.uint32).view(f32)
    d = (ap - c).astype(f32)
    p = _fma(_fma(_fma(c3, d, c2), d, c1), d, c0)
    return np.copysign(p, x)


def test_shipped_erf3_table_is_within_its_error_bound():
    words, mask, shift = _table()
    assert words.shape == (257, 4) and mask == 0xFFFFF800 and shift == 21
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-6, 6, 1 << 18), 10.0 ** rng.uniform(-30, 0.7, 1 << 18),
                        10.0 ** rng.uniform(-45.5, -30, 1 << 16),              # subnormal arguments and results
                        np.linspace(0, 4.1, (1 << 18) + 1), np.linspace(0, 2.0 ** -8, 1 << 16),
                        (np.arange(0, 258) / 256.0) ** 2 * 4, ((np.arange(0, 258) + 0.5) / 256.0) ** 2 * 4]).astype(f32)
    truth = erf(x.astype(np.float64))
    ulp = np.spacing(np.abs(truth).astype(f32)).astype(np.float64)
    for sqrt_err in (0.0, 3.0, -3.0):                      # the approximate square root only picks the row
        got = _erf3(x, words, mask, shift, sqrt_err)
        err = np.abs(got.astype(np.float64) - truth) / np.maximum(ulp, 1e-300)
        err[truth == 0] = 0
        assert err.max() <= 1.5, (sqrt_err, err.max(), x[err.argmax()])
        assert err[np.abs(x) >= 2.0 ** -9].max() <= 0.70
        assert np.array_equal(np.signbit(got), np.signbit(x))
    assert _erf3(np.array([4.0, 100.0, np.inf], f32), words, mask, shift).tolist() == [1.0, 1.0, 1.0]
