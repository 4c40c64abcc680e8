"""delayrepay.fft entry point  (reference fft.py:9-12).

The reference's ``fft(self, *args)`` carries a stray ``self`` parameter and is broken on its CPU
backend; the signature here is NumPy's.  The transform runs in cuFFT
behind the C ABI (``drc_fft_c2c_1d``); real input is packed to complex on the device first, so no
element ever visits the host.  Like NumPy 2, float32/complex64 input gives complex64 output.
"""
import numpy as np

from . import extras
from ._lib import check, lib
from .delayarray import NPArray, arg_to_numpy_ex
from .device import DeviceArray


def _transform(a, n, axis, norm, inverse):
    node = arg_to_numpy_ex(a) if not isinstance(a, DeviceArray) else NPArray(a)
    x = node._force()
    if x.ndim == 0:
        raise ValueError("fft needs at least one dimension")
    axis %= x.ndim
    if axis != x.ndim - 1:
        perm = [i for i in range(x.ndim) if i != axis] + [axis]
        x = x.transpose(*perm)
    x = x if x.is_contiguous else x.copy()
    n_in = x.shape[-1]
    n_out = int(n) if n is not None else n_in
    if n_out < 1:
        raise ValueError(f"Invalid number of FFT data points ({n_out}) specified.")
    single = x.dtype in (np.dtype(np.float32), np.dtype(np.complex64), np.dtype(np.float16))
    cdt = np.dtype(np.complex64 if single else np.complex128)
    src = x if x.dtype.kind in "fc" else x.astype(np.float64)
    lead = x.shape[:-1]
    rows = 1
    for s in lead:
        rows *= s
    out = extras.pack_complex(src.reshape(rows, n_in), n_out, cdt)
    if rows and out.dev >= 0:
        check(lib.drc_fft_c2c_1d(out.dev, 0, out.ptr, out.ptr, n_out, rows, 0 if single else 1,
                                 1 if inverse else 0))
    scale = None
    if norm == "ortho":
        scale = 1.0 / np.sqrt(n_out)
    elif (norm == "forward" and not inverse) or (norm in (None, "backward") and inverse):
        scale = 1.0 / n_out
    if scale is not None:
        part = np.dtype(np.float32 if single else np.float64)
        flat = DeviceArray(out.buf, (rows, 2 * n_out), part, None, out.offset)
        flat[...] = NPArray(flat) * part.type(scale)
    res = out.reshape(lead + (n_out,))
    if axis != x.ndim - 1:
        inv = list(range(x.ndim - 1))
        inv.insert(axis, x.ndim - 1)
        res = res.transpose(*inv)
    return NPArray(res)


def fft(a, n=None, axis=-1, norm=None):
    return _transform(a, n, axis, norm, False)


def ifft(a, n=None, axis=-1, norm=None):
    return _transform(a, n, axis, norm, True)
