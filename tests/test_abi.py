"""The C-ABI library loads without a GPU and exports every symbol include/drcuda.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "drcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "delayrepay_b200", "libdrcuda.so"))
    names = declared_symbols()
    assert len(names) >= 45
    for name in names:
        assert hasattr(lib, name), f"libdrcuda.so does not export {name}"


def test_binding_covers_header():
    from delayrepay_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    assert _lib.lib.drc_abi_version() == 1


def test_device_calls_fail_loudly_without_init_or_gpu():
    from delayrepay_b200 import _lib
    if _lib.gpu_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.DrcError):
        _lib.init()
    n = ctypes.c_uint64()
    assert _lib.lib.drc_malloc_async(0, 0, 16, ctypes.byref(n)) != 0
    assert b"drc_init" in _lib.lib.drc_last_error() or b"libcuda" in _lib.lib.drc_last_error()


def test_nvrtc_cross_compiles_sm100a_without_gpu():
    from delayrepay_b200 import _lib
    src = 'extern "C" __global__ void k(float* p) { p[threadIdx.x] *= 2.0f; }'
    cubin, _ = _lib.compile_cubin(src, "k.cu", ["--gpu-architecture=sm_100a"])
    assert cubin[:4] == b"\x7fELF"
    with pytest.raises(_lib.DrcError):
        _lib.compile_cubin("this is not CUDA", "bad.cu", ["--gpu-architecture=sm_100a"])
