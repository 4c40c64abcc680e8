"""ctypes binding of libdrcuda.so (include/drcuda.h).

The library is the only way this package touches a GPU.  It must exist (built in-tree by
``__graft_entry__.build()`` / ``make -C delayrepay_b200/csrc``); a missing library or a
missing device raises -- there is no CPU fallback behind this module.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrcuda.so")


class DrcError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise DrcError(
            f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). "
            "delayrepay_b200 has no CPU fallback.")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

u64, u32, i32, sz, vp = C.c_uint64, C.c_uint32, C.c_int, C.c_size_t, C.c_void_p
P = C.POINTER

_SIGS = {
    "drc_abi_version": ([], i32),
    "drc_last_error": ([], C.c_char_p),
    "drc_init": ([], i32),
    "drc_shutdown": ([], i32),
    "drc_device_count": ([P(i32)], i32),
    "drc_device_attr": ([i32, P(i32), P(i32), P(i32), P(sz), P(i32), P(i32)], i32),
    "drc_device_name": ([i32, C.c_char_p, sz], i32),
    "drc_device_pci_bus_id": ([i32, C.c_char_p, sz], i32),
    "drc_mem_info": ([i32, P(sz), P(sz)], i32),
    "drc_malloc_async": ([i32, i32, sz, P(u64)], i32),
    "drc_free_async": ([i32, i32, u64], i32),
    "drc_pool_trim": ([i32, sz], i32),
    "drc_memset_async": ([i32, i32, u64, i32, sz], i32),
    "drc_memcpy_h2d_async": ([i32, i32, u64, vp, sz], i32),
    "drc_memcpy_d2h_async": ([i32, i32, vp, u64, sz], i32),
    "drc_memcpy_d2d_async": ([i32, i32, u64, u64, sz], i32),
    "drc_memcpy_peer_async": ([i32, u64, i32, u64, sz, i32, i32], i32),
    "drc_enable_peer_access": ([i32, i32], i32),
    "drc_peer_alloc": ([i32, sz, P(u64)], i32),
    "drc_peer_free": ([i32, u64], i32),
    "drc_ipc_get_handle": ([i32, u64, vp], i32),
    "drc_ipc_open_handle": ([i32, vp, P(u64)], i32),
    "drc_ipc_close_handle": ([i32, u64], i32),
    "drc_host_alloc": ([sz, P(vp)], i32),
    "drc_host_free": ([vp], i32),
    "drc_host_register": ([vp, sz], i32),
    "drc_host_unregister": ([vp], i32),
    "drc_compile": ([C.c_char_p, C.c_char_p, P(C.c_char_p), i32, P(vp), P(sz), P(vp)], i32),
    "drc_free_blob": ([vp], i32),
    "drc_nvrtc_version": ([P(i32), P(i32), P(C.c_char_p)], i32),
    "drc_module_load": ([i32, vp, sz, P(u64)], i32),
    "drc_module_unload": ([i32, u64], i32),
    "drc_module_get_function": ([i32, u64, C.c_char_p, P(u64)], i32),
    "drc_func_set_max_dynamic_smem": ([i32, u64, i32], i32),
    "drc_func_attrs": ([i32, u64, P(i32), P(i32), P(i32), P(i32)], i32),
    "drc_occupancy": ([i32, u64, i32, sz, P(i32)], i32),
    "drc_launch": ([i32, i32, u64, P(u32), P(u32), u32, P(vp), i32], i32),
    "drc_launch_packed": ([i32, i32, u64, u32, u32, u32, u32, u32, u32, u32, u32, vp, vp, i32], i32),
    "drc_launch_count": ([], u64),
    "drc_tensormap_encode": ([i32, vp, i32, u32, u64, P(u64), P(u64), P(u32), i32, i32], i32),
    "drc_stream_sync": ([i32, i32], i32),
    "drc_device_sync": ([i32], i32),
    "drc_event_create": ([i32, P(u64)], i32),
    "drc_event_destroy": ([i32, u64], i32),
    "drc_event_record": ([i32, i32, u64], i32),
    "drc_event_sync": ([i32, u64], i32),
    "drc_event_elapsed_ms": ([i32, u64, u64, P(C.c_float)], i32),
    "drc_stream_wait_event": ([i32, i32, u64], i32),
    "drc_stream_handle": ([i32, i32, P(u64)], i32),
    "drc_nccl_available": ([], i32),
    "drc_nccl_get_unique_id": ([vp], i32),
    "drc_nccl_init_rank": ([i32, i32, i32, vp, P(u64)], i32),
    "drc_nccl_init_all": ([i32, P(i32), P(u64)], i32),
    "drc_nccl_destroy": ([u64], i32),
    "drc_nccl_allreduce": ([u64, i32, i32, u64, u64, sz, i32, i32], i32),
    "drc_nccl_sendrecv": ([u64, i32, i32, u64, sz, i32, u64, sz, i32], i32),
    "drc_nccl_allgather": ([u64, i32, i32, u64, u64, sz], i32),
    "drc_nccl_group_start": ([], i32),
    "drc_nccl_group_end": ([], i32),
    "drc_fft_c2c_1d": ([i32, i32, u64, u64, i32, i32, i32, i32], i32),
}
EXPORTS = tuple(_SIGS)

for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch
    _fn.argtypes = _args
    _fn.restype = _res


def check(rc):
    if rc != 0:
        raise DrcError(lib.drc_last_error().decode("utf-8", "replace"))


_state = {"up": False, "ndev": 0}


def init():
    """Bring the driver up (idempotent).  Raises DrcError when no GPU is usable."""
    if not _state["up"]:
        check(lib.drc_init())
        n = i32()
        check(lib.drc_device_count(C.byref(n)))
        _state["ndev"] = n.value
        _state["up"] = True
    return _state["ndev"]


def gpu_available():
    try:
        return init() > 0
    except DrcError:
        return False


def nvrtc_version():
    """(major, minor, path) of the NVRTC libdrcuda bound."""
    ma, mi, path = C.c_int(), C.c_int(), C.c_char_p()
    check(lib.drc_nvrtc_version(C.byref(ma), C.byref(mi), C.byref(path)))
    return ma.value, mi.value, (path.value or b"").decode()


def compile_cubin(source, name, options):
    """NVRTC: CUDA C++ text -> sm_100a cubin bytes.  Works without a GPU."""
    opts = (C.c_char_p * len(options))(*[o.encode() for o in options])
    blob, n, log = vp(), sz(), vp()
    rc = lib.drc_compile(source.encode(), name.encode(), opts, len(options),
                         C.byref(blob), C.byref(n), C.byref(log))
    log_text = ""
    if log.value:
        log_text = C.string_at(log.value).decode("utf-8", "replace")
        lib.drc_free_blob(log)
    if rc != 0:
        raise DrcError(lib.drc_last_error().decode("utf-8", "replace"))
    cubin = C.string_at(blob.value, n.value)
    lib.drc_free_blob(blob)
    return cubin, log_text
