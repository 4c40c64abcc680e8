"""Streamed (out-of-core) evaluation of host-resident arrays.

`map_chunks(fn, ins, outs, chunk)` pipelines H2D copy -> fused kernel(s) -> D2H copy over chunks
of the leading axis on three CUDA streams (copy-in, compute, copy-out) with double-buffered
device staging, so that PCIe moves data in both directions while the GPU computes.  `fn` is
ordinary DelayRepay code: it receives lazy arrays (one per input, chunk-sized) and returns one
lazy array per output.  The reference's only route for host data is eager
(`delayrepay.array(host)` -> evaluate -> `.get()`, delayarray.py:101-106,620): upload, compute
and download strictly one after the other.
"""
import ctypes as C

import numpy as np

from . import engine
from ._lib import check, lib
from .delayarray import NPArray, evaluate
from .device import DeviceArray, current_device

H2D, COMPUTE, D2H = 1, 0, 2          # stream indices inside libdrcuda


def _event(dev):
    e = C.c_uint64()
    check(lib.drc_event_create(dev, C.byref(e)))
    return e.value


def map_chunks(fn, ins, outs, chunk=1 << 26, depth=2):
    """ins / outs: lists of C-contiguous host arrays sharing their leading dimension (pinned
    memory -- `pinned_empty` -- makes the copies truly asynchronous)."""
    dev = current_device()
    n = ins[0].shape[0]
    assert all(a.shape[0] == n and a.flags.c_contiguous for a in list(ins) + list(outs))
    chunk = min(chunk, n)
    stage = [[DeviceArray.empty((chunk,) + a.shape[1:], a.dtype, dev) for a in ins]
             for _ in range(depth)]
    ev_in = [_event(dev) for _ in range(depth)]       # chunk uploaded
    ev_done = [_event(dev) for _ in range(depth)]     # kernels that read the stage have run
    ev_out = [_event(dev) for _ in range(depth)]      # results downloaded
    keep = [None] * depth                             # device results alive until downloaded
    check(lib.drc_stream_sync(dev, COMPUTE))          # the staging allocations are visible
    used = [False] * depth
    for c, lo in enumerate(range(0, n, chunk)):
        hi = min(lo + chunk, n)
        b = c % depth
        if used[b]:
            check(lib.drc_stream_wait_event(dev, H2D, ev_done[b]))   # stage b no longer read
            check(lib.drc_event_sync(dev, ev_out[b]))                # its results are on the host
            keep[b] = None
        for a, d in zip(ins, stage[b]):
            part = a[lo:hi]
            check(lib.drc_memcpy_h2d_async(dev, H2D, d.ptr, part.ctypes.data, part.nbytes))
            d.buf.version += 1
        check(lib.drc_event_record(dev, H2D, ev_in[b]))
        check(lib.drc_stream_wait_event(dev, COMPUTE, ev_in[b]))
        lazy = fn(*[NPArray(d[:hi - lo]) for d in stage[b]])
        lazy = list(lazy) if isinstance(lazy, (tuple, list)) else [lazy]
        evaluate(*lazy)
        devs = [x._force() for x in lazy]
        check(lib.drc_event_record(dev, COMPUTE, ev_done[b]))
        check(lib.drc_stream_wait_event(dev, D2H, ev_done[b]))
        for o, d in zip(outs, devs):
            part = o[lo:hi]
            assert d.is_contiguous and d.nbytes == part.nbytes
            check(lib.drc_memcpy_d2h_async(dev, D2H, part.ctypes.data, d.ptr, d.nbytes))
        check(lib.drc_event_record(dev, D2H, ev_out[b]))
        keep[b] = devs
        used[b] = True
        del lazy
    for s in (H2D, COMPUTE, D2H):
        check(lib.drc_stream_sync(dev, s))
    for e in ev_in + ev_done + ev_out:
        check(lib.drc_event_destroy(dev, e))
    return outs
