"""PCIe ceilings of the box and a chunk/depth sweep of dr.map_chunks on Black-Scholes.

usage: python tools/e2e_sweep.py [log2n]     (default 2^30 options, like bench.py's e2e leg)
Prints one JSON line per measurement: raw pinned H2D, D2H and both at once through libdrcuda's
copy streams (what the e2e path can reach at best), then ms per step of map_chunks for each
(chunk, depth).
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import delayrepay_b200 as dr
import workloads as wl
from delayrepay_b200._lib import check, lib
from delayrepay_b200.device import DeviceArray
from delayrepay_b200.stream import H2D, D2H


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    n = 1 << log2n
    dr.set_device(0)
    dev = 0
    hin = [dr.pinned_empty(n, np.float32) for _ in range(3)]
    hout = [dr.pinned_empty(n, np.float32) for _ in range(2)]
    rng = np.random.default_rng(2)
    blk = 1 << 24
    blk = min(blk, n)
    src = [rng.uniform(5, 30, blk).astype(np.float32), rng.uniform(1, 100, blk).astype(np.float32),
           rng.uniform(0.25, 10, blk).astype(np.float32)]
    for lo in range(0, n, blk):                     # one random block, repeated (timing only)
        for h, s_ in zip(hin, src):
            h[lo:lo + blk] = s_
    d_in = [DeviceArray.empty((n,), np.float32, dev) for _ in range(3)]
    d_out = [DeviceArray.empty((n,), np.float32, dev) for _ in range(2)]

    def sync():
        for s in (0, H2D, D2H):
            check(lib.drc_stream_sync(dev, s))

    def raw(do_in, do_out, piece=1 << 28):
        sync()
        t0 = time.perf_counter()
        for lo in range(0, n * 4, piece):
            if do_in:
                for h, d in zip(hin, d_in):
                    check(lib.drc_memcpy_h2d_async(dev, H2D, d.ptr + lo, h.ctypes.data + lo,
                                                   min(piece, n * 4 - lo)))
            if do_out:
                for h, d in zip(hout, d_out):
                    check(lib.drc_memcpy_d2h_async(dev, D2H, h.ctypes.data + lo, d.ptr + lo,
                                                   min(piece, n * 4 - lo)))
        sync()
        return time.perf_counter() - t0

    for name, a, b in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
        raw(a, b)
        dt = min(raw(a, b) for _ in range(2))
        gb_in = 12 * n / 1e9 if a else 0
        gb_out = 8 * n / 1e9 if b else 0
        print(json.dumps({"raw": name, "ms": round(dt * 1e3, 2),
                          "h2d_GBs": round(gb_in / dt, 1), "d2h_GBs": round(gb_out / dt, 1)}),
              flush=True)

    fn = lambda s, k, t: wl.black_scholes(dr, s, k, t)
    for log2c in (22, 24, 25, 26):
        for depth in (2, 3):
            kw = {"chunk": 1 << log2c, "depth": depth}
            dr.map_chunks(fn, hin, hout, **kw)
            best = 1e9
            for _ in range(2):
                sync()
                t0 = time.perf_counter()
                dr.map_chunks(fn, hin, hout, **kw)
                best = min(best, time.perf_counter() - t0)
            print(json.dumps({"map_chunks": kw, "ms": round(best * 1e3, 2)}), flush=True)


if __name__ == "__main__":
    main()
