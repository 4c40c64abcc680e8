#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the hot path (capture -> fuse -> launch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n L]
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline (BASELINE.json configs[1], the configuration the metric is quoted on): Black-Scholes
call/put pricing, one fused two-output kernel, float32, N = 2^30 options PER GPU (weak scaling:
the option axis is sharded across ranks, no data-path collective).  A "step" = one pass of the
hot path over the resident batch: build the lazy graph through the drop-in API, plan, launch.

The same line carries, under ``other_configs``, the other four BASELINE configs at full size
(C1 axpy, C3 fused reductions x3, C4 heat stencil, C5 n-body), each with its own roofline; with
WORLD_SIZE > 1 also the SHARDED C3 (weak scaling, per-GPU partial + ncclAllReduce through
libdrcuda's communicator) and C4 (strong scaling, row blocks, halo rows pushed into the
neighbours' memory over NVLink by the stencil kernel itself).  Everything that is timed is also
CHECKED against the oracle (oracle/refcpu.py) -- ``verified`` -- and the process exits non-zero
when a check fails.  Prints ONE JSON line (rank 0).  DESIGN.md section 5 defines every field.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_OPTION = 20          # 3 x f32 read + 2 x f32 written
NOMINAL_HBM_GBS = 8000.0       # the figure north_star's ">= 80 %" target is quoted on


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 25 ms while a timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = max((float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()),
                 default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arms
def cpu_black_scholes(log2n, reps=1):
    """The reference's CPU path (oracle port, oracle/refcpu.py) on a bounded sample."""
    from oracle import refcpu
    import workloads as wl
    n = 1 << log2n
    inp = wl.make_inputs("black_scholes", n)
    best = float("inf")
    for _ in range(reps):
        S, K, T = (refcpu.leaf(inp[k]) for k in ("S", "K", "T"))
        t0 = time.perf_counter()
        call, put = wl.black_scholes(refcpu, S, K, T)
        call.get()
        put.get()
        best = min(best, time.perf_counter() - t0)
    return n / best, best


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU path (the oracle port; the reference is pure Python
    + NumPy, there is nothing to compile into oracle/_ref) on bounded samples.  This arm imports
    NOTHING from the product package."""
    if rank != 0:
        return
    assert "delayrepay_b200" not in sys.modules
    log2n = args.cpu_log2n
    cpu_black_scholes(min(log2n, 20))
    times = []
    for _ in range(args.warmup):
        cpu_black_scholes(log2n)
    for _ in range(args.steps):
        times.append(cpu_black_scholes(log2n)[1])
    n = 1 << log2n
    total = sum(times)
    value = n * args.steps / total
    assert "delayrepay_b200" not in sys.modules
    line = {
        "impl": "reference", "metric": "fused elems/s (Black-Scholes options/s)",
        "value": value, "unit": "options/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "black_scholes_f32_call_put", "log2_options_per_step": log2n,
                   "note": "bounded sample of the 2^30 workload on host cores"},
        "cpu_baseline": {"value": value, "unit": "options/s", "cores": 1, "kind": "port",
                         "sample": f"2^{log2n} options/step, oracle/refcpu.py (unfused NumPy per "
                                   f"node, tree-recursive like reference cpu.py:13-31); "
                                   f"host has {os.cpu_count()} cores, NumPy elementwise uses 1"},
        "e2e": {"value": value, "unit": "options/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ device timing
class DevTimer:
    """CUDA events on libdrcuda's launch stream (torch.cuda.Event would only see torch's)."""

    def __init__(self, lib, check, dev):
        self.lib, self.check, self.dev = lib, check, dev

    def event(self):
        e = C.c_uint64()
        self.check(self.lib.drc_event_create(self.dev, C.byref(e)))
        return e.value

    def record(self, e):
        self.check(self.lib.drc_event_record(self.dev, 0, e))

    def elapsed(self, a, b):
        ms = C.c_float()
        self.check(self.lib.drc_event_sync(self.dev, b))
        self.check(self.lib.drc_event_elapsed_ms(self.dev, a, b, C.byref(ms)))
        return float(ms.value)

    def timed(self, fn, reps, warm, sync):
        """Mean device time of `fn` over `reps` calls after `warm` untimed ones."""
        for _ in range(warm):
            fn()
        a, b = self.event(), self.event()
        sync()
        self.record(a)
        for _ in range(reps):
            fn()
        self.record(b)
        return self.elapsed(a, b) / reps


def _roof(bound, achieved, peak, unit, **extra):
    d = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak}
    d.update(extra)
    return d


def _hbm_roof(bytes_per_launch, ms, peak, **extra):
    gbs = bytes_per_launch / (ms * 1e-3) / 1e9
    return _roof("hbm", gbs, peak, "GB/s", frac_of_nominal_8000=gbs / NOMINAL_HBM_GBS,
                 algorithmic_bytes_per_launch=bytes_per_launch, kernel_ms=ms, traffic=None, **extra)


# ------------------------------------------------------------------------------ verification
class Verifier:
    """Everything bench.py times is checked against the oracle (oracle/refcpu.py = the
    reference's CPU path) under the parity suite's bars.  Device arrays at the BASELINE sizes are
    tiled from a seeded host chunk, so the oracle only ever evaluates the chunk."""

    def __init__(self):
        self.results, self.failed = {}, []

    def put(self, name, ok, **info):
        self.results[name] = dict(ok=bool(ok), **info)
        if not ok:
            self.failed.append(name)


def verify_black_scholes(ver, dr, wl, host, call, put, n, chunk, name="black_scholes_f32"):
    """call/put over ALL n positions: (i) every chunk-sized block of the outputs is bitwise equal
    to block 0 (device-side compare: the inputs repeat with period `chunk`), (ii) block 0 and a
    strided sample spanning the first and last GiB equal the oracle within the parity bar
    (tests/test_parity_gpu.py: |d| <= 16 eps32 max(S, K))."""
    from oracle import refcpu
    rc, rp = wl.black_scholes(refcpu, *(refcpu.leaf(host[k]) for k in ("S", "K", "T")))
    rc, rp = rc.get(), rp.get()
    bar = 16 * np.finfo(np.float32).eps * np.maximum(host["S"], host["K"])
    worst, periodic = 0.0, True
    idx = np.unique(np.concatenate([np.arange(0, 1 << 15), np.arange(n - (1 << 15), n),
                                    (np.arange(0, n, max(n >> 16, 1)) + 7) % n]))
    for got, want in ((call, rc), (put, rp)):
        head = got[:chunk].get()
        err = np.abs(head.astype(np.float64) - want)
        worst = max(worst, float(np.max(err / bar)))
        if n > chunk:
            blocks = got.reshape(n // chunk, chunk)
            periodic &= bool(np.all(np.equal(blocks, blocks[0:1])))
        sample = got[dr.array(idx)].get()
        worst = max(worst, float(np.max(np.abs(sample.astype(np.float64) - want[idx % chunk]) / bar[idx % chunk])))
    # how slack is the bar?  errors of the engine and of the oracle itself against a float64
    # evaluation of the same float32 inputs, in ulps of the operand scale max(S, K) (call and put
    # are differences of terms of that size: near-zero prices have no meaningful ulp of their own)
    t64 = wl.black_scholes(np, *(host[k].astype(np.float64) for k in ("S", "K", "T")))
    scale = np.spacing(np.maximum(host["S"], host["K"])).astype(np.float64)
    ours = max(float(np.max(np.abs(g[:chunk].get() - t) / scale)) for g, t in ((call, t64[0]), (put, t64[1])))
    theirs = max(float(np.max(np.abs(w - t) / scale)) for w, t in ((rc, t64[0]), (rp, t64[1])))
    ver.put(name, worst <= 1.0 and periodic, max_err_over_bar=worst,
            bar="16*eps32*max(S,K) = 16 ulp of the operand scale", all_blocks_bitwise_equal_block0=periodic,
            positions_checked=int(n), sampled_positions=int(idx.size),
            max_err_vs_float64_truth_in_ulp_of_max_S_K={"engine": ours, "oracle_numpy": theirs})


def heat_oracle_block(wl, u0_fn, r0, r1, c0, c1, steps, g):
    """Oracle value of grid rows [r0, r1) x cols [c0, c1) after `steps` Jacobi steps: run the
    reference path on the block grown by `steps` cells (clipped at the true boundary, which
    stays fixed); influence travels one cell per step, so the core is exact."""
    from oracle import refcpu
    R0, R1, C0, C1 = max(r0 - steps, 0), min(r1 + steps, g), max(c0 - steps, 0), min(c1 + steps, g)
    blk = refcpu.leaf(u0_fn(R0, R1, C0, C1).copy())
    wl.heat(refcpu, blk, steps)
    return blk.get()[r0 - R0:r1 - R0, c0 - C0:c1 - C0]


# ------------------------------------------------------------------------------ other configs
def bench_others(dr, wl, tm, ver, dev, peak, sync, quick=False):
    """BASELINE.json configs 1, 3, 4, 5 at full size on ONE GPU: device-resident inputs, CUDA
    events on the launch stream, each checked against the oracle."""
    from oracle import refcpu
    out = {}

    # ---- C1 axpy f64 N=2^24 (24 B/elem), whole result bit-exact vs the oracle
    n = 1 << 24
    i = wl.make_inputs("axpy", n)
    x, y = dr.array(i["x"]), dr.array(i["y"])
    ms = tm.timed(lambda: wl.axpy(dr, i["a"], x, y).run(), 50, 5, sync)
    got = wl.axpy(dr, i["a"], x, y).get()
    want = wl.axpy(refcpu, i["a"], refcpu.leaf(i["x"]), refcpu.leaf(i["y"])).get()
    ver.put("axpy_f64_2^24", got.tobytes() == want.tobytes(), bar="bit-exact, all 2^24 elements")
    out["axpy_f64_2^24"] = {"ms": ms, "value": n / (ms * 1e-3), "unit": "elems/s",
                            "roofline": _hbm_roof(24 * n, ms, peak),
                            "note": "per step through the drop-in API (capture + plan-cache replay + launch)"}
    del x, y

    # ---- C3 fused reductions f64 N=2^30: 2^22 seeded chunk tiled 256x on the device
    n, chunk = (1 << 30, 1 << 22) if not quick else (1 << 24, 1 << 22)
    i = wl.make_inputs("l2", chunk)
    a = dr.tile(dr.array(i["a"]), n // chunk)
    b = dr.tile(dr.array(i["b"]), n // chunk)
    ra, rb = refcpu.leaf(i["a"]), refcpu.leaf(i["b"])
    reps_of_chunk = n // chunk
    # the device arrays repeat the chunk: sum over n = reps x sum over the chunk (rtol 1e-12 bar)
    def val(x):
        return float(x.get()) if hasattr(x, "get") else float(x)
    want = {"l2_distance": float(np.sqrt(reps_of_chunk)) * val(wl.l2_distance(refcpu, ra, rb)),
            "dot": reps_of_chunk * val(wl.dot(refcpu, ra, rb)),
            "norm": float(np.sqrt(reps_of_chunk)) * val(wl.norm(refcpu, ra))}
    for name, fn, bpe in (("l2_distance", lambda: wl.l2_distance(dr, a, b), 16),
                          ("dot", lambda: wl.dot(dr, a, b), 16), ("norm", lambda: wl.norm(dr, a), 8)):
        ms = tm.timed(lambda: fn().run(), 10, 2, sync)
        got = float(fn())
        rel = abs(got - want[name]) / abs(want[name])
        key = f"{name}_f64_2^{n.bit_length() - 1}"
        ver.put(key, rel <= 1e-12, rel_err=rel, bar="rtol 1e-12 vs oracle on the tiled chunk")
        out[key] = {"ms": ms, "value": n / (ms * 1e-3), "unit": "elems/s",
                    "roofline": _hbm_roof(bpe * n, ms, peak)}
    del a, b

    # ---- C4 heat 32768^2 f32, 100 steps (8 B/cell/step): 2048^2 seeded block tiled 16 x 16
    g, blk = (32768, 2048) if not quick else (4096, 2048)
    steps = 100
    h0 = wl.make_inputs("heat", blk)["u"]
    u = dr.tile(dr.array(h0), (g // blk, g // blk))

    def u0(r0, r1, c0, c1):
        return h0[np.ix_(np.arange(r0, r1) % blk, np.arange(c0, c1) % blk)]
    wl.heat(dr, u, steps)                                   # warm-up = the verified run
    worst = True
    spots = [(0, 96, 0, 96), (g - 96, g, g - 96, g), (g // 2 - 48, g // 2 + 48, g // 4 - 48, g // 4 + 48)]
    for (r0, r1, c0, c1) in spots:
        got = u[r0:r1, c0:c1].get()
        worst &= got.tobytes() == heat_oracle_block(wl, u0, r0, r1, c0, c1, steps, g).tobytes()
    ver.put("heat_f32_x100", worst, bar="bit-exact after 100 steps on 3 blocks of 96x96 "
            "(both corners incl. the fixed boundary and the last bytes of the 4 GiB grid, and the centre)")
    sampler = ClockSampler(dev)
    sampler.start()
    ms = tm.timed(lambda: wl.heat(dr, u, steps), 1, 0, sync)
    clk = sampler.stop()
    ms20 = tm.timed(lambda: wl.heat(dr, u, 20), 1, 0, sync)
    key = f"heat_f32_{g}^2_x100"
    out[key] = {"ms": ms / steps, "value": g * g / (ms / steps * 1e-3), "unit": "cell-steps/s",
                "roofline": _hbm_roof(8 * g * g, ms / steps, peak), "clocks": clk,
                "ms_burst_20_steps": ms20 / 20,
                "note": "100 back-to-back steps = the sustained (power-capped) regime"}
    del u

    # ---- C5 n-body N=65536: all-pairs producer fused into the contraction
    nb = 65536 if not quick else 8192
    i = wl.make_inputs("nbody", nb)
    pos, m = dr.array(i["pos"]), dr.array(i["m"])
    ms = tm.timed(lambda: wl.nbody_acc(dr, pos, m).run(), 5, 2, sync)
    acc = wl.nbody_acc(dr, pos, m).get()
    rows = np.arange(0, nb, nb // 64)
    p, mm = i["pos"], i["m"]
    d = p[None, :, :] - p[rows, None, :]                              # (64, N, 3) float32
    r2 = d[..., 0] ** 2 + d[..., 1] ** 2 + d[..., 2] ** 2 + np.float32(1e-3)
    w = mm[None, :] * r2 ** np.float32(-1.5)
    want = (w.astype(np.float64) @ p.astype(np.float64)) - p[rows].astype(np.float64) * w.astype(np.float64).sum(1)[:, None]
    scale = (np.abs(w).astype(np.float64) @ np.abs(p).astype(np.float64)) + np.abs(p[rows]) * np.abs(w).sum(1)[:, None]
    rel = float(np.max(np.abs(acc[rows] - want) / scale))
    ver.put("nbody_f32", rel <= 1e-5, err_over_term_scale=rel, bar="1e-5 of sum|w||pos| (64 sampled rows, float64 truth)")
    sm_ghz = 1.965
    fp32_peak_pairs = 148 * 128 * sm_ghz * 1e9 / 16.0                 # 16 FP32-pipe ops per pair
    key = f"nbody_f32_{nb}"
    out[key] = {"ms": ms, "value": nb * nb / (ms * 1e-3), "unit": "pairs/s",
                "roofline": _roof("fp32", nb * nb / (ms * 1e-3) / 1e12, fp32_peak_pairs / 1e12, "Tpair/s",
                                  note="FP32-pipe bound at 16 ops/pair (DESIGN.md section 4); not HBM")}
    return out


# ------------------------------------------------------------------------------ sharded configs
def bench_sharded(dr, wl, tm, ver, dev, peak, sync, rank, world, max_over_ranks, quick=False):
    """WORLD_SIZE > 1: the collective-exercising configs through the sharding layer
    (delayrepay_b200/shard.py).  C3: weak scaling, 2^30 float64 per GPU, per-GPU partial +
    ncclAllReduce via libdrcuda's communicator.  C4: strong scaling, the 32768^2 grid split into
    row blocks, halo rows pushed into the neighbours' memory by the stencil kernel."""
    from oracle import refcpu
    out = {}
    mesh = dr.sharding.init()
    # ---- C3
    n, chunk = (1 << 30, 1 << 22) if not quick else (1 << 24, 1 << 22)
    loc = wl.make_inputs("l2", chunk, seed=3 + rank)
    a = dr.sharding.from_local(dr.tile(dr.array(loc["a"]), n // chunk))
    b = dr.sharding.from_local(dr.tile(dr.array(loc["b"]), n // chunk))
    ms = max_over_ranks(tm.timed(lambda: wl.l2_distance(dr, a, b).run(), 10, 2, sync))
    got = float(wl.l2_distance(dr, a, b))
    tot = 0.0
    for r in range(world):
        ir = wl.make_inputs("l2", chunk, seed=3 + r)
        tot += (n // chunk) * float(np.sum((ir["a"] - ir["b"]) ** 2))
    rel = abs(got - np.sqrt(tot)) / np.sqrt(tot)
    ver.put("sharded_l2_distance", rel <= 1e-12, rel_err=rel, bar="rtol 1e-12 vs oracle over all ranks' chunks")
    out["l2_distance_f64_sharded"] = {
        "scaling": "weak", "elems_per_gpu": n, "ms": ms, "value": world * n / (ms * 1e-3), "unit": "elems/s",
        "collective": "ncclAllReduce(sum, 1 x f64) on libdrcuda's communicator",
        "roofline": _hbm_roof(16 * n, ms, peak, note="per GPU")}
    del a, b
    # ---- C2 through the layer: S, K, T sharded (2^30 options per GPU, weak), call and put co-evaluated
    # per row block by sharding.run_many -- one two-output kernel per block, no communication
    nopt, chunk = (1 << 30, 1 << 22) if not quick else (1 << 24, 1 << 22)
    hb = wl.make_inputs("black_scholes", chunk, seed=2 + rank)
    S, K, T = (dr.sharding.from_local(dr.tile(dr.array(hb[k]), nopt // chunk)) for k in ("S", "K", "T"))

    def bs_step():
        call, put = wl.black_scholes(dr, S, K, T)
        dr.evaluate(call, put)
        return call, put
    ms = max_over_ranks(tm.timed(lambda: bs_step(), 10, 3, sync))
    call, put = bs_step()
    mine = [dr.NPArray(x._force().base.blocks[rank]) for x in (call, put)]
    verify_black_scholes(ver, dr, wl, hb, mine[0], mine[1], nopt, chunk, name="sharded_black_scholes")
    out["black_scholes_f32_sharded"] = {
        "scaling": "weak", "options_per_gpu": nopt, "ms": ms, "value": world * nopt / (ms * 1e-3), "unit": "options/s",
        "collective": "none (elementwise regions are localised per row block)",
        "roofline": _hbm_roof(BYTES_PER_OPTION * nopt, ms, peak, note="per GPU")}
    del S, K, T, call, put, mine
    # ---- C4
    g, blk = (32768, 2048) if not quick else (4096, 2048)
    steps = 100
    h0 = wl.make_inputs("heat", blk)["u"]
    lo, hi = mesh.bounds(g)

    def u0(r0, r1, c0, c1):
        return h0[np.ix_(np.arange(r0, r1) % blk, np.arange(c0, c1) % blk)]
    u = dr.sharding.from_global_fn(lambda r0, r1: dr.tile(dr.array(h0), ((r1 - r0) // blk + 2, g // blk))[
        (r0 % blk):(r0 % blk) + (r1 - r0)], (g, g), np.float32)
    assert u.array.base.H == 1
    wl.heat(dr, u, steps)                                   # warm-up = the verified run
    ok = True
    for (r0, r1, c0, c1) in [(lo, min(lo + 64, hi), 0, 96), (max(hi - 64, lo), hi, g - 96, g)]:
        got = u.array.local_rows(r0, r1)[:, c0:c1].get()
        ok &= got.tobytes() == heat_oracle_block(wl, u0, r0, r1, c0, c1, steps, g).tobytes()
    ver.put("sharded_heat_x100", ok, bar="bit-exact after 100 steps on the first and last 64 rows of this "
            "rank's block (they depend on the neighbours' rows through 100 halo exchanges)")
    sampler = ClockSampler(dev)
    sampler.start()
    ms = max_over_ranks(tm.timed(lambda: wl.heat(dr, u, steps), 1, 0, sync))
    clk = sampler.stop()
    out["heat_f32_sharded_x100"] = {
        "scaling": "strong", "grid": [g, g], "rows_per_gpu": hi - lo, "ms": ms / steps,
        "value": g * g / (ms / steps * 1e-3), "unit": "cell-steps/s", "clocks": clk,
        "exchange": "one halo row per neighbour per step, stored into the neighbour's block by the "
                    "stencil kernel over NVLink peer mappings + release/acquire flags (no NCCL, no host sync)",
        "roofline": _hbm_roof(8 * g * (hi - lo), ms / steps, peak, note="per GPU, this rank's rows")}
    del u
    # ---- C5: rows of W sharded, pos / m gathered once per step (strong scaling)
    nb = 65536 if not quick else 8192
    i = wl.make_inputs("nbody", nb)
    pos, m = dr.shard(i["pos"], halo=0), dr.array(i["m"])
    ms = max_over_ranks(tm.timed(lambda: wl.nbody_acc(dr, pos, m).run(), 5, 2, sync))
    acc = wl.nbody_acc(dr, pos, m)._force()
    lo, hi = acc.base.bounds[rank]
    rows = np.arange(lo, hi, max((hi - lo) // 16, 1))
    got = acc.local_rows(lo, hi).get()[rows - lo]
    p, mm = i["pos"], i["m"]
    d = p[None, :, :] - p[rows, None, :]
    r2 = d[..., 0] ** 2 + d[..., 1] ** 2 + d[..., 2] ** 2 + np.float32(1e-3)
    w = (mm[None, :] * r2 ** np.float32(-1.5)).astype(np.float64)
    want = w @ p.astype(np.float64) - p[rows].astype(np.float64) * w.sum(1)[:, None]
    scale = np.abs(w) @ np.abs(p).astype(np.float64) + np.abs(p[rows]) * np.abs(w).sum(1)[:, None]
    rel = float(np.max(np.abs(got - want) / scale))
    ver.put("sharded_nbody", rel <= 1e-5, err_over_term_scale=rel, bar="1e-5 of sum|w||pos| (16 rows of this rank's block, float64 truth)")
    out["nbody_f32_sharded"] = {
        "scaling": "strong", "bodies": nb, "rows_per_gpu": hi - lo, "ms": ms, "value": nb * nb / (ms * 1e-3),
        "unit": "pairs/s", "exchange": "pos columns and the right operand all-gathered per step "
                                       "(drc_nccl_allgather, 1 MiB); rows of W never leave their GPU"}
    return out


# ------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=30, help="options per GPU = 2^log2n")
    ap.add_argument("--cpu-log2n", type=int, default=24, help="CPU baseline sample size")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-log2chunk", type=int, default=25)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip configs 1, 3, 4, 5")
    ap.add_argument("--others", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--quick", action="store_true", help="reduced sizes for the other configs (smoke)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine
    import workloads as wl
    from delayrepay_b200._lib import lib, check

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dr.set_device(local_rank)
    dev = local_rank
    # host side of the e2e path: threads and pinned staging buffers on the GPU's NUMA node
    from delayrepay_b200.device import bind_to_device_numa
    numa_cpus = None if os.environ.get("DR_NO_NUMA_BIND") else bind_to_device_numa(local_rank)
    tm = DevTimer(lib, check, dev)
    ver = Verifier()

    def barrier():
        dr.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident synthetic inputs: a seeded 2^22 host chunk per rank, tiled on the device
    n = 1 << args.log2n
    chunk = min(n, 1 << 22)
    host = wl.make_inputs("black_scholes", chunk, seed=2 + rank)
    S, K, T = (dr.tile(dr.array(host[k]), n // chunk) if n > chunk else dr.array(host[k])
               for k in ("S", "K", "T"))
    barrier()

    def step():
        call, put = wl.black_scholes(dr, S, K, T)
        dr.evaluate(call, put)
        return call, put

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = lib.drc_launch_count()
    marks = [tm.event() for _ in range(args.steps + 1)]
    barrier()
    tm.record(marks[0])
    t_host0 = time.perf_counter()
    last = None
    for i in range(args.steps):
        # drop the previous step's results BEFORE capturing again: while they are alive the
        # memo table would hand back the already evaluated nodes and nothing would launch
        last = None
        last = step()
        tm.record(marks[i + 1])
    barrier()
    host_s = time.perf_counter() - t_host0
    total_ms = tm.elapsed(marks[0], marks[-1])
    per_step = [tm.elapsed(marks[i], marks[i + 1]) for i in range(args.steps)]
    launches = lib.drc_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(total_ms)
    kernel_ms = float(np.mean(per_step))          # one kernel per step: step time == launch time
    value = world * n * args.steps / (total_ms * 1e-3)
    bs_kernel = engine.last_kernel_name()
    # ---- what was timed is checked: the outputs of the LAST timed step, all n positions
    verify_black_scholes(ver, dr, wl, host, last[0], last[1], n, chunk)
    del last

    # ---- e2e: host buffers in, host buffers out, through the public API
    e2e = None
    if not args.no_e2e:
        import psutil
        n_e = n
        while n_e * BYTES_PER_OPTION * world * 2 > psutil.virtual_memory().available and n_e > (1 << 20):
            n_e >>= 1
        hin = [dr.pinned_empty(n_e, np.float32) for _ in range(3)]
        hout = [dr.pinned_empty(n_e, np.float32) for _ in range(2)]
        for h, k in zip(hin, ("S", "K", "T")):
            h[:] = np.resize(host[k], n_e)

        def e2e_step():
            s, k, t = (dr.array(h) for h in hin)
            call, put = wl.black_scholes(dr, s, k, t)
            dr.evaluate(call, put)
            call.get(out=hout[0])
            put.get(out=hout[1])

        def e2e_streamed():
            dr.map_chunks(lambda s, k, t: wl.black_scholes(dr, s, k, t), hin, hout,
                          chunk=1 << args.e2e_log2chunk)
        del S, K, T
        results = {}
        for name, fn in (("eager", e2e_step), ("streamed", e2e_streamed)):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                fn()
            barrier()
            results[name] = max_over_ranks(time.perf_counter() - t0)
        # the host result buffers of the streamed run against block 0 of the oracle values
        from oracle import refcpu
        rc, _rp = wl.black_scholes(refcpu, *(refcpu.leaf(host[k]) for k in ("S", "K", "T")))
        bar = 16 * np.finfo(np.float32).eps * np.maximum(host["S"], host["K"])
        m = min(n_e, chunk)
        e_err = float(np.max(np.abs(hout[0][-m:].astype(np.float64) - np.resize(rc.get(), n_e)[-m:]) / np.resize(bar, n_e)[-m:]))
        ver.put("e2e_host_buffers", e_err <= 1.0, max_err_over_bar=e_err, bar="16*eps32*max(S,K), last block of the host call buffer")
        dt = results["streamed"]
        e2e = {"value": world * n_e * args.e2e_steps / dt, "unit": "options/s",
               "h2d_bytes_per_step": 12 * n_e, "d2h_bytes_per_step": 8 * n_e,
               "options_per_step_per_gpu": n_e, "steps": args.e2e_steps,
               "ms_per_step": 1e3 * dt / args.e2e_steps,
               "eager_ms_per_step": 1e3 * results["eager"] / args.e2e_steps,
               "note": "pinned host buffers -> dr.map_chunks(black_scholes): chunked H2D / fused "
                       "kernel / D2H on three streams, every byte crosses PCIe inside the timed "
                       "region; wall clock, max over ranks.  eager_ms_per_step = dr.array(h) -> "
                       "evaluate -> .get(out=), copies and kernel strictly serial",
               "host_numa_cpus": None if numa_cpus is None else len(numa_cpus)}
        hin = hout = None
    else:
        del S, K, T

    peak, peak_src = _peaks()
    others = sharded = None
    if not args.no_others:
        if world > 1:
            sharded = bench_sharded(dr, wl, tm, ver, dev, peak, barrier, rank, world, max_over_ranks, args.quick)
        if rank == 0:
            others = bench_others(dr, wl, tm, ver, dev, peak, dr.synchronize, args.quick)
    failed = ver.failed
    if world > 1:
        flag = torch.tensor([len(failed)], dtype=torch.float64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
        any_failed = flag.item() > 0
        barrier()
    else:
        any_failed = bool(failed)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if failed:
            print(f"rank {rank}: verification FAILED: {failed} {ver.results}", file=sys.stderr, flush=True)
            sys.exit(3)
        return

    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.log2n == 30:
        with open(tpath) as f:
            ent = json.load(f).get("black_scholes_f32_call_put", {})
        if ent.get("kernel") == bs_kernel:
            traffic = ent.get("dram_bytes_per_launch")
            traffic_note = f"ncu --set full of kernel {bs_kernel}: {ent.get('source')}"
        else:
            traffic_note = (f"profiles/traffic.json was captured for kernel {ent.get('kernel')}, this run "
                            f"launched {bs_kernel}: stale entry refused")
    achieved = BYTES_PER_OPTION * n / (kernel_ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu:
        v, secs = cpu_black_scholes(args.cpu_log2n)
        cpu = {"value": v, "unit": "options/s", "cores": 1, "kind": "port",
               "sample": f"2^{args.cpu_log2n} options once ({secs:.1f} s), oracle/refcpu.py = "
                         f"reference cpu.py path (unfused NumPy, single-threaded); host has "
                         f"{os.cpu_count()} cores"}
    line = {
        "metric": "fused elems/s (Black-Scholes options/s)", "value": value,
        "unit": "options/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "black_scholes_f32_call_put", "options_per_gpu": n,
                   "global_options": n * world, "kernel": "one fused two-output flat kernel",
                   "cache_hygiene": "inputs 12 GiB + outputs 8 GiB per GPU >> 126 MB L2",
                   "parallelism": f"option axis sharded over {world} rank(s), no collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "frac_of_nominal_8000": achieved / NOMINAL_HBM_GBS,
                     "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_OPTION * n,
                     "kernel_ms": kernel_ms, "kernel": bs_kernel},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "host_wall_s": host_s,
        "verified": {"ok": not any_failed, "checks": ver.results},
        "other_configs": others, "sharded_configs": sharded,
        "engine": {k: (round(v, 1) if isinstance(v, float) else v) for k, v in engine.stats.items()},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if any_failed:
        print(f"verification FAILED: {failed} (all ranks: see stderr)", file=sys.stderr, flush=True)
        sys.exit(3)


if __name__ == "__main__":
    main()
