#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the hot path (capture -> fuse -> launch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n L]
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): Black-Scholes
call/put pricing, one fused two-output kernel, float32, N = 2^30 options PER GPU (weak scaling:
the option axis is sharded across ranks, no data-path collective).  A "step" = one pass of the
hot path over the resident batch: build the lazy graph through the drop-in API, plan, launch.
Prints ONE JSON line (rank 0).  See DESIGN.md section 5 for every field's definition.
"""
import argparse
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


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""

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
    if rank != 0:
        return
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


# ------------------------------------------------------------------------------ other configs
def bench_others(dr, wl, lib, check, dev, peak, reps=10):
    """BASELINE.json configs 1, 3, 4, 5 at full size: device-resident inputs, CUDA events on the
    launch stream, mean over `reps` after 2 warm-ups.  Reported for context next to the headline
    (they are parity-test cases, not the bench line)."""
    import ctypes as C
    out = {}

    def timed(fn, n_rep=reps, warm=2):
        for _ in range(warm):
            fn()
        a, b = C.c_uint64(), C.c_uint64()
        check(lib.drc_event_create(dev, C.byref(a)))
        check(lib.drc_event_create(dev, C.byref(b)))
        dr.synchronize()
        check(lib.drc_event_record(dev, 0, a.value))
        for _ in range(n_rep):
            fn()
        check(lib.drc_event_record(dev, 0, b.value))
        check(lib.drc_event_sync(dev, b.value))
        ms = C.c_float()
        check(lib.drc_event_elapsed_ms(dev, a.value, b.value, C.byref(ms)))
        return ms.value / n_rep

    def entry(name, ms, units, unit, bytes_per_unit):
        gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "value": units / (ms * 1e-3), "unit": unit, "achieved_GBs": gbs,
                     "frac_of_measured_hbm": gbs / peak, "bytes_per_unit": bytes_per_unit}

    # C1 axpy f64 N=2^24
    n = 1 << 24
    i = wl.make_inputs("axpy", n)
    x, y = dr.array(i["x"]), dr.array(i["y"])
    entry("axpy_f64_2^24", timed(lambda: wl.axpy(dr, i["a"], x, y).run(), 50), n, "elems/s", 24)
    del x, y
    # C3 fused reductions f64 N=2^30 (seeded 2^22 chunk tiled on the device)
    n = 1 << 30
    i = wl.make_inputs("l2", 1 << 22)
    a = dr.tile(dr.array(i["a"]), n >> 22)
    b = dr.tile(dr.array(i["b"]), n >> 22)
    entry("l2_distance_f64_2^30", timed(lambda: wl.l2_distance(dr, a, b).run()), n, "elems/s", 16)
    entry("dot_f64_2^30", timed(lambda: wl.dot(dr, a, b).run()), n, "elems/s", 16)
    entry("norm_f64_2^30", timed(lambda: wl.norm(dr, a).run()), n, "elems/s", 8)
    del a, b
    # C4 heat 32768^2 f32, 100 steps
    g = 32768
    u = dr.tile(dr.array(wl.make_inputs("heat", 2048)["u"]), (g // 2048, g // 2048))
    steps = 100
    sampler = ClockSampler(dev)
    sampler.start()
    ms = timed(lambda: wl.heat(dr, u, steps), 1, 1)
    clk = sampler.stop()
    entry("heat_f32_32768^2_x100", ms / steps, g * g, "cell-steps/s", 8)
    out["heat_f32_32768^2_x100"]["clocks"] = clk      # 100 steps: the sustained (power-capped) regime
    ms20 = timed(lambda: wl.heat(dr, u, 20), 1, 0)
    out["heat_f32_32768^2_x100"]["ms_burst_20_steps"] = ms20 / 20
    del u
    # C5 n-body N=65536 (all-pairs producer fused into the contraction)
    nb = 65536
    i = wl.make_inputs("nbody", nb)
    pos, m = dr.array(i["pos"]), dr.array(i["m"])
    ms = timed(lambda: wl.nbody_acc(dr, pos, m).run(), 5, 2)
    out["nbody_f32_65536"] = {"ms": ms, "value": nb * nb / (ms * 1e-3), "unit": "pairs/s"}
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
    ap.add_argument("--others", action="store_true", help="also time configs 1, 3, 4, 5 (N=1)")
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
    import ctypes as C

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dr.set_device(local_rank)
    dev = local_rank
    # host side of the e2e path: threads and pinned staging buffers on the GPU's NUMA node
    from delayrepay_b200.device import bind_to_device_numa
    numa_cpus = None if os.environ.get("DR_NO_NUMA_BIND") else bind_to_device_numa(local_rank)

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

    def event():
        e = C.c_uint64()
        check(lib.drc_event_create(dev, C.byref(e)))
        return e.value

    def record(e):
        check(lib.drc_event_record(dev, 0, e))

    def elapsed(a, b):
        ms = C.c_float()
        check(lib.drc_event_sync(dev, b))
        check(lib.drc_event_elapsed_ms(dev, a, b, C.byref(ms)))
        return float(ms.value)

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
    marks = [event() for _ in range(args.steps + 1)]
    barrier()
    record(marks[0])
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        step()
        record(marks[i + 1])
    barrier()
    host_s = time.perf_counter() - t_host0
    total_ms = elapsed(marks[0], marks[-1])
    per_step = [elapsed(marks[i], marks[i + 1]) for i in range(args.steps)]
    launches = lib.drc_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(total_ms)
    kernel_ms = float(np.mean(per_step))          # one kernel per step: step time == launch time
    value = world * n * args.steps / (total_ms * 1e-3)

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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    others = None
    if args.others and world == 1:
        hin = hout = None
        others = bench_others(dr, wl, lib, check, dev, peak)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.log2n == 30:
        with open(tpath) as f:
            traffic = json.load(f).get("black_scholes_f32_call_put", {}).get("dram_bytes_per_launch")
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
                     "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_OPTION * n,
                     "kernel_ms": kernel_ms},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "host_wall_s": host_s, "other_configs": others,
        "engine": {k: (round(v, 1) if isinstance(v, float) else v) for k, v in engine.stats.items()},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
