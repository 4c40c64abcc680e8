"""Pipe microbenchmarks on the B200 through libdrcuda (NVRTC + launch + events):
issue cost of FFMA vs FFMA2 (packed f32x2), DFMA rate vs number of independent chains,
F2F conversion rate.  usage: python tools/microbench.py"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delayrepay_b200 as dr
from delayrepay_b200 import engine
from delayrepay_b200._lib import lib, check

SRC = r'''
#define ITERS 4096
extern "C" __global__ void k_ffma(float* out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" __global__ void k_ffma2(float* out, float a, float b) {
  float2 x[4];
  for (int i = 0; i < 4; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0; for (int i = 0; i < 4; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> __device__ void dfma_body(float* out, double a, double b) {
  double x[CH];
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS * 8 / CH / 4; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0; for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
extern "C" __global__ void k_dfma1(float* out, double a, double b) { dfma_body<1>(out, a, b); }
extern "C" __global__ void k_dfma2(float* out, double a, double b) { dfma_body<2>(out, a, b); }
extern "C" __global__ void k_dfma4(float* out, double a, double b) { dfma_body<4>(out, a, b); }
extern "C" __global__ void k_dfma8(float* out, double a, double b) { dfma_body<8>(out, a, b); }
// mixed: 4 DFMA chains + 8 FFMA (or 4 FFMA2) per iteration: does packed f32 free issue slots?
extern "C" __global__ void k_mix_scalar(float* out, double a, double b, float c, float d) {
  double x[4]; float y[16];
  for (int i = 0; i < 4; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 16; ++i) y[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = fmaf(y[i], c, d);
  }
  double s = 0; for (int i = 0; i < 4; ++i) s += x[i];
  for (int i = 0; i < 16; ++i) s += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
extern "C" __global__ void k_mix_packed(float* out, double a, double b, float c, float d) {
  double x[4]; float2 y[8];
  for (int i = 0; i < 4; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 8; ++i) y[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  const float2 cc = make_float2(c, c), dd = make_float2(d, d);
  for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = __ffma2_rn(y[i], cc, dd);
  }
  double s = 0; for (int i = 0; i < 4; ++i) s += x[i];
  for (int i = 0; i < 8; ++i) s += y[i].x + y[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
__constant__ double HC[12] = {1e-7, 2e-6, -3e-5, 2e-4, -8e-4, 2e-3, -3e-4, -2e-2, 0.1, 0.6, 1.1, 0.5};
// Horner chains with constant-bank coefficients, LANES chains interleaved, plus per-trip
// conversion f32->f64 at the start and f64->f32 at the end (the shape of dr_erf4_fast)
template <int LANES> __device__ void horner_body(float* out, float seed) {
  float x[LANES];
  for (int l = 0; l < LANES; ++l) x[l] = seed + threadIdx.x * 1e-4f + l * 0.01f;
  for (int it = 0; it < 256; ++it) {
    double a[LANES], p[LANES];
#pragma unroll
    for (int l = 0; l < LANES; ++l) { a[l] = (double)x[l]; p[l] = HC[0]; }
#pragma unroll
    for (int i = 1; i < 12; ++i)
#pragma unroll
      for (int l = 0; l < LANES; ++l)
        asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(p[l]) : "d"(p[l]), "d"(a[l]), "d"(HC[i]));
#pragma unroll
    for (int l = 0; l < LANES; ++l) x[l] = (float)p[l];
  }
  float s = 0; for (int l = 0; l < LANES; ++l) s += x[l];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" __global__ void k_horner1(float* out, float s) { horner_body<1>(out, s); }
extern "C" __global__ void k_horner2(float* out, float s) { horner_body<2>(out, s); }
extern "C" __global__ void k_horner4(float* out, float s) { horner_body<4>(out, s); }
extern "C" __global__ void k_f2f(float* out, float a) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { double d = (double)x[i]; d += 1.0; x[i] = (float)d; }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
'''


def main():
    dr.set_device(0)
    src, cubin = engine.compile_source("microbench", SRC)
    dev = 0
    mod = C.c_uint64()
    check(lib.drc_module_load(dev, cubin, len(cubin), C.byref(mod)))
    blocks, threads = 148 * 8, 256
    out = dr.DeviceArray.empty((blocks * threads,), "f4")
    ev = [C.c_uint64(), C.c_uint64()]
    for e in ev:
        check(lib.drc_event_create(dev, C.byref(e)))

    def run(name, args, ops_per_thread, label):
        fn = C.c_uint64()
        check(lib.drc_module_get_function(dev, mod, name.encode(), C.byref(fn)))
        k = engine.Kernel(name, "", cubin, {})
        k.funcs[dev] = fn.value
        best = 1e9
        for _ in range(3):
            check(lib.drc_event_record(dev, 0, ev[0].value))
            engine.launch(k, dev, blocks, threads, args)
            check(lib.drc_event_record(dev, 0, ev[1].value))
            check(lib.drc_event_sync(dev, ev[1].value))
            ms = C.c_float()
            check(lib.drc_event_elapsed_ms(dev, ev[0].value, ev[1].value, C.byref(ms)))
            best = min(best, ms.value)
        total = ops_per_thread * blocks * threads
        clk = 1.9e9
        per_sm_clk = total / (best * 1e-3) / 148 / clk
        print(f"{label:34s} {best:8.3f} ms  {per_sm_clk:7.1f} thread-ops/clk/SM (at 1.9 GHz)")

    def a_f(*vals):
        a = engine.Args(); a.ptr(out.ptr)
        for v, t in vals:
            a.scalar(v, t)
        return a
    IT = 4096
    run("k_ffma", a_f((1.0001, "f4"), (1e-4, "f4")), IT * 8, "FFMA x8 chains")
    run("k_ffma2", a_f((1.0001, "f4"), (1e-4, "f4")), IT * 8, "FFMA2 x4 chains (8 lanes)")
    for ch in (1, 2, 4, 8):
        run(f"k_dfma{ch}", a_f((1.0001, "f8"), (1e-4, "f8")), IT * 8, f"DFMA {ch} chain(s)/thread")
    run("k_mix_scalar", a_f((1.0001, "f8"), (1e-4, "f8"), (1.0001, "f4"), (1e-4, "f4")), IT // 2 * 20, "4 DFMA + 16 FFMA  (ops=20/it)")
    run("k_mix_packed", a_f((1.0001, "f8"), (1e-4, "f8"), (1.0001, "f4"), (1e-4, "f4")), IT // 2 * 20, "4 DFMA + 8 FFMA2  (ops=20/it)")
    run("k_f2f", a_f((1.0, "f4"),), IT // 4 * 8 * 2, "F2F f32->f64->f32 (conv ops)")
    # latency / occupancy study: limit resident warps with a small grid
    def run_occ(name, args, ops_per_thread, label, blocks_per_sm, thr):
        fn = C.c_uint64()
        check(lib.drc_module_get_function(dev, mod, name.encode(), C.byref(fn)))
        k = engine.Kernel(name, "", cubin, {})
        k.funcs[dev] = fn.value
        best = 1e9
        for _ in range(3):
            check(lib.drc_event_record(dev, 0, ev[0].value))
            engine.launch(k, dev, 148 * blocks_per_sm, thr, args)
            check(lib.drc_event_record(dev, 0, ev[1].value))
            check(lib.drc_event_sync(dev, ev[1].value))
            ms = C.c_float()
            check(lib.drc_event_elapsed_ms(dev, ev[0].value, ev[1].value, C.byref(ms)))
            best = min(best, ms.value)
        total = ops_per_thread * 148 * blocks_per_sm * thr
        print(f"{label:44s} {best:8.3f} ms  {total / (best * 1e-3) / 148 / 1.9e9:7.1f} ops/clk/SM")
    for warps_per_sm in (4, 8, 16, 24, 32):
        for ch in (1, 4):
            run_occ(f"k_dfma{ch}", a_f((1.0001, "f8"), (1e-4, "f8")), IT * 8,
                    f"DFMA {ch} chain(s), {warps_per_sm} warps/SM", 1, warps_per_sm * 32)
    for warps_per_sm in (8, 16, 24, 32):
        for lanes in (1, 2, 4):
            run_occ(f"k_horner{lanes}", a_f((0.5, "f4"),), 256 * 11 * lanes,
                    f"Horner(11 DFMA)+2 F2F, {lanes} lane(s), {warps_per_sm} warps/SM", 1, warps_per_sm * 32)
    for warps_per_sm in (8, 16, 24, 32):
        run_occ("k_mix_packed", a_f((1.0001, "f8"), (1e-4, "f8"), (1.0001, "f4"), (1e-4, "f4")),
                IT // 2 * 20, f"4 DFMA + 8 FFMA2, {warps_per_sm} warps/SM", 1, warps_per_sm * 32)


if __name__ == "__main__":
    main()
