"""Issue-slot / pipe microbenchmarks for the instruction mix of the fused Black-Scholes kernel:
FFMA vs FFMA2 rate, FFMA2 mixed with ALU-pipe integer ops, LDS.64 vs LDS.128 with divergent
table indices, MUFU.  Numbers are warp-instructions per clock per SM sub-partition (SMSP) at the
SM clock sampled from the event time (assumes 1.965 GHz boost; the ratio between rows is what
matters).   usage: python tools/microbench2.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delayrepay_b200 as dr
from delayrepay_b200 import engine
from delayrepay_b200._lib import lib, check

SRC = r'''
#define ITERS 2048
typedef unsigned long long p2;
__device__ __forceinline__ p2 pk(float a, float b) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(p2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x + y; }
#define FMA2(x, a, b) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(a), "l"(b))
#define FMA1(x, a, b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b))
#define LOP(x, a) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(x) : "r"(a))
#define IADD(x, a) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(a))
#define FMNMX(x, a) asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(a))
#define MUFU(x) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x))

extern "C" __global__ void k_ffma(float* out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) FMA1(x[i], a, b);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" __global__ void k_ffma2(float* out, float a, float b) {
  p2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = pk(threadIdx.x * 1e-3f + i, i);
  const p2 aa = pk(a, a), bb = pk(b, b);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) FMA2(x[i], aa, bb);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += lo(x[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// NF FFMA2 + NA ALU ops per group, 16 instructions per trip in total
template <int NF, int NA> __device__ void mix_body(float* out, float a, float b, int c) {
  p2 x[8]; int y[8];
  for (int i = 0; i < 8; ++i) { x[i] = pk(threadIdx.x * 1e-3f + i, i); y[i] = threadIdx.x + i; }
  const p2 aa = pk(a, a), bb = pk(b, b);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int g = 0; g < 16 / (NF + NA); ++g) {
#pragma unroll
      for (int i = 0; i < NF; ++i) FMA2(x[(g * NF + i) & 7], aa, bb);
#pragma unroll
      for (int i = 0; i < NA; ++i) LOP(y[(g * NA + i) & 7], y[(g * NA + i + 3) & 7]);
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += lo(x[i]) + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" __global__ void k_mix_3_1(float* out, float a, float b, int c) { mix_body<3, 1>(out, a, b, c); }
extern "C" __global__ void k_mix_1_1(float* out, float a, float b, int c) { mix_body<1, 1>(out, a, b, c); }
extern "C" __global__ void k_mix_1_3(float* out, float a, float b, int c) { mix_body<1, 3>(out, a, b, c); }
extern "C" __global__ void k_alu(float* out, float a, float b, int c) {
  int y[8];
  for (int i = 0; i < 8; ++i) y[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) LOP(y[i], y[(i + 3) & 7]);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" __global__ void k_alu_mix(float* out, float a, float b, int c) {   // LOP3 + IADD + FMNMX
  int y[8]; float z[8];
  for (int i = 0; i < 8; ++i) { y[i] = threadIdx.x + i; z[i] = threadIdx.x * 0.5f + i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) LOP(y[i], y[i + 4]);
#pragma unroll
    for (int i = 4; i < 8; ++i) IADD(y[i], y[i - 4]);
#pragma unroll
    for (int i = 0; i < 8; ++i) FMNMX(z[i], a);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += y[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" __global__ void k_mufu(float* out, float a, float b, int c) {
  float z[8];
  for (int i = 0; i < 8; ++i) z[i] = threadIdx.x * 0.5f + i + 1.0f;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) MUFU(z[i]);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// table look-ups with divergent indices: 4 x LDS.64 (stride 40 rows) vs 2 x LDS.128 per element
extern "C" __global__ void k_lds64(float* out, float a, float b, int c) {
  __shared__ float2 tab[160];
  for (int i = threadIdx.x; i < 160; i += blockDim.x) tab[i] = make_float2(i, -i);
  __syncthreads();
  unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
  float s = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      h = h * 1664525u + 1013904223u;
      const float2* row = tab + (h >> 27) + 4;           // 32 of the 40 rows
      const float2 t0 = row[0], t1 = row[40], t2 = row[80], t3 = row[120];
      s += t0.x + t1.y + t2.x + t3.y;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" __global__ void k_lds128(float* out, float a, float b, int c) {
  __shared__ float4 tab[80];
  for (int i = threadIdx.x; i < 80; i += blockDim.x) tab[i] = make_float4(i, -i, 1, 2);
  __syncthreads();
  unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
  float s = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      h = h * 1664525u + 1013904223u;
      const float4* row = tab + (h >> 27) + 4;
      const float4 t0 = row[0], t1 = row[40];
      s += t0.x + t0.w + t1.y + t1.z;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the same index stream without the loads (cost of the surrounding arithmetic)
extern "C" __global__ void k_lds0(float* out, float a, float b, int c) {
  unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
  float s = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      h = h * 1664525u + 1013904223u;
      s += (float)(h >> 27) + a; s += b; s += a; s += b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
'''


def main():
    dr.set_device(0)
    src, cubin = engine.compile_source("microbench2", SRC)
    dev = 0
    mod = C.c_uint64()
    check(lib.drc_module_load(dev, cubin, len(cubin), C.byref(mod)))
    out = dr.DeviceArray.empty((148 * 8 * 256,), "f4")
    ev = [C.c_uint64(), C.c_uint64()]
    for e in ev:
        check(lib.drc_event_create(dev, C.byref(e)))

    def run(name, instr_per_thread, label, blocks_per_sm=8, threads=256):
        fn = C.c_uint64()
        check(lib.drc_module_get_function(dev, mod, name.encode(), C.byref(fn)))
        k = engine.Kernel(name, "", cubin, {})
        k.funcs[dev] = fn.value
        a = engine.Args()
        a.ptr(out.ptr); a.scalar(1.0001, "f4"); a.scalar(1e-4, "f4"); a.scalar(12345, "i4")
        best = 1e9
        for _ in range(4):
            check(lib.drc_event_record(dev, 0, ev[0].value))
            engine.launch(k, dev, 148 * blocks_per_sm, threads, a)
            check(lib.drc_event_record(dev, 0, ev[1].value))
            check(lib.drc_event_sync(dev, ev[1].value))
            ms = C.c_float()
            check(lib.drc_event_elapsed_ms(dev, ev[0].value, ev[1].value, C.byref(ms)))
            best = min(best, ms.value)
        warps = 148 * blocks_per_sm * threads / 32
        wi = instr_per_thread * warps
        per_smsp_clk = wi / (best * 1e-3) / (148 * 4) / 1.965e9
        print(f"{label:46s} {best:8.3f} ms  {per_smsp_clk:6.3f} warp-instr/clk/SMSP")

    IT = 2048
    for bps in (8, 4, 2):
        print(f"-- {bps} blocks x 256 threads per SM")
        run("k_ffma", IT * 16, "FFMA, 8 chains", bps)
        run("k_ffma2", IT * 16, "FFMA2, 8 chains", bps)
        run("k_alu", IT * 16, "SHF, 8 chains", bps)
        run("k_alu_mix", IT * 12, "4 SHF + 4 IADD(->IMAD) + 8 FMNMX(->4 FMNMX3)", bps)
        run("k_mix_3_1", IT * 16, "3 FFMA2 : 1 SHF", bps)
        run("k_mix_1_1", IT * 16, "1 FFMA2 : 1 SHF", bps)
        run("k_mix_1_3", IT * 16, "1 FFMA2 : 3 SHF", bps)
        run("k_mufu", IT // 4 * 16, "MUFU.RCP, 8 chains", bps)
        run("k_lds0", IT * 4, "index stream only (per look-up)", bps)
        run("k_lds64", IT * 4, "look-up: 4 x LDS.64 (per look-up)", bps)
        run("k_lds128", IT * 4, "look-up: 2 x LDS.128 (per look-up)", bps)


if __name__ == "__main__":
    main()
