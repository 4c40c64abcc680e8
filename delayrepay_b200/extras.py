"""Device kernels beside the fused-region generator: prefix sums (np.cumsum), the Philox counter
RNG behind delayrepay.random, and the real->complex packing used by fft.  Each is a hand-written
CUDA C++ template specialised by dtype and compiled through the same NVRTC/cubin cache
(engine.get_kernel).  Reference: these are the eager CuPy calls of delayarray.py:555-558 (cumsum),
random.py:8-13 (cuRAND) and fft.py:12 (cuFFT).
"""
import os

import numpy as np

from . import engine
from .codegen import ctype
from .device import DeviceArray, current_device
from .engine import Args, get_kernel, launch

# ------------------------------------------------------------------------------ cumsum
_SCAN_SRC = r'''
// scan along the middle axis of (outer, n, inner): one thread per (outer, inner) pair, serial in
// n, coalesced along inner
extern "C" __global__ void __launch_bounds__(256) NAME_serial(const TIN* __restrict__ in,
    TACC* __restrict__ out, i64 outer, i64 n, i64 inner) {
  const i64 total = outer * inner;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 o = idx / inner, c = idx - o * inner;
    const TIN* p = in + o * n * inner + c;
    TACC* q = out + o * n * inner + c;
    TACC acc = (TACC)0;
    for (i64 k = 0; k < n; ++k) { acc += (TACC)p[k * inner]; q[k * inner] = acc; }
  }
}
// rows of a matrix, scanned along the contiguous axis: one block per row walks it in tiles of
// 256 * ITEMS elements (thread-local scan, warp shuffles, running carry): one read, one write
#define ITEMS 8
extern "C" __global__ void __launch_bounds__(256) NAME_rowscan(const TIN* __restrict__ in,
    TACC* __restrict__ out, i64 rows, i64 n) {
  __shared__ TACC warp_tot[8];
  for (i64 r = blockIdx.x; r < rows; r += gridDim.x) {
    const TIN* p = in + r * n;
    TACC* q = out + r * n;
    TACC carry = (TACC)0;
    for (i64 base = 0; base < n; base += 256 * ITEMS) {
      const i64 b = base + (i64)threadIdx.x * ITEMS;
      TACC v[ITEMS], s = (TACC)0;
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) { v[k] = b + k < n ? (TACC)p[b + k] : (TACC)0; s += v[k]; v[k] = s; }
      TACC x = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { TACC y = dr_shfl_up(x, o); if ((threadIdx.x & 31) >= o) x += y; }
      if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
      __syncthreads();
      TACC before = carry + (x - s), tile = (TACC)0;
#pragma unroll
      for (int w = 0; w < 8; ++w) { if (w < (threadIdx.x >> 5)) before += warp_tot[w]; tile += warp_tot[w]; }
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) if (b + k < n) q[b + k] = before + v[k];
      carry += tile;
      __syncthreads();
    }
  }
}
// scan along the middle axis with the axis cut into `parts` chunks (enough threads when
// outer * inner alone is small): (1) chunk totals, (2) exclusive scan of the totals over the
// chunks, (3) rescan every chunk from its offset.  Coalesced along inner throughout.
extern "C" __global__ void __launch_bounds__(256) NAME_chunk_totals(const TIN* __restrict__ in,
    TACC* __restrict__ totals, i64 outer, i64 n, i64 inner, i64 chunk, i64 parts) {
  const i64 total = outer * parts * inner;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 c = idx % inner, s = (idx / inner) % parts, o = idx / (inner * parts);
    const i64 lo = s * chunk, hi = lo + chunk < n ? lo + chunk : n;
    const TIN* p = in + o * n * inner + c;
    TACC acc = (TACC)0;
#pragma unroll 4
    for (i64 k = lo; k < hi; ++k) acc += (TACC)p[k * inner];
    totals[idx] = acc;
  }
}
extern "C" __global__ void __launch_bounds__(256) NAME_chunk_offsets(TACC* __restrict__ totals,
    i64 outer, i64 inner, i64 parts) {
  const i64 total = outer * inner;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 o = idx / inner, c = idx - o * inner;
    TACC acc = (TACC)0;
    for (i64 s = 0; s < parts; ++s) {
      TACC* t = totals + (o * parts + s) * inner + c;
      const TACC v = *t; *t = acc; acc += v;
    }
  }
}
extern "C" __global__ void __launch_bounds__(256) NAME_chunk_final(const TIN* __restrict__ in,
    TACC* __restrict__ out, const TACC* __restrict__ totals, i64 outer, i64 n, i64 inner, i64 chunk, i64 parts) {
  const i64 total = outer * parts * inner;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 c = idx % inner, s = (idx / inner) % parts, o = idx / (inner * parts);
    const i64 lo = s * chunk, hi = lo + chunk < n ? lo + chunk : n;
    const TIN* p = in + o * n * inner + c;
    TACC* q = out + o * n * inner + c;
    TACC acc = totals[idx];
#pragma unroll 4
    for (i64 k = lo; k < hi; ++k) { acc += (TACC)p[k * inner]; q[k * inner] = acc; }
  }
}
// 1-d, three phases: (1) per-block totals, (2) exclusive scan of the totals by one block,
// (3) per-block inclusive scan (warp shuffles + shared memory) offset by its prefix
extern "C" __global__ void __launch_bounds__(256) NAME_partials(const TIN* __restrict__ in,
    TACC* __restrict__ totals, i64 n) {
  __shared__ TACC scratch[32];
  const i64 base = (i64)blockIdx.x * 256 * ITEMS;
  TACC s = (TACC)0;
  for (int k = 0; k < ITEMS; ++k) {
    const i64 i = base + (i64)k * 256 + threadIdx.x;
    if (i < n) s += (TACC)in[i];
  }
  s = dr_block_reduce<DrSum>(s, (TACC)0, scratch);
  if (threadIdx.x == 0) totals[blockIdx.x] = s;
}
extern "C" __global__ void __launch_bounds__(1024) NAME_offsets(TACC* __restrict__ totals, int nblocks) {
  __shared__ TACC warp_tot[32];
  __shared__ TACC carry;
  if (threadIdx.x == 0) carry = (TACC)0;
  __syncthreads();
  for (int start = 0; start < nblocks; start += 1024) {
    const int i = start + threadIdx.x;
    TACC v = i < nblocks ? totals[i] : (TACC)0, x = v;
    for (int o = 1; o < 32; o <<= 1) { TACC y = dr_shfl_up(x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      TACC w = warp_tot[threadIdx.x], z = w;
      for (int o = 1; o < 32; o <<= 1) { TACC y = dr_shfl_up(z, o); if (threadIdx.x >= o) z += y; }
      warp_tot[threadIdx.x] = z - w;
    }
    __syncthreads();
    const TACC incl = x + warp_tot[threadIdx.x >> 5] + carry;
    if (i < nblocks) totals[i] = incl - v;               // exclusive prefix of block i
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
}
extern "C" __global__ void __launch_bounds__(256) NAME_final(const TIN* __restrict__ in,
    TACC* __restrict__ out, const TACC* __restrict__ offsets, i64 n) {
  __shared__ TACC warp_tot[8];
  const i64 base = (i64)blockIdx.x * 256 * ITEMS + (i64)threadIdx.x * ITEMS;
  TACC v[ITEMS], s = (TACC)0;
  for (int k = 0; k < ITEMS; ++k) { v[k] = base + k < n ? (TACC)in[base + k] : (TACC)0; s += v[k]; v[k] = s; }
  TACC x = s;
  for (int o = 1; o < 32; o <<= 1) { TACC y = dr_shfl_up(x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  TACC before = offsets[blockIdx.x] + (x - s);
  for (int w = 0; w < (threadIdx.x >> 5); ++w) before += warp_tot[w];
  for (int k = 0; k < ITEMS; ++k) if (base + k < n) out[base + k] = before + v[k];
}
'''
# Second-generation scans: 128-bit accesses laid out per warp and, for the 1-d case, ONE pass over
# the data.  A tile = THREADS * NV vectors of VE = 16 / sizeof(TIN) elements; warp w owns a
# contiguous chunk of it and lane l its vectors j * 32 + l (every load / store instruction of a
# warp covers 512 contiguous bytes).  Element order inside a chunk is (j, lane, e): per j a lane-
# local scan, a shuffle scan over the lanes and the running total of the earlier j.
#
# NAME_chain (1-d): G co-resident CTAs walk the tiles round by round (tile t = round * G + p).
# A CTA reads each of its tiles twice: one round AHEAD, from HBM, only to publish the tile's total
# as a 16-byte {value, flag} record (single 128-bit relaxed store, no fence), and in its own round,
# from L2, for the scan itself.  In between, the records of the whole round have long been
# published: every thread fetches one of them (the look-back is one round trip to L2), the tiles
# in front give the tile's prefix and all of them the next carry, which every CTA therefore
# computes for itself.  HBM traffic stays 8 B / element for float32 (three-phase scan: 12), and —
# unlike decoupled look-back — every prefix is the same fixed-order sum on every run: the result
# is reproducible bit for bit.  All G CTAs must be resident at once (the host sizes G from the
# occupancy query; a lost peer ends in a trap after ~10^8 polls, never in a hang).
_SCAN2_SRC = r'''
#define VE (16 / (int)sizeof(TIN))
#define NV NVVAL
#define SV ((int)sizeof(TACC) * VE / 16)
typedef dr_raw<16> V16;
union InVec { V16 q; TIN e[VE]; __device__ InVec() {} };
union OutVec { V16 q[SV]; TACC e[VE]; __device__ OutVec() {} };

template <typename T> __device__ __forceinline__ T dr_shfl_idx(T v, int src) {
  if (sizeof(T) == 8) {
    union { T t; struct { u32 lo, hi; } s; } u;
    u.t = v;
    u.s.lo = __shfl_sync(0xffffffffu, u.s.lo, src);
    u.s.hi = __shfl_sync(0xffffffffu, u.s.hi, src);
    return u.t;
  } else {
    union { T t; u32 w; } u;
    u.w = 0; u.t = v;
    u.w = __shfl_sync(0xffffffffu, u.w, src);
    return u.t;
  }
}
// A published total is ONE 16-byte record {value, flag}, written and polled with single 128-bit
// relaxed accesses (single-copy atomic): no fence on either side.  A fence, or the release /
// acquire pair of a separate flag word, waits for the thread's outstanding loads — here the
// prefetch of the next tile, i.e. a full DRAM latency on the critical path of every round
// (measured: 10.5 us per round with fences, see DESIGN.md).
struct alignas(16) ScRec { unsigned long long value, flag; };
__device__ __forceinline__ void sc_publish(ScRec* p, TACC v) {
  union { TACC t; unsigned long long w; } u; u.w = 0ull; u.t = v;
  asm volatile("{ .reg .b128 q; mov.b128 q, {%1, %2}; st.relaxed.gpu.global.b128 [%0], q; }"
               :: "l"(p), "l"(u.w), "l"(1ull) : "memory");
}
__device__ __forceinline__ bool sc_peek(const ScRec* p, TACC& v) {
  unsigned long long lo, hi;
  asm volatile("{ .reg .b128 q; ld.relaxed.gpu.global.b128 q, [%2]; mov.b128 {%0, %1}, q; }"
               : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
  union { TACC t; unsigned long long w; } u; u.w = lo;
  v = u.t;
  return hi != 0ull;
}
// vectors of one thread for the tile that starts at `src` and has `rem` elements left in its row
// (a full tile takes the path without a single bounds test)
template <int THREADS>
__device__ __forceinline__ void sc_load(const TIN* __restrict__ src, int rem, V16 (&raw)[NV]) {
  const int off = ((threadIdx.x >> 5) * (32 * NV) + (threadIdx.x & 31)) * VE;
  if (rem >= THREADS * NV * VE) {
    const V16* q = reinterpret_cast<const V16*>(src + off);
#pragma unroll
    for (int j = 0; j < NV; ++j) raw[j] = dr_ld_raw<true>(q + j * 32);
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int idx = off + j * 32 * VE;
      if (idx + VE <= rem) raw[j] = dr_ld_raw<true>(reinterpret_cast<const V16*>(src + idx));
      else {
        InVec u;
#pragma unroll
        for (int e = 0; e < VE; ++e) u.e[e] = idx + e < rem ? src[idx + e] : (TIN)0;
        raw[j] = u.q;
      }
    }
  }
}
// inclusive prefixes relative to the start of the warp's chunk; returns the chunk total
__device__ __forceinline__ TACC sc_warp_scan(const V16 (&raw)[NV], TACC (&v)[NV][VE]) {
  const int lane = threadIdx.x & 31;
  TACC run = (TACC)0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    InVec u; u.q = raw[j];
    TACC s = (TACC)0;
#pragma unroll
    for (int e = 0; e < VE; ++e) { s += (TACC)u.e[e]; v[j][e] = s; }
    TACC x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const TACC y = dr_shfl_up(x, o); if (lane >= o) x += y; }
    TACC ex = dr_shfl_up(x, 1);
    if (lane == 0) ex = (TACC)0;
    const TACC base = run + ex;
#pragma unroll
    for (int e = 0; e < VE; ++e) v[j][e] = base + v[j][e];
    run += dr_shfl_idx(x, 31);
  }
  return run;
}
// warp 0: exclusive scan of the warp totals (sW -> sEx), returns the tile total in every lane
template <int THREADS>
__device__ __forceinline__ TACC sc_block_offsets(const TACC* sW, TACC* sEx) {
  const int lane = threadIdx.x & 31;
  const TACC w = lane < THREADS / 32 ? sW[lane] : (TACC)0;
  TACC x = w;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const TACC y = dr_shfl_up(x, o); if (lane >= o) x += y; }
  TACC ex = dr_shfl_up(x, 1);
  if (lane == 0) ex = (TACC)0;
  sEx[lane] = ex;
  return dr_shfl_idx(x, 31);
}
template <int THREADS>
__device__ __forceinline__ void sc_store(TACC* __restrict__ dst, int rem, const TACC (&v)[NV][VE], TACC add) {
  const int off = ((threadIdx.x >> 5) * (32 * NV) + (threadIdx.x & 31)) * VE;
  if (rem >= THREADS * NV * VE) {
    V16* q = reinterpret_cast<V16*>(dst + off);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      OutVec u;
#pragma unroll
      for (int e = 0; e < VE; ++e) u.e[e] = add + v[j][e];
#pragma unroll
      for (int k = 0; k < SV; ++k) dr_st_raw<true>(q + j * 32 * SV + k, u.q[k]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int idx = off + j * 32 * VE;
      OutVec u;
#pragma unroll
      for (int e = 0; e < VE; ++e) u.e[e] = add + v[j][e];
      if (idx + VE <= rem) {
#pragma unroll
        for (int k = 0; k < SV; ++k) dr_st_raw<true>(reinterpret_cast<V16*>(dst + idx) + k, u.q[k]);
      } else {
#pragma unroll
        for (int e = 0; e < VE; ++e) if (idx + e < rem) dst[idx + e] = u.e[e];
      }
    }
  }
}

// total of the warp's chunk (local sums, one butterfly)
__device__ __forceinline__ TACC sc_warp_total(const V16 (&raw)[NV]) {
  const int lane = threadIdx.x & 31;
  TACC s = (TACC)0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    InVec u; u.q = raw[j];
#pragma unroll
    for (int e = 0; e < VE; ++e) s += (TACC)u.e[e];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += dr_shfl_idx(s, lane ^ o);
  return s;
}
__device__ __forceinline__ int sc_rem(i64 n, i64 base, i64 tile) {
  const i64 left = n - base;
  return left < tile ? (int)left : (int)tile;
}

extern "C" __global__ void __launch_bounds__(CTHREADS, CPS) NAME_chain(const TIN* __restrict__ in,
    TACC* __restrict__ out, i64 n, i64 ntiles, ScRec* agg) {
  __shared__ TACC sW[32], sWn[32], sEx[32], sExn[32], sLt[32], sAll[32];
  constexpr i64 TILE = (i64)CTHREADS * NV * VE;
  const int G = gridDim.x, p = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  V16 cur[NV], nxt[NV];
  // prologue: the totals of the CTA's first SLACK tiles; the tile after them starts to load
#pragma unroll 1
  for (int r = 0; r < SLACK; ++r) {
    const i64 t = (i64)p + (i64)r * G;
    if (t >= ntiles) break;
    sc_load<CTHREADS>(in + t * TILE, sc_rem(n, t * TILE, TILE), nxt);
    const TACC W0 = sc_warp_total(nxt);
    if (lane == 0) sWn[warp] = W0;
    __syncthreads();
    if (warp == 0) {
      const TACC total = sc_block_offsets<CTHREADS>(sWn, sExn);
      if (lane == 0) sc_publish(agg + t, total);
    }
    __syncthreads();
  }
  if ((i64)p + (i64)SLACK * G < ntiles)
    sc_load<CTHREADS>(in + ((i64)p + (i64)SLACK * G) * TILE, sc_rem(n, ((i64)p + (i64)SLACK * G) * TILE, TILE), nxt);
  TACC carry = (TACC)0;
  i64 rho = 0;
  for (i64 t = p; t < ntiles; t += G, ++rho) {
    const i64 ahead = t + (i64)SLACK * G;
    const bool has_next = ahead < ntiles;
    // a tile is read twice: SLACK rounds ahead for its total (from HBM, prefetched a round earlier
    // still), and now for the scan itself (from L2: it was read SLACK + 1 rounds ago)
    sc_load<CTHREADS>(in + t * TILE, sc_rem(n, t * TILE, TILE), cur);
    if (has_next) {
      const TACC Wn = sc_warp_total(nxt);
      if (lane == 0) sWn[warp] = Wn;
    }
    if (ahead + G < ntiles)
      sc_load<CTHREADS>(in + (ahead + G) * TILE, sc_rem(n, (ahead + G) * TILE, TILE), nxt);
    TACC v[NV][VE];
    const TACC W = sc_warp_scan(cur, v);
    if (lane == 0) sW[warp] = W;
    __syncthreads();
    if (warp == 0) {
      if (has_next) {                                   // published SLACK rounds before it is needed
        const TACC total = sc_block_offsets<CTHREADS>(sWn, sExn);
        if (lane == 0) sc_publish(agg + ahead, total);
      }
      sc_block_offsets<CTHREADS>(sW, sEx);
    }
    // all totals of this round, one record per thread (a single round trip to L2): the tiles in
    // front of this one give its prefix, all of them the next carry — the same fixed-order sums
    // in every CTA, so no carry has to be communicated and every run produces the same bits
    const ScRec* round = agg + rho * G;
    const i64 left = ntiles - rho * G;
    const int Gr = left < G ? (int)left : G;
    TACC lt, all;
    for (long long spins = 0;; ++spins) {
      bool ok = true;
      lt = (TACC)0; all = (TACC)0;
      for (int q = threadIdx.x; q < Gr; q += CTHREADS) {      // one pass: G <= CTHREADS
        TACC a;
        ok = ok & sc_peek(round + q, a);
        all += a;
        if (q < p) lt += a;
      }
      if (__syncthreads_and(ok)) break;
      if (spins > 100000000ll) __trap();                // tens of seconds: a lost peer, not a slow one
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lt += dr_shfl_idx(lt, lane ^ o); all += dr_shfl_idx(all, lane ^ o); }
    if (lane == 0) { sLt[warp] = lt; sAll[warp] = all; }
    __syncthreads();
    // the warps' partial sums, again in a fixed order (every warp repeats the same butterfly)
    TACC prefix = lane < CTHREADS / 32 ? sLt[lane] : (TACC)0;
    TACC round_total = lane < CTHREADS / 32 ? sAll[lane] : (TACC)0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      prefix += dr_shfl_idx(prefix, lane ^ o);
      round_total += dr_shfl_idx(round_total, lane ^ o);
    }
    prefix = carry + prefix;
    carry += round_total;
    sc_store<CTHREADS>(out + t * TILE, sc_rem(n, t * TILE, TILE), v, prefix + sEx[warp]);
  }
}

// rows of a matrix scanned along the contiguous axis: one CTA per row, tile by tile with a carry
extern "C" __global__ void __launch_bounds__(256, 4) NAME_rowscan2(const TIN* __restrict__ in,
    TACC* __restrict__ out, i64 rows, i64 n) {
  __shared__ TACC sW[32], sEx[32];
  __shared__ TACC sTotal;
  constexpr i64 TILE = (i64)256 * NV * VE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (i64 r = blockIdx.x; r < rows; r += gridDim.x) {
    const TIN* p = in + r * n;
    TACC* q = out + r * n;
    TACC carry = (TACC)0;
    V16 cur[NV], nxt[NV];
    sc_load<256>(p, sc_rem(n, 0, TILE), cur);
    for (i64 base = 0; base < n; base += TILE) {
      if (base + TILE < n) sc_load<256>(p + base + TILE, sc_rem(n, base + TILE, TILE), nxt);
      TACC v[NV][VE];
      const TACC W = sc_warp_scan(cur, v);
      if (lane == 0) sW[warp] = W;
      __syncthreads();
      if (warp == 0) {
        const TACC total = sc_block_offsets<256>(sW, sEx);
        if (lane == 0) sTotal = total;
      }
      __syncthreads();
      sc_store<256>(q + base, sc_rem(n, base, TILE), v, carry + sEx[warp]);
      carry += sTotal;
#pragma unroll
      for (int j = 0; j < NV; ++j) cur[j] = nxt[j];
    }
    __syncthreads();
  }
}
'''
_SHFL_UP = r'''
template <typename T> __device__ __forceinline__ T dr_shfl_up(T v, int delta) {
  if (sizeof(T) == 8) {
    union { T t; struct { u32 lo, hi; } s; } u;
    u.t = v;
    u.s.lo = __shfl_up_sync(0xffffffffu, u.s.lo, delta);
    u.s.hi = __shfl_up_sync(0xffffffffu, u.s.hi, delta);
    return u.t;
  } else {
    union { T t; u32 w; } u;
    u.w = 0; u.t = v;
    u.w = __shfl_up_sync(0xffffffffu, u.w, delta);
    return u.t;
  }
}
'''


_scan_sets = {}                 # key -> {suffix: Kernel}: built once (the source hash alone costs ~0.7 ms)


def _scan_kernels(in_dt, acc_dt):
    # one module holds all four entry points; load each by name
    key = ("scan", np.dtype(in_dt).str, np.dtype(acc_dt).str)
    hit = _scan_sets.get(key)
    if hit is not None:
        return hit
    name = engine.kernel_name(key)
    src = _SHFL_UP + _SCAN_SRC.replace("NAME", name).replace("TIN", ctype(in_dt)).replace("TACC", ctype(acc_dt))
    _, cubin = engine.compile_source(name, src)
    out = {}
    for suffix in ("serial", "partials", "offsets", "final", "rowscan", "chunk_totals", "chunk_offsets",
                   "chunk_final"):
        k = engine._kernels.get(key + (suffix,))
        if k is None:
            k = engine._kernels[key + (suffix,)] = engine.Kernel(f"{name}_{suffix}", src, cubin, {})
        out[suffix] = k
    _scan_sets[key] = out
    return out


# 2 CTAs of 512 threads per SM; totals are published one round ahead.  Measured alternatives
# (profiles/r2_scan_chain_experiments.txt): 1024 x 1 and 256 x 4 are within 2 %, 2 or 3 rounds
# ahead are 5 - 8 % slower (the second read starts to miss L2), smaller tiles (8 or 4 elements per
# thread with 3 or 4 CTAs per SM) are 30 - 100 % slower: the cost is per round
_CHAIN_THREADS, _CHAIN_CPS, _CHAIN_SLACK = 512, 2, 1


def _scan2_kernels(in_dt, acc_dt, sm_count):
    """(kernels, NV) of the second-generation scans for one (input, accumulator) type pair."""
    in_dt, acc_dt = np.dtype(in_dt), np.dtype(acc_dt)
    ve = 16 // in_dt.itemsize
    nv = max(1, min(4, 64 // (ve * acc_dt.itemsize)))
    key = ("scan2", in_dt.str, acc_dt.str, nv, _CHAIN_THREADS, _CHAIN_CPS, _CHAIN_SLACK)
    hit = _scan_sets.get(key)
    if hit is not None:
        return hit, nv
    name = engine.kernel_name(key)
    src = _SHFL_UP + (_SCAN2_SRC.replace("NAME", name).replace("NVVAL", str(nv)).replace("SLACK", str(_CHAIN_SLACK))
                      .replace("CTHREADS", str(_CHAIN_THREADS)).replace("CPS", str(_CHAIN_CPS))
                      .replace("TIN", ctype(in_dt)).replace("TACC", ctype(acc_dt)))
    _, cubin = engine.compile_source(name, src)
    out = {}
    for suffix in ("chain", "rowscan2"):
        k = engine._kernels.get(key + (suffix,))
        if k is None:
            k = engine._kernels[key + (suffix,)] = engine.Kernel(f"{name}_{suffix}", src, cubin, {})
        out[suffix] = k
    _scan_sets[key] = out
    return out, nv


def _chain_scan(src, out, n, in_dt, acc_dt):
    """1-d inclusive scan of `n` contiguous elements in one pass (NAME_chain above)."""
    dev = src.dev
    st = engine.dev_state(dev)
    ks, nv = _scan2_kernels(in_dt, acc_dt, st.sm_count)
    tile = _CHAIN_THREADS * nv * (16 // in_dt.itemsize)
    ntiles = -(-n // tile)
    kern = ks["chain"]
    G = st.sm_count * _CHAIN_CPS
    if dev >= 0 and kern.blocks_per_sm(dev, _CHAIN_THREADS) < _CHAIN_CPS:
        return False                      # the CTAs of a round must be co-resident
    G = min(G, ntiles)
    rounds = -(-ntiles // G)
    d = dev if dev >= 0 else None
    recs = DeviceArray.empty((ntiles, 2), np.uint64, d)                   # {total, flag} per tile
    if dev >= 0:
        engine.check(engine.lib.drc_memset_async(dev, 0, recs.ptr, 0, recs.nbytes))
    a = Args()
    a.ptr(src.ptr); a.ptr(out.ptr); a.i64(n); a.i64(ntiles); a.ptr(recs.ptr)
    launch(kern, dev, G, _CHAIN_THREADS, a)
    return True


def _grid_1d(n):
    return max(1, min(148 * 16, -(-n // 256)))


def cumsum(arr, axis=None):
    """np.cumsum on a DeviceArray (accumulator dtype = NumPy's add.reduce promotion)."""
    src = arr if arr.is_contiguous else arr.copy()
    in_dt = src.dtype
    acc_dt = np.cumsum(np.empty(0, in_dt)).dtype
    if axis is None:
        outer, n, inner, out_shape = 1, src.size, 1, (src.size,)
    else:
        axis %= src.ndim
        outer = int(np.prod(src.shape[:axis], dtype=np.int64))
        n = src.shape[axis]
        inner = int(np.prod(src.shape[axis + 1:], dtype=np.int64))
        out_shape = src.shape
    dev = src.dev
    out = DeviceArray.empty(out_shape, acc_dt, dev if dev >= 0 else None)
    if src.size == 0:
        return out
    ks = _scan_kernels(in_dt, acc_dt)
    d = dev if dev >= 0 else None
    gen2 = not os.environ.get("DR_SCAN_GEN1") and src.ptr % 16 == 0
    if gen2 and inner == 1 and outer >= 64 and n >= 1024 and (n * in_dt.itemsize) % 16 == 0:
        ks2, _ = _scan2_kernels(in_dt, acc_dt, engine.dev_state(dev).sm_count)
        a = Args()
        a.ptr(src.ptr); a.ptr(out.ptr); a.i64(outer); a.i64(n)
        launch(ks2["rowscan2"], dev, min(outer, 148 * 8), 256, a)
        return out
    if gen2 and outer == 1 and inner == 1 and n >= (1 << 20) and _chain_scan(src, out, n, in_dt, acc_dt):
        return out
    if inner == 1 and outer >= 64 and n >= 64:
        a = Args()
        a.ptr(src.ptr); a.ptr(out.ptr); a.i64(outer); a.i64(n)
        launch(ks["rowscan"], dev, min(outer, 148 * 8), 256, a)
        return out
    parts = min(-(-148 * 2048 // max(outer * inner, 1)), n // 16) if inner > 1 else 1
    if parts > 1:
        chunk = -(-n // parts)
        parts = -(-n // chunk)
        totals = DeviceArray.empty((outer, parts, inner), acc_dt, d)
        a = Args()
        a.ptr(src.ptr); a.ptr(totals.ptr); a.i64(outer); a.i64(n); a.i64(inner); a.i64(chunk); a.i64(parts)
        launch(ks["chunk_totals"], dev, _grid_1d(outer * parts * inner), 256, a)
        a = Args()
        a.ptr(totals.ptr); a.i64(outer); a.i64(inner); a.i64(parts)
        launch(ks["chunk_offsets"], dev, _grid_1d(outer * inner), 256, a)
        a = Args()
        a.ptr(src.ptr); a.ptr(out.ptr); a.ptr(totals.ptr); a.i64(outer); a.i64(n); a.i64(inner)
        a.i64(chunk); a.i64(parts)
        launch(ks["chunk_final"], dev, _grid_1d(outer * parts * inner), 256, a)
        return out
    if outer * inner >= 4096 or n < 65536:
        a = Args()
        a.ptr(src.ptr); a.ptr(out.ptr); a.i64(outer); a.i64(n); a.i64(inner)
        launch(ks["serial"], dev, max(1, min(148 * 8, -(-outer * inner // 256))), 256, a)
        return out
    per_block = 256 * 8
    rows_out = out.reshape(outer, n) if outer > 1 else None
    for r in range(outer):
        sptr = src.ptr + r * n * in_dt.itemsize
        optr = out.ptr + r * n * acc_dt.itemsize
        nblocks = -(-n // per_block)
        totals = DeviceArray.empty((nblocks,), acc_dt, dev if dev >= 0 else None)
        a = Args(); a.ptr(sptr); a.ptr(totals.ptr); a.i64(n)
        launch(ks["partials"], dev, nblocks, 256, a)
        a = Args(); a.ptr(totals.ptr); a.scalar(nblocks, np.int32)
        launch(ks["offsets"], dev, 1, 1024, a)
        a = Args(); a.ptr(sptr); a.ptr(optr); a.ptr(totals.ptr); a.i64(n)
        launch(ks["final"], dev, nblocks, 256, a)
    return out


# ------------------------------------------------------------------------------ Philox RNG
_PHILOX_SRC = r'''
// Philox4x32-10 (Salmon et al., SC'11): counter = (element index / 4, stream), key = seed.
__device__ __forceinline__ void philox_round(u32 (&c)[4], u32 (&k)[2]) {
  const u32 hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const u32 hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const u32 n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__device__ __forceinline__ void philox4x32(u64 counter, u64 stream, u64 seed, u32 (&out)[4]) {
  u32 c[4] = {(u32)counter, (u32)(counter >> 32), (u32)stream, (u32)(stream >> 32)};
  u32 k[2] = {(u32)seed, (u32)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__device__ __forceinline__ double u53(u32 a, u32 b) {     // [0, 1) with 53 random bits
  return (double)((((u64)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}
// mode 0: uniform [0,1)   1: standard normal (Box-Muller)   2: integers in [lo, lo+span)
extern "C" __global__ void __launch_bounds__(256) NAME(TOUT* __restrict__ out, i64 n, u64 seed,
    u64 stream, int mode, i64 lo, u64 span) {
  const i64 pairs = (n + 1) / 2;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (i64)gridDim.x * blockDim.x) {
    u32 r[4];
    philox4x32((u64)p, stream, seed, r);
    double a = u53(r[0], r[1]), b = u53(r[2], r[3]);
    double v0, v1;
    if (mode == 1) {
      const double rad = sqrt(-2.0 * log(1.0 - a)), ang = 6.283185307179586 * b;
      v0 = rad * cos(ang); v1 = rad * sin(ang);
    } else if (mode == 2) {
      v0 = (double)(lo + (i64)((((u64)r[0] << 32) | r[1]) % span));
      v1 = (double)(lo + (i64)((((u64)r[2] << 32) | r[3]) % span));
    } else { v0 = a; v1 = b; }
    out[2 * p] = (TOUT)v0;
    if (2 * p + 1 < n) out[2 * p + 1] = (TOUT)v1;
  }
}
'''
_rng = {"seed": 0x5DEECE66D, "stream": 0}


def seed(s=None):
    _rng["seed"] = 0x5DEECE66D if s is None else int(s) & 0xFFFFFFFFFFFFFFFF
    _rng["stream"] = 0


def philox(shape, dtype, mode, lo=0, span=1):
    """Fill a fresh DeviceArray from the Philox stream (each call consumes one stream id)."""
    dtype = np.dtype(dtype)
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    shape = tuple(int(s) for s in shape)
    dev = current_device() if not engine.is_dry() else -1
    out = DeviceArray.empty(shape, dtype, dev if dev >= 0 else None)
    n = out.size
    if n == 0:
        return out
    key = ("philox", dtype.str)
    kern = get_kernel(key, lambda name: _PHILOX_SRC.replace("NAME", name).replace("TOUT", ctype(dtype)))
    a = Args()
    a.ptr(out.ptr); a.i64(n)
    a.raw(int(_rng["seed"]).to_bytes(8, "little"), 8)
    a.raw(int(_rng["stream"]).to_bytes(8, "little"), 8)
    a.scalar(mode, np.int32); a.i64(lo)
    a.raw(int(max(span, 1)).to_bytes(8, "little"), 8)
    _rng["stream"] += 1
    launch(kern, out.dev, max(1, min(148 * 8, -(-((n + 1) // 2) // 256))), 256, a)
    return out


# ------------------------------------------------------------------------------ fft helpers
_R2C_SRC = r'''
// real (or complex) rows of length n_in -> complex rows of length n_out (zero padded / truncated)
extern "C" __global__ void __launch_bounds__(256) NAME(const TIN* __restrict__ in, TC* __restrict__ out,
    i64 rows, i64 n_in, i64 n_out, int in_is_complex) {
  const i64 total = rows * n_out;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 r = idx / n_out, c = idx - r * n_out;
    TC re = (TC)0, im = (TC)0;
    if (c < n_in) {
      if (in_is_complex) { re = (TC)in[2 * (r * n_in + c)]; im = (TC)in[2 * (r * n_in + c) + 1]; }
      else re = (TC)in[r * n_in + c];
    }
    out[2 * idx] = re; out[2 * idx + 1] = im;
  }
}
'''


def pack_complex(src, n_out, complex_dtype):
    """(rows, n_in) real/complex DeviceArray -> (rows, n_out) complex DeviceArray."""
    cdt = np.dtype(complex_dtype)
    part = np.dtype(np.float32 if cdt == np.complex64 else np.float64)
    rows, n_in = src.shape
    is_c = src.dtype.kind == "c"
    in_part = np.dtype(np.float32 if src.dtype == np.complex64 else np.float64) if is_c else src.dtype
    out = DeviceArray.empty((rows, n_out), cdt, src.dev if src.dev >= 0 else None)
    key = ("r2c", in_part.str, part.str)
    kern = get_kernel(key, lambda name: _R2C_SRC.replace("NAME", name).replace("TIN", ctype(in_part))
                      .replace("TC", ctype(part)))
    a = Args()
    a.ptr(src.ptr); a.ptr(out.ptr); a.i64(rows); a.i64(n_in); a.i64(n_out)
    a.scalar(1 if is_c else 0, np.int32)
    launch(kern, src.dev, max(1, min(148 * 8, -(-rows * n_out // 256))), 256, a)
    return out


# ------------------------------------------------------------------------------ gather / compact
# Integer-array and boolean-mask indexing (reference delayarray.py:123-128 hands the key to CuPy).
# Elements move as opaque 1/2/4/8-byte words, so one specialisation per item size and index type.
_INDEX_SRC = r'''
// out[i, :] = src[idx[i], :]   (rows of `inner` words; negative indices wrap like NumPy's)
extern "C" __global__ void __launch_bounds__(256) NAME_take(const W* __restrict__ src,
    const IDX* __restrict__ idx, W* __restrict__ out, i64 n_idx, i64 inner, i64 n_src) {
  const i64 total = n_idx * inner;
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (i64)gridDim.x * blockDim.x) {
    const i64 i = k / inner, c = k - i * inner;
    i64 j = (i64)idx[i];
    if (j < 0) j += n_src;
    out[k] = src[j * inner + c];
  }
}
// dst[idx[i], :] = vals[i, :]   (val_rows == 1: one row broadcast to every index)
extern "C" __global__ void __launch_bounds__(256) NAME_put(W* __restrict__ dst,
    const IDX* __restrict__ idx, const W* __restrict__ vals, i64 n_idx, i64 inner, i64 n_dst, i64 val_rows) {
  const i64 total = n_idx * inner;
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (i64)gridDim.x * blockDim.x) {
    const i64 i = k / inner, c = k - i * inner;
    i64 j = (i64)idx[i];
    if (j < 0) j += n_dst;
    dst[j * inner + c] = vals[(val_rows == 1 ? 0 : i) * inner + c];
  }
}
// rows whose mask is set, packed in order: pos = inclusive prefix sum of the mask
// mode 0: out[pos-1, :] = src[i, :]     1: src[i, :] = out[pos-1, :] (masked assignment)
// mode 2: out[pos-1] = i (flatnonzero; W = long long, inner = 1)
extern "C" __global__ void __launch_bounds__(256) NAME_compact(W* __restrict__ src,
    const unsigned char* __restrict__ mask, const i64* __restrict__ pos, W* __restrict__ out,
    i64 n, i64 inner, int mode) {
  const i64 total = n * inner;
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (i64)gridDim.x * blockDim.x) {
    const i64 i = k / inner, c = k - i * inner;
    if (!mask[i]) continue;
    const i64 o = (pos[i] - 1) * inner + c;
    if (mode == 0) out[o] = src[k];
    else if (mode == 1) src[k] = out[o];
    else out[o] = (W)i;
  }
}
// ---- element masks (inner == 1) without a position array: tiles of 4096 mask bytes
// (1) per-tile counts, (2) exclusive scan of the counts (scan kernels), (3) per-tile scatter:
// every thread owns 16 consecutive mask bytes (one 128-bit load when aligned), its rank inside
// the tile comes from warp shuffles + 8 warp totals, the tile's base from the scanned counts.
__device__ __forceinline__ unsigned NAME_flags16(const unsigned char* __restrict__ mask, i64 first, i64 n) {
  unsigned bits = 0;
  if (first + 16 <= n && ((unsigned long long)(mask + first) & 15ull) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(mask + first);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      unsigned x = w[q];
      x |= x >> 4; x |= x >> 2; x |= x >> 1; x &= 0x01010101u;          // byte != 0 -> bit 0 of the byte
      bits |= ((x & 1u) | ((x >> 7) & 2u) | ((x >> 14) & 4u) | ((x >> 21) & 8u)) << (4 * q);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (first + j < n && mask[first + j]) bits |= 1u << j;
  }
  return bits;
}
extern "C" __global__ void __launch_bounds__(256) NAME_mcount(const unsigned char* __restrict__ mask,
    i64* __restrict__ counts, i64 n, i64 ntiles) {
  __shared__ int warp_tot[8];
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[ntiles] = 0;    // slot of the grand total
  for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    int c = __popc(NAME_flags16(mask, t * 4096 + (i64)threadIdx.x * 16, n));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int w = 0; w < 8; ++w) s += warp_tot[w];
      counts[t] = s;
    }
    __syncthreads();
  }
}
// mode 0: out[rank] = data[i]   1: data[i] = out[rank]   2: out[rank] = i
extern "C" __global__ void __launch_bounds__(256) NAME_mscatter(W* __restrict__ data,
    const unsigned char* __restrict__ mask, const i64* __restrict__ offsets, W* __restrict__ out,
    i64 n, i64 ntiles, int mode) {
  __shared__ int warp_tot[8];
  for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const i64 first = t * 4096 + (i64)threadIdx.x * 16;
    const unsigned bits = NAME_flags16(mask, first, n);
    const int c = __popc(bits);
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    i64 rank = offsets[t] + (x - c);
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) rank += warp_tot[w];
    unsigned rest = bits;
    while (rest) {
      const int j = __ffs(rest) - 1;
      rest &= rest - 1;
      if (mode == 0) out[rank] = data[first + j];
      else if (mode == 1) data[first + j] = out[rank];
      else out[rank] = (W)(first + j);
      ++rank;
    }
    __syncthreads();
  }
}
'''
_WORD = {1: "unsigned char", 2: "unsigned short", 4: "unsigned int", 8: "unsigned long long"}


def _index_kernels(itemsize, idx_dt):
    key = ("index", itemsize, np.dtype(idx_dt).str)
    if key + ("take",) in engine._kernels:
        return {suffix: engine._kernels[key + (suffix,)]
                for suffix in ("take", "put", "compact", "mcount", "mscatter")}
    name = engine.kernel_name(key)
    src = _INDEX_SRC.replace("NAME", name).replace("IDX", ctype(idx_dt))
    src = src.replace("W*", _WORD[itemsize] + "*").replace("(W)", f"({_WORD[itemsize]})")
    _, cubin = engine.compile_source(name, src)
    out = {}
    for suffix in ("take", "put", "compact", "mcount", "mscatter"):
        k = engine._kernels.get(key + (suffix,))
        if k is None:
            k = engine._kernels[key + (suffix,)] = engine.Kernel(f"{name}_{suffix}", src, cubin, {})
        out[suffix] = k
    return out


def _grid(n):
    return max(1, min(148 * 8, -(-n // 256)))


def _word_size(dt):
    if dt.itemsize not in _WORD:
        raise NotImplementedError(f"indexing arrays of {dt}")
    return dt.itemsize


def _check_bounds(idx, n):
    """IndexError for an out-of-range index, like NumPy (two fused reductions, one sync)."""
    from .delayarray import NPArray
    if idx.size == 0 or idx.dev < 0:
        return
    lo, hi = int(np.min(NPArray(idx)).get()), int(np.max(NPArray(idx)).get())
    if lo < -n or hi >= n:
        bad = lo if lo < -n else hi
        raise IndexError(f"index {bad} is out of bounds for axis 0 with size {n}")


# rows whose byte length is a multiple of 16: the same gather in 128-bit words (one index
# decomposition per 16 bytes; a 64 KiB row moves as 4096 coalesced vector copies)
_TAKE16_SRC = r'''
extern "C" __global__ void __launch_bounds__(256) NAME(const dr_raw<16>* __restrict__ src,
    const IDX* __restrict__ idx, dr_raw<16>* __restrict__ out, i64 n_idx, i64 inner, i64 n_src) {
  const i64 total = n_idx * inner;
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (i64)gridDim.x * blockDim.x) {
    const i64 i = k / inner, c = k - i * inner;
    i64 j = (i64)idx[i];
    if (j < 0) j += n_src;
    dr_st_raw<true>(out + k, dr_ld_raw<true>(src + j * inner + c));
  }
}
'''


def take(src, idx):
    """src[idx] for an integer DeviceArray idx of any shape (gathers along axis 0)."""
    if idx.dtype.kind not in "iu":
        raise IndexError("arrays used as indices must be of integer (or boolean) type")
    src = src if src.is_contiguous else src.copy()
    idx = idx if idx.is_contiguous else idx.copy()
    if src.ndim == 0:
        raise IndexError("too many indices for array")
    n_src = src.shape[0]
    inner = int(np.prod(src.shape[1:], dtype=np.int64))
    _check_bounds(idx, n_src)
    out = DeviceArray.empty(tuple(idx.shape) + tuple(src.shape[1:]), src.dtype, src.dev if src.dev >= 0 else None)
    if out.size == 0:
        return out
    row_bytes = inner * src.dtype.itemsize
    if row_bytes >= 64 and row_bytes % 16 == 0 and src.ptr % 16 == 0 and out.ptr % 16 == 0:
        kern = get_kernel(("take16", idx.dtype.str), lambda name: _TAKE16_SRC.replace("NAME", name)
                          .replace("IDX", ctype(idx.dtype)))
        a = Args()
        a.ptr(src.ptr); a.ptr(idx.ptr); a.ptr(out.ptr); a.i64(idx.size); a.i64(row_bytes // 16); a.i64(n_src)
        launch(kern, src.dev, max(1, min(148 * 16, -(-(idx.size * (row_bytes // 16)) // 256))), 256, a)
        return out
    ks = _index_kernels(_word_size(src.dtype), idx.dtype)
    a = Args()
    a.ptr(src.ptr); a.ptr(idx.ptr); a.ptr(out.ptr); a.i64(idx.size); a.i64(inner); a.i64(n_src)
    launch(ks["take"], src.dev, _grid(out.size), 256, a)
    return out


def put(dst, idx, vals):
    """dst[idx] = vals (integer index array along axis 0; duplicate indices: one of the writes
    wins, unspecified which -- NumPy keeps the last).  dst must be contiguous."""
    if idx.dtype.kind not in "iu":
        raise IndexError("arrays used as indices must be of integer (or boolean) type")
    if not dst.is_contiguous:
        raise NotImplementedError("integer-array assignment into a non-contiguous view")
    idx = idx if idx.is_contiguous else idx.copy()
    n_dst = dst.shape[0]
    inner = int(np.prod(dst.shape[1:], dtype=np.int64))
    _check_bounds(idx, n_dst)
    row_shape = tuple(dst.shape[1:])
    full = tuple(idx.shape) + row_shape
    vals = vals.astype(dst.dtype) if vals.dtype != dst.dtype else vals
    if tuple(vals.shape) == full:
        rows = idx.size
    else:
        np.broadcast_shapes(tuple(vals.shape), full)                  # raises like NumPy
        if vals.ndim <= len(row_shape) and vals.size in (1, max(inner, 1)):
            vals, rows = vals.broadcast_to(row_shape).copy() if tuple(vals.shape) != row_shape else vals, 1
        else:
            vals, rows = vals.broadcast_to(full).copy(), idx.size
    vals = vals if vals.is_contiguous else vals.copy()
    if idx.size == 0 or dst.size == 0:
        return
    ks = _index_kernels(_word_size(dst.dtype), idx.dtype)
    a = Args()
    a.ptr(dst.ptr); a.ptr(idx.ptr); a.ptr(vals.ptr); a.i64(idx.size); a.i64(inner); a.i64(n_dst); a.i64(rows)
    launch(ks["put"], dst.dev, _grid(idx.size * max(inner, 1)), 256, a)
    dst.buf.version += 1


def _mask_layout(src, mask):
    """(rows, inner) such that mask selects rows of `inner` words of the contiguous src."""
    if mask.dtype != np.dtype(bool):
        raise IndexError("boolean index expected")
    if tuple(mask.shape) != tuple(src.shape[:mask.ndim]):
        raise IndexError(f"boolean index did not match indexed array: mask shape {tuple(mask.shape)}, "
                         f"array shape {tuple(src.shape)}")
    rows = int(np.prod(mask.shape, dtype=np.int64))
    inner = int(np.prod(src.shape[mask.ndim:], dtype=np.int64))
    return rows, inner


def _mask_positions(mask):
    mask = mask if mask.is_contiguous else mask.copy()
    pos = cumsum(mask.reshape(-1), None)                              # intp inclusive prefix
    if pos.dtype != np.dtype(np.int64):
        pos = pos.astype(np.int64)
    count = int(pos[-1:].get()[0]) if (mask.size and mask.dev >= 0) else 0
    return mask, pos, count


def _mask_tiles(mask, itemsize):
    """Element masks: per-tile counts scanned in place -> (mask, offsets, count, ntiles, kernels).
    offsets[t] = number of set elements before tile t (tiles of 4096), offsets[ntiles] = total."""
    mask = mask if mask.is_contiguous else mask.copy()
    n = mask.size
    ntiles = -(-n // 4096)
    dev = mask.dev
    ks = _index_kernels(itemsize, np.int64)
    counts = DeviceArray.empty((ntiles + 1,), np.int64, dev if dev >= 0 else None)
    a = Args(); a.ptr(mask.ptr); a.ptr(counts.ptr); a.i64(n); a.i64(ntiles)
    launch(ks["mcount"], dev, max(1, min(ntiles, 148 * 8)), 256, a)
    a = Args(); a.ptr(counts.ptr); a.scalar(ntiles + 1, np.int32)
    launch(_scan_kernels(np.int64, np.int64)["offsets"], dev, 1, 1024, a)
    count = int(counts[ntiles:].get()[0]) if dev >= 0 else 0
    return mask, counts, count, ntiles, ks


def _mask_scatter(ks, data, mask, offsets, out, n, ntiles, mode):
    a = Args()
    a.ptr(data.ptr); a.ptr(mask.ptr); a.ptr(offsets.ptr); a.ptr(out.ptr); a.i64(n); a.i64(ntiles)
    a.scalar(mode, np.int32)
    launch(ks["mscatter"], mask.dev, max(1, min(ntiles, 148 * 8)), 256, a)


def compress(src, mask):
    """src[mask] for a boolean DeviceArray covering the leading dimensions of src."""
    src = src if src.is_contiguous else src.copy()
    rows, inner = _mask_layout(src, mask)
    if inner == 1 and rows > 0:
        mask, offsets, count, ntiles, ks = _mask_tiles(mask, _word_size(src.dtype))
        # (trailing dimensions of size 1 stay: a row mask on an (n, 1) array gives (count, 1))
        out = DeviceArray.empty((count,) + tuple(src.shape[mask.ndim:]), src.dtype,
                                src.dev if src.dev >= 0 else None)
        if count or src.dev < 0:
            _mask_scatter(ks, src, mask, offsets, out, rows, ntiles, 0)
        return out
    mask, pos, count = _mask_positions(mask)
    out = DeviceArray.empty((count,) + tuple(src.shape[mask.ndim:]), src.dtype, src.dev if src.dev >= 0 else None)
    if out.size == 0 and src.dev >= 0:
        return out
    ks = _index_kernels(_word_size(src.dtype), np.int64)
    a = Args()
    a.ptr(src.ptr); a.ptr(mask.ptr); a.ptr(pos.ptr); a.ptr(out.ptr); a.i64(rows); a.i64(inner)
    a.scalar(0, np.int32)
    launch(ks["compact"], src.dev, _grid(rows * max(inner, 1)), 256, a)
    return out


def put_mask(dst, mask, vals):
    """dst[mask] = vals where vals holds one row per selected position (NumPy's compacted form)."""
    if not dst.is_contiguous:
        raise NotImplementedError("masked assignment of an array into a non-contiguous view")
    rows, inner = _mask_layout(dst, mask)
    fused = inner == 1 and rows > 0
    if fused:
        mask, pos, count, ntiles, fks = _mask_tiles(mask, _word_size(dst.dtype))
    else:
        mask, pos, count = _mask_positions(mask)
    want = (count,) + tuple(dst.shape[mask.ndim:])
    vals = vals.astype(dst.dtype) if vals.dtype != dst.dtype else vals
    if tuple(vals.shape) != want:
        if dst.dev >= 0:
            np.broadcast_shapes(tuple(vals.shape), want)
            if vals.ndim > len(want) or (vals.ndim == len(want) and vals.shape[0] != count):
                raise ValueError(f"NumPy boolean array indexing assignment cannot assign {vals.shape[0]} "
                                 f"input values to the {count} output values where the mask is true")
            vals = vals.broadcast_to(want).copy()
    vals = vals if vals.is_contiguous else vals.copy()
    if count == 0 and dst.dev >= 0:
        return
    if fused:
        _mask_scatter(fks, dst, mask, pos, vals, rows, ntiles, 1)
        dst.buf.version += 1
        return
    ks = _index_kernels(_word_size(dst.dtype), np.int64)
    a = Args()
    a.ptr(dst.ptr); a.ptr(mask.ptr); a.ptr(pos.ptr); a.ptr(vals.ptr); a.i64(rows); a.i64(inner)
    a.scalar(1, np.int32)
    launch(ks["compact"], dst.dev, _grid(rows * max(inner, 1)), 256, a)
    dst.buf.version += 1


def flatnonzero(mask):
    """Indices (int64) of the set elements of a boolean DeviceArray, flattened, ascending."""
    if mask.size:
        mask, offsets, count, ntiles, ks = _mask_tiles(mask, 8)
        out = DeviceArray.empty((count,), np.int64, mask.dev if mask.dev >= 0 else None)
        if count or mask.dev < 0:
            _mask_scatter(ks, out, mask, offsets, out, mask.size, ntiles, 2)
        return out
    mask, pos, count = _mask_positions(mask)
    out = DeviceArray.empty((count,), np.int64, mask.dev if mask.dev >= 0 else None)
    if count == 0 and mask.dev >= 0:
        return out
    ks = _index_kernels(8, np.int64)
    a = Args()
    a.ptr(out.ptr); a.ptr(mask.ptr); a.ptr(pos.ptr); a.ptr(out.ptr); a.i64(mask.size); a.i64(1)
    a.scalar(2, np.int32)
    launch(ks["compact"], mask.dev, _grid(mask.size), 256, a)
    return out


# ------------------------------------------------------------------------------ argmax / argmin
# (value, first index) pairs reduced in two phases over an (outer, n, inner) contiguous array.
# NumPy's rules: the FIRST extreme element wins; a nan counts as the extreme.
_ARG_SRC = r'''
struct NAME_pair { T v; i64 i; };
__device__ __forceinline__ bool NAME_isnan(T x) { return x != x; }
// does a beat b?  (ISMAX is 1 for argmax, 0 for argmin)
__device__ __forceinline__ bool NAME_beats(const NAME_pair& a, const NAME_pair& b) {
  if (b.i < 0) return a.i >= 0;
  if (a.i < 0) return false;
  const bool an = NAME_isnan(a.v), bn = NAME_isnan(b.v);
  if (an || bn) return an && (!bn || a.i < b.i);
  if (a.v == b.v) return a.i < b.i;
  return ISMAX ? a.v > b.v : a.v < b.v;
}
__device__ __forceinline__ NAME_pair NAME_shfl(NAME_pair p, int o) {
  NAME_pair q;
  q.v = dr_shfl_xor(p.v, o);
  q.i = dr_shfl_xor(p.i, o);
  return q;
}
// inner == 1: block (blockIdx.y = row, blockIdx.x = chunk of the row) -> one pair per (row, chunk).
// 16-byte aligned chunks are read as 128-bit vectors with a 32-bit running index and a two-
// predicate test per element (strict, so the first of equal values wins; nan beats everything
// once and is never beaten); the ragged rest goes through the general pair comparison.
extern "C" __global__ void __launch_bounds__(256) NAME_rows(const T* __restrict__ in,
    T* __restrict__ pv, i64* __restrict__ pi, i64 n, i64 chunk) {
  __shared__ T sv[8];
  __shared__ i64 si[8];
  const i64 row = blockIdx.y, lo = (i64)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
  const T* p = in + row * n;
  NAME_pair best; best.v = (T)0; best.i = -1;
  constexpr int V = 16 / (int)sizeof(T);
  i64 k0 = lo;
  if ((reinterpret_cast<unsigned long long>(p + lo) & 15ull) == 0ull) {
    const int nvec = (int)((hi - lo) / V);
    T bv = (T)0;
    int bi = -1;
#pragma unroll 4
    for (int q = threadIdx.x; q < nvec; q += 256) {
      const Vec<T, V> x = dr_ld<true, T, V>(p + lo + (i64)q * V);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const T c = x.v[e];
        const bool wins = !(ISMAX ? c <= bv : c >= bv) && !NAME_isnan(bv);
        if (bi < 0 || wins) { bv = c; bi = q * V + e; }
      }
    }
    if (bi >= 0) { best.v = bv; best.i = lo + bi; }
    k0 = lo + (i64)nvec * V;
  }
  for (i64 k = k0 + threadIdx.x; k < hi; k += 256) {
    NAME_pair c; c.v = p[k]; c.i = k;
    if (NAME_beats(c, best)) best = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const NAME_pair q = NAME_shfl(best, o); if (NAME_beats(q, best)) best = q; }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best.v; si[threadIdx.x >> 5] = best.i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { NAME_pair q; q.v = sv[w]; q.i = si[w]; if (NAME_beats(q, best)) best = q; }
    pv[row * gridDim.x + blockIdx.x] = best.v;
    pi[row * gridDim.x + blockIdx.x] = best.i;
  }
}
// fold many partial pairs of ONE output element per block (inner == 1, parts large)
extern "C" __global__ void __launch_bounds__(256) NAME_finalb(const T* __restrict__ pv,
    const i64* __restrict__ pi, i64* __restrict__ out, i64 parts) {
  __shared__ T sv[8];
  __shared__ i64 si[8];
  const i64 o = blockIdx.x;
  NAME_pair best; best.v = (T)0; best.i = -1;
  for (i64 s = threadIdx.x; s < parts; s += 256) {
    NAME_pair q; q.v = pv[o * parts + s]; q.i = pi[o * parts + s];
    if (NAME_beats(q, best)) best = q;
  }
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) { const NAME_pair q = NAME_shfl(best, w); if (NAME_beats(q, best)) best = q; }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best.v; si[threadIdx.x >> 5] = best.i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { NAME_pair q; q.v = sv[w]; q.i = si[w]; if (NAME_beats(q, best)) best = q; }
    out[o] = best.i;
  }
}
// inner > 1: one thread per (outer, inner) element and chunk of n (blockIdx.y), coalesced along inner
extern "C" __global__ void __launch_bounds__(256) NAME_cols(const T* __restrict__ in,
    T* __restrict__ pv, i64* __restrict__ pi, i64 outer, i64 n, i64 inner, i64 chunk) {
  const i64 total = outer * inner;
  const i64 lo = (i64)blockIdx.y * chunk, hi = lo + chunk < n ? lo + chunk : n;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 o = idx / inner, c = idx - o * inner;
    const T* p = in + o * n * inner + c;
    NAME_pair best; best.v = (T)0; best.i = -1;
#pragma unroll 4
    for (i64 k = lo; k < hi; ++k) {
      NAME_pair q; q.v = p[k * inner]; q.i = k;
      if (NAME_beats(q, best)) best = q;
    }
    pv[(o * gridDim.y + blockIdx.y) * inner + c] = best.v;
    pi[(o * gridDim.y + blockIdx.y) * inner + c] = best.i;
  }
}
// fold the `parts` partial pairs of every output element, in order
extern "C" __global__ void __launch_bounds__(256) NAME_final(const T* __restrict__ pv,
    const i64* __restrict__ pi, i64* __restrict__ out, i64 outer, i64 parts, i64 inner) {
  const i64 total = outer * inner;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 o = idx / inner, c = idx - o * inner;
    NAME_pair best; best.v = (T)0; best.i = -1;
    for (i64 s = 0; s < parts; ++s) {
      NAME_pair q; q.v = pv[(o * parts + s) * inner + c]; q.i = pi[(o * parts + s) * inner + c];
      if (NAME_beats(q, best)) best = q;
    }
    out[idx] = best.i;
  }
}
'''


def _arg_kernels(dt, is_max):
    key = ("arg", np.dtype(dt).str, bool(is_max))
    if key + ("rows",) in engine._kernels:
        return {s: engine._kernels[key + (s,)] for s in ("rows", "cols", "final", "finalb")}
    name = engine.kernel_name(key)
    T = ctype(dt) if np.dtype(dt) != np.dtype(bool) else "unsigned char"
    # the element type is a macro: this text follows the prelude, whose templates are already parsed
    src = f"#define T {T}\n" + _ARG_SRC.replace("NAME", name).replace("ISMAX", "1" if is_max else "0")
    _, cubin = engine.compile_source(name, src)
    out = {}
    for suffix in ("rows", "cols", "final", "finalb"):
        out[suffix] = engine._kernels[key + (suffix,)] = engine.Kernel(f"{name}_{suffix}", src, cubin, {})
    return out


def argreduce(src, axis, is_max):
    """np.argmax / np.argmin of a DeviceArray along one axis (None: flattened); int64 result."""
    src = src if src.is_contiguous else src.copy()
    if axis is None:
        outer, n, inner, out_shape = 1, src.size, 1, ()
    else:
        axis %= src.ndim
        outer = int(np.prod(src.shape[:axis], dtype=np.int64))
        n = src.shape[axis]
        inner = int(np.prod(src.shape[axis + 1:], dtype=np.int64))
        out_shape = tuple(src.shape[:axis]) + tuple(src.shape[axis + 1:])
    if n == 0:
        raise ValueError("attempt to get argmax of an empty sequence")
    dev = src.dev
    out = DeviceArray.empty(out_shape, np.int64, dev if dev >= 0 else None)
    if out.size == 0:
        return out
    ks = _arg_kernels(src.dtype, is_max)
    target = 148 * 8
    if inner == 1:
        parts = max(1, min(-(-target // outer), -(-n // 1024), 65535 if outer > 1 else 1 << 20))
        chunk = -(-(-(-n // parts)) // 1024) * 1024          # whole vectors per chunk; the index inside fits 32 bits
        parts = -(-n // chunk)
        rows_y = outer
        if rows_y > 65535:                      # grid.y limit: fall back to the column kernel
            inner, outer_c = 1, outer
            parts, chunk = 1, n
            pv = DeviceArray.empty((outer, 1), src.dtype, dev if dev >= 0 else None)
            pi = DeviceArray.empty((outer, 1), np.int64, dev if dev >= 0 else None)
            a = Args(); a.ptr(src.ptr); a.ptr(pv.ptr); a.ptr(pi.ptr); a.i64(outer_c); a.i64(n); a.i64(1); a.i64(chunk)
            launch(ks["cols"], dev, (_grid(outer), 1, 1), 256, a)
        else:
            pv = DeviceArray.empty((outer, parts), src.dtype, dev if dev >= 0 else None)
            pi = DeviceArray.empty((outer, parts), np.int64, dev if dev >= 0 else None)
            a = Args(); a.ptr(src.ptr); a.ptr(pv.ptr); a.ptr(pi.ptr); a.i64(n); a.i64(chunk)
            launch(ks["rows"], dev, (parts, rows_y, 1), 256, a)
    else:
        blocks_x = -(-(outer * inner) // 256)
        parts = max(1, min(-(-target // blocks_x), n // 16 if n >= 32 else 1, 65535))
        chunk = -(-n // parts)
        parts = -(-n // chunk)
        pv = DeviceArray.empty((outer, parts, inner), src.dtype, dev if dev >= 0 else None)
        pi = DeviceArray.empty((outer, parts, inner), np.int64, dev if dev >= 0 else None)
        a = Args(); a.ptr(src.ptr); a.ptr(pv.ptr); a.ptr(pi.ptr); a.i64(outer); a.i64(n); a.i64(inner); a.i64(chunk)
        launch(ks["cols"], dev, (min(blocks_x, 148 * 16), parts, 1), 256, a)
    if inner == 1 and parts >= 64 and outer <= 65535:
        a = Args(); a.ptr(pv.ptr); a.ptr(pi.ptr); a.ptr(out.ptr); a.i64(parts)
        launch(ks["finalb"], dev, outer, 256, a)
        return out
    a = Args(); a.ptr(pv.ptr); a.ptr(pi.ptr); a.ptr(out.ptr); a.i64(outer); a.i64(parts); a.i64(inner)
    launch(ks["final"], dev, _grid(outer * inner), 256, a)
    return out


# ------------------------------------------------------------------------------ transpose
# out[b, c, r] = in[b, r, c] through 32 x 33 shared-memory tiles: both the loads and the stores
# are coalesced (the index-decomposing `nd` kernel reads a transposed view a row pitch apart).
_TRANSPOSE_SRC = r'''
// tile = 32 source rows x TC source columns (TC = 64 for 4-byte words: 8 independent loads per
// thread, the same bytes in flight as 8-byte words with TC = 32)
#define TC TCVAL
extern "C" __global__ void __launch_bounds__(256) NAME(const W* __restrict__ in, W* __restrict__ out,
    i64 rows, i64 cols, i64 in_pitch, i64 out_pitch, i64 in_batch, i64 out_batch, i64 tiles_x, i64 tiles_y,
    i64 ntiles) {
  __shared__ W tile[32][TC + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
  for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const i64 b = t / (tiles_x * tiles_y), rem = t - b * tiles_x * tiles_y;
    const i64 by = rem / tiles_x, bx = rem - by * tiles_x;
    const W* src = in + b * in_batch;
    W* dst = out + b * out_batch;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const i64 r = by * 32 + ty + k * 8;
#pragma unroll
      for (int j = 0; j < TC / 32; ++j) {
        const i64 c = bx * TC + tx + j * 32;
        if (r < rows && c < cols) tile[ty + k * 8][tx + j * 32] = src[r * in_pitch + c];
      }
    }
    __syncthreads();
    const i64 oc = by * 32 + tx;                                   // output column = input row
#pragma unroll
    for (int k = 0; k < TC / 8; ++k) {
      const i64 orow = bx * TC + ty + k * 8;                       // output row = input column
      if (orow < cols && oc < rows) dst[orow * out_pitch + oc] = tile[tx][ty + k * 8];
    }
    __syncthreads();
  }
}
'''


def transposed_source(arr):
    """If `arr` (ndim >= 2) is exactly the last-two-axes transpose of a C-contiguous array, return
    (batch, rows, cols) of that source, else None."""
    if arr.ndim < 2 or arr.dtype.itemsize not in (4, 8) or arr.size == 0:
        return None
    item = arr.dtype.itemsize
    n_r, n_c = arr.shape[-1], arr.shape[-2]                        # source is (.., n_r, n_c)
    if arr.strides[-2] != item or n_r < 32 or n_c < 32:
        return None
    if arr.ndim == 2 and arr.strides[-1] % item == 0 and arr.strides[-1] >= n_c * item:
        return 1, n_r, n_c                                         # a column block of a matrix: pitch > n_c
    if arr.strides[-1] != n_c * item:
        return None
    batch, pitch = 1, n_r * n_c * item
    for n, st in zip(reversed(arr.shape[:-2]), reversed(arr.strides[:-2])):
        if n != 1 and st != pitch:
            return None
        pitch *= n
        batch *= n
    return batch, n_r, n_c


def transpose_copy(arr, out=None):
    """C-contiguous copy of a transposed view (see transposed_source); None if not applicable."""
    info = transposed_source(arr)
    if info is None:
        return None
    batch, rows, cols = info                                       # source (batch, rows, cols)
    item = arr.dtype.itemsize
    if out is None:
        out = DeviceArray.empty(arr.shape, arr.dtype, arr.dev if arr.dev >= 0 else None)
    key = ("transpose", item)
    tc = 64 if item == 4 else 32
    kern = get_kernel(key, lambda name: _TRANSPOSE_SRC.replace("NAME", name).replace("TCVAL", str(tc))
                      .replace("W", _WORD[item]))
    tiles_x, tiles_y = -(-cols // tc), -(-rows // 32)
    ntiles = batch * tiles_x * tiles_y
    a = Args()
    a.ptr(arr.ptr); a.ptr(out.ptr); a.i64(rows); a.i64(cols); a.i64(arr.strides[-1] // item); a.i64(rows)
    a.i64(rows * cols); a.i64(rows * cols); a.i64(tiles_x); a.i64(tiles_y); a.i64(ntiles)
    launch(kern, arr.dev, max(1, min(ntiles, 148 * 16)), 256, a)
    return out
