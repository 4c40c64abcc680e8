"""Device kernels beside the fused-region generator: prefix sums (np.cumsum), the Philox counter
RNG behind delayrepay.random, and the real->complex packing used by fft.  Each is a hand-written
CUDA C++ template specialised by dtype and compiled through the same NVRTC/cubin cache
(engine.get_kernel).  Reference: these are the eager CuPy calls of delayarray.py:555-558 (cumsum),
random.py:8-13 (cuRAND) and fft.py:12 (cuFFT).
"""
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
// 1-d, three phases: (1) per-block totals, (2) exclusive scan of the totals by one block,
// (3) per-block inclusive scan (warp shuffles + shared memory) offset by its prefix
#define ITEMS 8
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


def _scan_kernels(in_dt, acc_dt):
    # one module holds all four entry points; load each by name
    key = ("scan", np.dtype(in_dt).str, np.dtype(acc_dt).str)
    name = engine.kernel_name(key)
    src = _SHFL_UP + _SCAN_SRC.replace("NAME", name).replace("TIN", ctype(in_dt)).replace("TACC", ctype(acc_dt))
    _, cubin = engine.compile_source(name, src)
    out = {}
    for suffix in ("serial", "partials", "offsets", "final"):
        k = engine._kernels.get(key + (suffix,))
        if k is None:
            k = engine._kernels[key + (suffix,)] = engine.Kernel(f"{name}_{suffix}", src, cubin, {})
        out[suffix] = k
    return out


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
