"""Dense float32 `A @ B` on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), used when a
fused region ends in a genuine dense contraction (both dimensions of the result large).

Precision: NumPy's float32 matmul is an fp32 SGEMM; plain TF32 (10-bit mantissa) would miss the
rtol 1e-5 bar by two orders of magnitude.  The operands are therefore split once, in a fused
elementwise pre-pass that also evaluates their lazy producers:  hi = x rounded to TF32 (10-bit
mantissa), lo = x - hi (exact in fp32) rounded to TF32 as well.  The kernel
accumulates  A_hi B_hi + A_hi B_lo + A_lo B_hi  in one fp32 TMEM accumulator ("3xTF32", relative
error ~2^-21).  B is written transposed by the pre-pass so that both operands are K-major.

Kernel (one CTA per 128 x 128 output tile, 192 threads, warp-specialised):
  warp 0  lane 0   TMA producer: 4 x (128 x 32 fp32, SWIZZLE_128B) boxes per k-block into a 3-stage ring
  warp 1  lane 0   MMA issuer: 12 tcgen05.mma.kind::tf32 (M128 N128 K8) per k-block; tcgen05.commit
                   releases the stage / signals the epilogue
  warps 2-5        epilogue: tcgen05.ld 32x32b.x32 -> registers -> 128-bit global stores
"""
import ctypes as C

import numpy as np

from . import engine, planner
from ._lib import check, lib
from .device import DeviceArray
from .engine import Args, get_kernel, launch

BM, BN, BK, NS = 128, 128, 32, 3
TILE_BYTES = BM * BK * 4
SMEM_BYTES = NS * 4 * TILE_BYTES + 1024

_SRC = r'''
#define BM 128
#define BN 128
#define BK 32
#define NS 3
#define TILE_BYTES (BM * BK * 4)
extern "C" __global__ void __launch_bounds__(192, 1) NAME(
    const __grid_constant__ DrTensorMap map_ah, const __grid_constant__ DrTensorMap map_al,
    const __grid_constant__ DrTensorMap map_bh, const __grid_constant__ DrTensorMap map_bl,
    float* __restrict__ Cmat, int M, int N, int K, i64 ldc) {
  extern __shared__ unsigned char dr_smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[NS], empty_bar[NS], accum_bar;
  __shared__ unsigned tmem_slot;
  unsigned char* smem = dr_smem_raw + ((1024u - (dr_smem_addr(dr_smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { dr_mbar_init(&full_bar[s], 1); dr_mbar_init(&empty_bar[s], 1); }
    dr_mbar_init(&accum_bar, 1);
    dr_fence_barrier_init();
  }
  if (warp == 1) dr_tmem_alloc(&tmem_slot, 128);
  dr_tc_fence_before();
  __syncthreads();
  dr_tc_fence_after();
  const unsigned tmem_acc = tmem_slot;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = (K + BK - 1) / BK;
  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NS;
        dr_mbar_wait(&empty_bar[s], ((kb / NS) & 1) ^ 1);
        dr_mbar_expect_tx(&full_bar[s], 4 * TILE_BYTES);
        unsigned char* st = smem + s * 4 * TILE_BYTES;
        dr_tma_load_2d(st + 0 * TILE_BYTES, &map_ah, kb * BK, m0, &full_bar[s]);
        dr_tma_load_2d(st + 1 * TILE_BYTES, &map_al, kb * BK, m0, &full_bar[s]);
        dr_tma_load_2d(st + 2 * TILE_BYTES, &map_bh, kb * BK, n0, &full_bar[s]);
        dr_tma_load_2d(st + 3 * TILE_BYTES, &map_bl, kb * BK, n0, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BN >> 3) << 17)
                           | ((unsigned)(BM >> 4) << 24);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NS;
        dr_mbar_wait(&full_bar[s], (kb / NS) & 1);
        dr_tc_fence_after();
        const unsigned base = dr_smem_addr(smem + s * 4 * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const unsigned long long ah = dr_umma_desc(base + 0 * TILE_BYTES + k * 32);
          const unsigned long long al = dr_umma_desc(base + 1 * TILE_BYTES + k * 32);
          const unsigned long long bh = dr_umma_desc(base + 2 * TILE_BYTES + k * 32);
          const unsigned long long bl = dr_umma_desc(base + 3 * TILE_BYTES + k * 32);
          dr_umma_tf32(tmem_acc, ah, bh, idesc, (kb | k) != 0);
          dr_umma_tf32(tmem_acc, ah, bl, idesc, 1);
          dr_umma_tf32(tmem_acc, al, bh, idesc, 1);
        }
        dr_umma_commit(&empty_bar[s]);
      }
      dr_umma_commit(&accum_bar);
    }
  } else {
    dr_mbar_wait(&accum_bar, 0);
    dr_tc_fence_after();
    const int q = warp & 3;                           // this warp's quarter of the 128 TMEM lanes
    const i64 row = (i64)m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      unsigned r[32];
      dr_tmem_ld32(tmem_acc + ((unsigned)(q * 32) << 16) + (unsigned)(c * 32), r);
      const int col0 = n0 + c * 32;
      if (row < M) {
        float* dst = Cmat + row * ldc + col0;
        if (col0 + 32 <= N && (ldc & 3) == 0) {
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<uint4*>(dst + 4 * v) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
        } else {
          for (int v = 0; v < 32; ++v) if (col0 + v < N) dst[v] = __uint_as_float(r[v]);
        }
      }
    }
  }
  dr_tc_fence_before();
  __syncthreads();
  if (warp == 1) dr_tmem_dealloc(tmem_acc, 128);
}
'''


def _tensor_map(dev, arr, rows, cols):
    """(rows, cols) row-major fp32 -> box (BK, 128), SWIZZLE_128B."""
    return engine.encode_tensormap(dev, "float32", arr.ptr, (cols, rows), (arr.strides[0],),
                                   (BK, BM), swizzle=3)


def split_operands(a_node, b_node):
    """Fused pre-pass: evaluate both lazy operands once and write A_hi, A_lo (M, K) and
    B_hi^T, B_lo^T (N, K): hi = TF32-exact part, lo = exact remainder."""
    from .delayarray import BinaryNumpyEx, RawOp
    m, k = a_node.shape
    n = b_node.shape[1]
    f32 = np.dtype(np.float32)
    from .delayarray import as_dtype
    a_node, b_node = as_dtype(a_node, f32), as_dtype(b_node, f32)
    # hi = x rounded to TF32, lo = (x - hi) rounded to TF32: the tensor core then has nothing
    # left to truncate, so no biased error accumulates along K
    a_hi = RawOp("tf32_hi", a_node)
    a_lo = RawOp("tf32_hi", BinaryNumpyEx(np.subtract, a_node, a_hi))
    b_hi = RawOp("tf32_hi", b_node)
    b_lo = RawOp("tf32_hi", BinaryNumpyEx(np.subtract, b_node, b_hi))
    from .device import current_device
    dev = -1 if engine.is_dry() else current_device()
    kp = -(-k // 4) * 4                 # TMA wants row pitches that are multiples of 16 bytes
    # operands with fewer rows than one tile are padded with zero rows, so that a TMA box never
    # exceeds the extent of its tensor map
    bufs = [DeviceArray.empty((max(rows, BM), kp), f32, dev if dev >= 0 else None) for rows in (m, m, n, n)]
    for buf, rows in zip(bufs, (m, m, n, n)):
        if rows < BM and dev >= 0:
            buf.fill(0)
    ah, al, bth, btl = (buf[:rows, :k] for buf, rows in zip(bufs, (m, m, n, n)))
    engine.evaluate_nodes([a_hi, a_lo], outs=[ah, al])
    from . import extras
    from .delayarray import NPArray
    b_t = None
    if b_node.dtype == f32 and min(k, n) >= 32:
        # transpose B once through shared-memory tiles (coalesced both ways), then split B^T
        # with contiguous stores; writing hi/lo straight into transposed outputs stores 4-byte
        # words a row pitch apart
        b_dev = b_node._force()
        b_dev = b_dev if b_dev.is_contiguous else b_dev.copy()
        b_t = extras.transpose_copy(b_dev.T)
    if b_t is not None:
        bt_node = NPArray(b_t)
        bt_hi = RawOp("tf32_hi", bt_node)
        bt_lo = RawOp("tf32_hi", BinaryNumpyEx(np.subtract, bt_node, bt_hi))
        engine.evaluate_nodes([bt_hi, bt_lo], outs=[bth, btl])
    else:
        engine.evaluate_nodes([b_hi, b_lo], outs=[bth.T, btl.T])
    return ah, al, bth, btl


K_CHUNK = 8192      # the tensor core's fp32 accumulation error grows ~sqrt(K): 3.8e-6 of scale at 4096


def matmul_tf32x3(a_node, b_node):
    m, k = a_node.shape
    n = b_node.shape[1]
    if k > K_CHUNK:
        # long contractions: accumulate K-chunks in separate passes and add them in fp32
        from .delayarray import NPArray
        a_dev, b_dev = a_node._force(), b_node._force()
        total = None
        for lo in range(0, k, K_CHUNK):
            part = NPArray(matmul_tf32x3(NPArray(a_dev[:, lo:lo + K_CHUNK]), NPArray(b_dev[lo:lo + K_CHUNK, :])))
            total = part if total is None else total + part
        return total._force()
    ah, al, bth, btl = split_operands(a_node, b_node)
    dev = ah.dev
    out = DeviceArray.empty((m, n), np.float32, dev if dev >= 0 else None)
    kern = get_kernel(("tcgen05_gemm", BM, BN, BK, NS), lambda name: _SRC.replace("NAME", name))
    a = Args()
    for arr, rows in ((ah, m), (al, m), (bth, n), (btl, n)):
        a.raw(_tensor_map(dev, arr, max(rows, BM), k), 64)
    a.ptr(out.ptr)
    for v in (m, n, k):
        a.scalar(v, np.int32)
    a.i64(n)
    if dev >= 0 and not kern.meta.get("smem_set"):
        check(lib.drc_func_set_max_dynamic_smem(dev, kern.func(dev), SMEM_BYTES))
        kern.meta["smem_set"] = True
    launch(kern, dev, (-(-n // BN), -(-m // BM), 1), 192, a, smem=SMEM_BYTES)
    return out


# --------------------------------------------------------------------------- generic tiled GEMM
# float64 and integer A @ B have no tcgen05 path (the tensor cores take TF32 / bf16 / fp8 here).
# Round 1 sent them to the `cols` reduction kernel: one thread per output element, no reuse --
# 16384 x 16384 @ 16384 x 64 float64 took 20.4 ms (105 GB/s, profiles/r1_surface_ops_after.txt).
# This is the classic shared-memory / register-tiled kernel: a 64 x 64 output tile per CTA of 256
# threads, BK = 16 deep slices of A and B staged in shared memory (A transposed on the way in, so
# both are read conflict-free), a 4 x 4 accumulator tile per thread.  Accumulation is in the result
# type, one term at a time in k order (float64: same order as the reference's BLAS is not defined;
# the parity bar is rtol 1e-12 of the terms' scale; integers: exact, wrap-around like NumPy).
_TILED_SRC = r"""
extern "C" __global__ void __launch_bounds__(256) NAME(const T* __restrict__ A, const T* __restrict__ B,
    T* __restrict__ Cout, int M, int N, int K, i64 lda, i64 ldb, i64 ldc) {
  __shared__ T As[16][64 + PAD];          // As[k][m]
  __shared__ T Bs[16][64 + PAD];          // Bs[k][n]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  ACC acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (ACC)0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    // A tile 64 x 16: thread -> (row = tid / 4 ... ), 4 elements each; coalesced along k
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + r * 256, mm = e >> 4, kk = e & 15;
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[(i64)gm * lda + gk] : (T)0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + r * 256, kk = e >> 6, nn = e & 63;
      const int gk = k0 + kk, gn = n0 + nn;
      Bs[kk][nn] = (gk < K && gn < N) ? B[(i64)gk * ldb + gn] : (T)0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      ACC a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = (ACC)As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = (ACC)Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = MAC(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) Cout[(i64)gm * ldc + gn] = (T)acc[i][j];
    }
  }
}
"""


def matmul_tiled(a_node, b_node, res_dt):
    """A(M,K) @ B(K,N) in `res_dt` (float64 / float32 without a tensor path / integers) on the
    register-tiled kernel.  Operands are materialised contiguously in the result type first (their
    elementwise producers are fused into that pass)."""
    from .delayarray import as_dtype
    from .codegen import ctype
    res_dt = np.dtype(res_dt)
    a_dev = as_dtype(a_node, res_dt)._force()
    b_dev = as_dtype(b_node, res_dt)._force()
    if not a_dev.is_contiguous:
        a_dev = a_dev.copy()
    if not b_dev.is_contiguous:
        b_dev = b_dev.copy()
    m, k = a_dev.shape
    n = b_dev.shape[1]
    dev = a_dev.dev
    out = DeviceArray.empty((m, n), res_dt, dev if dev >= 0 else None)
    T = ctype(res_dt)
    if res_dt.kind == "f":
        acc, mac = T, "fma"
    elif res_dt.kind == "b":
        raise TypeError("boolean matmul is not supported")
    else:
        # signed overflow is undefined in C++ (and NVRTC uses that): accumulate unsigned, like the
        # elementwise integer path
        acc = {1: "unsigned char", 2: "unsigned short", 4: "unsigned int", 8: "unsigned long long"}[res_dt.itemsize]
        mac = "DR_IMAC"
    src = ("#define DR_IMAC(a, b, c) ((a) * (b) + (c))\n" if mac == "DR_IMAC" else "") + \
        _TILED_SRC.replace("NAME", "KNAME").replace("ACC", acc).replace("MAC", mac) \
        .replace("PAD", "4" if res_dt.itemsize == 4 else "2").replace(" T ", f" {T} ") \
        .replace("const T*", f"const {T}*").replace("(T)", f"({T})").replace("T* __restrict__ Cout", f"{T}* __restrict__ Cout")
    kern = get_kernel(("gemm_tiled", res_dt.str), lambda name: src.replace("KNAME", name))
    a = Args()
    a.ptr(a_dev.ptr)
    a.ptr(b_dev.ptr)
    a.ptr(out.ptr)
    for v in (m, n, k):
        a.scalar(v, np.int32)
    a.i64(k)
    a.i64(n)
    a.i64(n)
    launch(kern, dev, (-(-n // 64), -(-m // 64), 1), 256, a)
    return out
