/* libdrcuda -- the C-ABI boundary of delayrepay_b200.
 *
 * The reference (magnusmorton/DelayRepay) has no FFI of its own: everything below its
 * backend-module protocol lives in CuPy.  Each entry point here names the reference
 * interface (file:line under /root/reference) whose CuPy service it replaces.
 *
 * Conventions
 *   - plain C types only; device pointers travel as uint64_t; handles are opaque uint64_t;
 *   - every call returns int: 0 = ok, non-zero = failure; drc_last_error() returns the
 *     thread-local message of the last failure on the calling thread;
 *   - all work is stream-ordered; each device owns DRC_NUM_STREAMS streams, stream 0 is the
 *     compute stream (the reference uses CuPy's single default stream, cuda.py:96);
 *   - the library loads without a GPU (libcuda / libnccl are resolved in drc_init /
 *     drc_nccl_*); drc_compile works without a GPU (NVRTC cross-compiles sm_100a);
 *   - there is NO CPU fallback: without a device every device call fails with an error.
 */
#ifndef DRCUDA_H
#define DRCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRC_NUM_STREAMS 4
#define DRC_ABI_VERSION 1

/* ---- lifetime ------------------------------------------------------------------------
 * Replaces the implicit `import cupy` device bring-up (cuda.py:17). */
int drc_abi_version(void);
const char* drc_last_error(void);
int drc_init(void);                              /* cuInit, primary contexts, streams, pools */
int drc_shutdown(void);
int drc_device_count(int* count);
int drc_device_attr(int dev, int* sm_count, int* cc_major, int* cc_minor,
                    size_t* total_mem, int* l2_bytes, int* max_smem_optin);
int drc_device_name(int dev, char* buf, size_t buflen);
/* PCI address ("0000:1b:00.0"): lets the host side pin its threads and its pinned staging
 * memory to the GPU's NUMA node (/sys/bus/pci/devices/<id>/local_cpulist). */
int drc_device_pci_bus_id(int dev, char* buf, size_t buflen);
int drc_mem_info(int dev, size_t* free_bytes, size_t* total_bytes);

/* ---- memory: stream-ordered pool -----------------------------------------------------
 * Replaces CuPy's memory pool behind kernel outputs (cuda.py:96 allocates `out`) and behind
 * the creation functions (delayarray.py:596-644). */
int drc_malloc_async(int dev, int stream, size_t bytes, uint64_t* dptr);
int drc_free_async(int dev, int stream, uint64_t dptr);
int drc_pool_trim(int dev, size_t keep_bytes);
int drc_memset_async(int dev, int stream, uint64_t dptr, int byte_value, size_t bytes);

/* Replaces cupy.ndarray.get() / cupy.asarray H2D (delayarray.py:101-106, :620). */
int drc_memcpy_h2d_async(int dev, int stream, uint64_t dst, const void* src, size_t bytes);
int drc_memcpy_d2h_async(int dev, int stream, void* dst, uint64_t src, size_t bytes);
int drc_memcpy_d2d_async(int dev, int stream, uint64_t dst, uint64_t src, size_t bytes);
/* NVLink peer copy (new: leading-axis sharding halo exchange; no reference equivalent). */
int drc_memcpy_peer_async(int dst_dev, uint64_t dst, int src_dev, uint64_t src, size_t bytes,
                          int stream_dev, int stream);
int drc_enable_peer_access(int dev, int peer);
/* Peer-visible device memory for sharded arrays (new; north_star item 4): plain cuMemAlloc
 * allocations, reachable from the other GPUs of the box either through peer access (one process
 * driving several devices) or through a 64-byte IPC handle opened by the neighbour rank's process
 * (one process per GPU).  The stencil kernel stores halo rows and release-flags straight into
 * memory obtained this way. */
#define DRC_IPC_HANDLE_BYTES 64
int drc_peer_alloc(int dev, size_t bytes, uint64_t* dptr);
int drc_peer_free(int dev, uint64_t dptr);
int drc_ipc_get_handle(int dev, uint64_t dptr, void* handle64);
int drc_ipc_open_handle(int dev, const void* handle64, uint64_t* dptr);
int drc_ipc_close_handle(int dev, uint64_t dptr);
/* Pinned host staging for the e2e path. */
int drc_host_alloc(size_t bytes, void** hptr);
int drc_host_free(void* hptr);
int drc_host_register(void* hptr, size_t bytes);
int drc_host_unregister(void* hptr);

/* ---- compile + load: NVRTC -> sm_100a cubin ------------------------------------------
 * Replaces cupy.ElementwiseKernel's NVRTC compile + source-keyed cache (cuda.py:35-43). */
int drc_compile(const char* source, const char* name, const char* const* options,
                int num_options, void** cubin, size_t* cubin_len, char** log);
int drc_free_blob(void* blob);
/* Version and path of the NVRTC the library bound (dlopen by absolute path, toolkit first: the
 * first libnvrtc.so.12 a process happens to load is not necessarily the toolkit's). */
int drc_nvrtc_version(int* major, int* minor, const char** path);
int drc_module_load(int dev, const void* cubin, size_t cubin_len, uint64_t* module);
int drc_module_unload(int dev, uint64_t module);
int drc_module_get_function(int dev, uint64_t module, const char* entry, uint64_t* func);
int drc_func_set_max_dynamic_smem(int dev, uint64_t func, int bytes);
int drc_func_attrs(int dev, uint64_t func, int* num_regs, int* static_smem, int* local_bytes,
                   int* max_threads);
int drc_occupancy(int dev, uint64_t func, int block_threads, size_t dyn_smem, int* blocks_per_sm);

/* ---- launch ----------------------------------------------------------------------------
 * Replaces `kern(*inputs)` (cuda.py:95-96).  `args` is an array of num_args pointers to the
 * argument values (cuLaunchKernel convention).  drc_launch_packed takes one contiguous blob
 * plus the byte offset of every argument inside it (one ctypes call, no per-arg objects). */
int drc_launch(int dev, int stream, uint64_t func, const uint32_t grid[3],
               const uint32_t block[3], uint32_t dyn_smem, void** args, int num_args);
int drc_launch_packed(int dev, int stream, uint64_t func, uint32_t gx, uint32_t gy, uint32_t gz,
                      uint32_t bx, uint32_t by, uint32_t bz, uint32_t dyn_smem,
                      uint32_t cluster_x, const void* blob, const uint32_t* offsets,
                      int num_args);
uint64_t drc_launch_count(void);                  /* kernels launched since drc_init */

/* TMA descriptor for slice-stencil tiles (new; the reference reads views through CuPy's
 * strided indexer, delayarray.py:123-128).  Writes a 128-byte CUtensorMap into `out128`.
 * dims/strides are innermost-first; strides in BYTES for dims 1..rank-1. */
int drc_tensormap_encode(int dev, void* out128, int dtype_code, uint32_t rank, uint64_t gptr,
                         const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle, int l2_promotion);

/* ---- ordering / timing ---------------------------------------------------------------- */
int drc_stream_sync(int dev, int stream);
int drc_device_sync(int dev);
int drc_event_create(int dev, uint64_t* event);
int drc_event_destroy(int dev, uint64_t event);
int drc_event_record(int dev, int stream, uint64_t event);
int drc_event_sync(int dev, uint64_t event);
int drc_event_elapsed_ms(int dev, uint64_t start, uint64_t stop, float* ms);
int drc_stream_wait_event(int dev, int stream, uint64_t event);
int drc_stream_handle(int dev, int stream, uint64_t* custream);

/* ---- collectives: NCCL over NVLink 5 / NVSwitch (new; the reference has none) ---------
 * One rank per process (torchrun) via unique-id, or all local devices in one process. */
#define DRC_NCCL_UNIQUE_ID_BYTES 128
int drc_nccl_available(void);
int drc_nccl_get_unique_id(void* id128);
int drc_nccl_init_rank(int dev, int nranks, int rank, const void* id128, uint64_t* comm);
int drc_nccl_init_all(int ndev, const int* devs, uint64_t* comms);
int drc_nccl_destroy(uint64_t comm);
/* dtype_code: 0=f32 1=f64 2=i32 3=i64 4=u8 ; op: 0=sum 1=prod 2=max 3=min */
int drc_nccl_allreduce(uint64_t comm, int dev, int stream, uint64_t sendbuf, uint64_t recvbuf,
                       size_t count, int dtype_code, int op);
int drc_nccl_sendrecv(uint64_t comm, int dev, int stream, uint64_t sendbuf, size_t send_bytes,
                      int send_peer, uint64_t recvbuf, size_t recv_bytes, int recv_peer);
int drc_nccl_allgather(uint64_t comm, int dev, int stream, uint64_t sendbuf, uint64_t recvbuf,
                       size_t bytes_per_rank);
int drc_nccl_group_start(void);
int drc_nccl_group_end(void);

/* ---- FFT: cuFFT behind the C ABI -------------------------------------------------------
 * Replaces `cupy.fft.fft` (reference fft.py:12).  Batched 1-d complex-to-complex transform of
 * `batch` contiguous rows of length n (in place when in == out); plans are cached per
 * (device, n, batch, precision).  libcufft is resolved at first use. */
int drc_fft_c2c_1d(int dev, int stream, uint64_t in, uint64_t out, int n, int batch,
                   int is_double, int inverse);

#ifdef __cplusplus
}
#endif
#endif /* DRCUDA_H */
