// libdrcuda: thin C-ABI runtime under delayrepay_b200 (see include/drcuda.h).
//
// CUDA *driver* API only (resolved with dlopen so the library loads on a box without a GPU),
// NVRTC linked directly (compiles sm_100a cubins without a GPU), NCCL resolved lazily.
// No cudart, no CuPy, no PyTorch types.  One primary context, DRC_NUM_STREAMS streams and the
// device's default stream-ordered memory pool per device; devices are brought up on first use
// so that one-process-per-GPU launches (torchrun) never touch the other seven devices.
#include "drcuda.h"

#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvrtc.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

// ---------------------------------------------------------------- driver entry points
#define DRC_XSTR(x) DRC_STR(x)
#define DRC_STR(x) #x
#define DRV_FUNCS(X)                                                                         \
  X(cuInit) X(cuGetErrorString) X(cuGetErrorName) X(cuDeviceGetCount) X(cuDeviceGet)         \
  X(cuDeviceGetName) X(cuDeviceGetPCIBusId) X(cuDeviceGetAttribute) X(cuDeviceTotalMem) X(cuDevicePrimaryCtxRetain) \
  X(cuDevicePrimaryCtxRelease) X(cuCtxSetCurrent) X(cuCtxGetCurrent) X(cuStreamCreate)       \
  X(cuStreamDestroy) X(cuStreamSynchronize) X(cuCtxSynchronize) X(cuMemGetInfo)              \
  X(cuMemAllocAsync) X(cuMemFreeAsync) X(cuDeviceGetDefaultMemPool) X(cuMemPoolSetAttribute) \
  X(cuMemPoolTrimTo) X(cuMemsetD8Async) X(cuMemcpyHtoDAsync) X(cuMemcpyDtoHAsync)            \
  X(cuMemcpyDtoDAsync) X(cuMemcpyPeerAsync) X(cuCtxEnablePeerAccess)                         \
  X(cuMemAlloc) X(cuMemFree) X(cuIpcGetMemHandle) X(cuIpcOpenMemHandle) X(cuIpcCloseMemHandle) \
  X(cuDeviceCanAccessPeer) X(cuMemHostAlloc) X(cuMemFreeHost) X(cuMemHostRegister)           \
  X(cuMemHostUnregister) X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction)        \
  X(cuFuncSetAttribute) X(cuFuncGetAttribute) X(cuOccupancyMaxActiveBlocksPerMultiprocessor) \
  X(cuLaunchKernel) X(cuLaunchKernelEx) X(cuTensorMapEncodeTiled) X(cuEventCreate)           \
  X(cuEventDestroy) X(cuEventRecord) X(cuEventSynchronize) X(cuEventElapsedTime)             \
  X(cuStreamWaitEvent)

#define DECL(fn) decltype(&fn) p_##fn = nullptr;
DRV_FUNCS(DECL)
#undef DECL

void* g_libcuda = nullptr;
std::mutex g_mu;
bool g_inited = false;
int g_ndev = 0;
std::atomic<uint64_t> g_launches{0};

struct Dev {
  bool up = false;
  CUdevice dev = 0;
  CUcontext ctx = nullptr;
  CUstream streams[DRC_NUM_STREAMS] = {};
  CUmemoryPool pool = nullptr;
};
std::vector<Dev> g_devs;
int g_cur_dev = 0;   // last device used; host allocations attach to it

int cu_fail(CUresult r, const char* what) {
  const char* name = nullptr;
  const char* msg = nullptr;
  if (p_cuGetErrorName) p_cuGetErrorName(r, &name);
  if (p_cuGetErrorString) p_cuGetErrorString(r, &msg);
  return fail("%s failed: %s (%d): %s", what, name ? name : "?", (int)r, msg ? msg : "?");
}
#define CU(call)                                 \
  do {                                           \
    CUresult _r = (call);                        \
    if (_r != CUDA_SUCCESS) return cu_fail(_r, #call); \
  } while (0)

int load_driver() {
  if (g_libcuda) return 0;
  g_libcuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!g_libcuda)
    return fail("libcuda.so.1 not found (%s): no NVIDIA driver on this machine; "
                "delayrepay_b200 has no CPU fallback", dlerror());
#define LOAD(fn)                                                         \
  p_##fn = (decltype(&fn))dlsym(g_libcuda, DRC_XSTR(fn));                \
  if (!p_##fn) return fail("libcuda.so.1 lacks symbol %s", DRC_XSTR(fn));
  DRV_FUNCS(LOAD)
#undef LOAD
  return 0;
}

// Bring device `dev` up (context, streams, pool) and make its context current.
int use(int dev) {
  if (!g_inited) return fail("drc_init() has not been called (or failed)");
  if (dev < 0 || dev >= g_ndev) return fail("device %d out of range (have %d)", dev, g_ndev);
  Dev& d = g_devs[dev];
  if (!d.up) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!d.up) {
      CU(p_cuDeviceGet(&d.dev, dev));
      CU(p_cuDevicePrimaryCtxRetain(&d.ctx, d.dev));
      CU(p_cuCtxSetCurrent(d.ctx));
      for (int s = 0; s < DRC_NUM_STREAMS; ++s)
        CU(p_cuStreamCreate(&d.streams[s], CU_STREAM_NON_BLOCKING));
      CU(p_cuDeviceGetDefaultMemPool(&d.pool, d.dev));
      cuuint64_t keep = ~0ull;   // never hand memory back to the OS between regions
      CU(p_cuMemPoolSetAttribute(d.pool, CU_MEMPOOL_ATTR_RELEASE_THRESHOLD, &keep));
      d.up = true;
    }
  }
  CUcontext cur = nullptr;
  p_cuCtxGetCurrent(&cur);
  if (cur != d.ctx) CU(p_cuCtxSetCurrent(d.ctx));
  g_cur_dev = dev;
  return 0;
}

inline int stream_of(int dev, int stream, CUstream* out) {
  if (stream < 0 || stream >= DRC_NUM_STREAMS) return fail("stream %d out of range", stream);
  *out = g_devs[dev].streams[stream];
  return 0;
}
#define USE(dev) do { if (int _e = use(dev)) return _e; } while (0)
#define STREAM(dev, s, var) CUstream var = nullptr; do { if (int _e = stream_of(dev, s, &var)) return _e; } while (0)

// ---------------------------------------------------------------- NCCL entry points
void* g_libnccl = nullptr;
#define NCCL_FUNCS(X)                                                                       \
  X(ncclGetUniqueId) X(ncclCommInitRank) X(ncclCommInitAll) X(ncclCommDestroy)              \
  X(ncclAllReduce) X(ncclSend) X(ncclRecv) X(ncclAllGather) X(ncclGroupStart)               \
  X(ncclGroupEnd) X(ncclGetErrorString)
#define DECL(fn) decltype(&fn) p_##fn = nullptr;
NCCL_FUNCS(DECL)
#undef DECL

int load_nccl() {
  if (g_libnccl) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_libnccl) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail("libnccl.so.2 not found: %s", dlerror());
#define LOAD(fn)                                        \
  p_##fn = (decltype(&fn))dlsym(h, #fn);                \
  if (!p_##fn) return fail("libnccl.so.2 lacks symbol %s", #fn);
  NCCL_FUNCS(LOAD)
#undef LOAD
  g_libnccl = h;
  return 0;
}
#define NC(call)                                                                    \
  do {                                                                              \
    ncclResult_t _r = (call);                                                       \
    if (_r != ncclSuccess) return fail("%s failed: %s", #call, p_ncclGetErrorString(_r)); \
  } while (0)

int nccl_dtype(int code, ncclDataType_t* t) {
  switch (code) {
    case 0: *t = ncclFloat32; return 0;
    case 1: *t = ncclFloat64; return 0;
    case 2: *t = ncclInt32; return 0;
    case 3: *t = ncclInt64; return 0;
    case 4: *t = ncclUint8; return 0;
  }
  return fail("bad dtype code %d", code);
}

}  // namespace

extern "C" {

int drc_abi_version(void) { return DRC_ABI_VERSION; }
const char* drc_last_error(void) { return g_err.c_str(); }

int drc_init(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_inited) return 0;
  if (int e = load_driver()) return e;
  CU(p_cuInit(0));
  CU(p_cuDeviceGetCount(&g_ndev));
  if (g_ndev <= 0) return fail("no CUDA device visible");
  g_devs.assign(g_ndev, Dev());
  g_inited = true;
  return 0;
}

int drc_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_inited) return 0;
  for (int i = 0; i < g_ndev; ++i) {
    Dev& d = g_devs[i];
    if (!d.up) continue;
    p_cuCtxSetCurrent(d.ctx);
    p_cuCtxSynchronize();
    for (auto& s : d.streams) if (s) p_cuStreamDestroy(s);
    p_cuDevicePrimaryCtxRelease(d.dev);
    d = Dev();
  }
  g_inited = false;
  return 0;
}

int drc_device_count(int* count) {
  if (!g_inited) return fail("drc_init() has not been called (or failed)");
  *count = g_ndev;
  return 0;
}

int drc_device_attr(int dev, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem,
                    int* l2_bytes, int* max_smem_optin) {
  USE(dev);
  CUdevice d = g_devs[dev].dev;
  CU(p_cuDeviceGetAttribute(sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, d));
  CU(p_cuDeviceGetAttribute(cc_major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, d));
  CU(p_cuDeviceGetAttribute(cc_minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, d));
  CU(p_cuDeviceGetAttribute(l2_bytes, CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE, d));
  CU(p_cuDeviceGetAttribute(max_smem_optin, CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN, d));
  CU(p_cuDeviceTotalMem(total_mem, d));
  return 0;
}

int drc_device_name(int dev, char* buf, size_t buflen) {
  USE(dev);
  CU(p_cuDeviceGetName(buf, (int)buflen, g_devs[dev].dev));
  return 0;
}

// "0000:1b:00.0"-style PCI address: the key to /sys/bus/pci/devices/<id>/numa_node, local_cpulist
int drc_device_pci_bus_id(int dev, char* buf, size_t buflen) {
  USE(dev);
  CU(p_cuDeviceGetPCIBusId(buf, (int)buflen, g_devs[dev].dev));
  return 0;
}

int drc_mem_info(int dev, size_t* free_bytes, size_t* total_bytes) {
  USE(dev);
  CU(p_cuMemGetInfo(free_bytes, total_bytes));
  return 0;
}

// ---------------------------------------------------------------- memory
int drc_malloc_async(int dev, int stream, size_t bytes, uint64_t* dptr) {
  USE(dev);
  STREAM(dev, stream, s);
  CUdeviceptr p = 0;
  CU(p_cuMemAllocAsync(&p, bytes ? bytes : 1, s));
  *dptr = (uint64_t)p;
  return 0;
}

int drc_free_async(int dev, int stream, uint64_t dptr) {
  USE(dev);
  STREAM(dev, stream, s);
  CU(p_cuMemFreeAsync((CUdeviceptr)dptr, s));
  return 0;
}

int drc_pool_trim(int dev, size_t keep_bytes) {
  USE(dev);
  CU(p_cuMemPoolTrimTo(g_devs[dev].pool, keep_bytes));
  return 0;
}

int drc_memset_async(int dev, int stream, uint64_t dptr, int byte_value, size_t bytes) {
  USE(dev);
  STREAM(dev, stream, s);
  CU(p_cuMemsetD8Async((CUdeviceptr)dptr, (unsigned char)byte_value, bytes, s));
  return 0;
}

int drc_memcpy_h2d_async(int dev, int stream, uint64_t dst, const void* src, size_t bytes) {
  USE(dev);
  STREAM(dev, stream, s);
  CU(p_cuMemcpyHtoDAsync((CUdeviceptr)dst, src, bytes, s));
  return 0;
}

int drc_memcpy_d2h_async(int dev, int stream, void* dst, uint64_t src, size_t bytes) {
  USE(dev);
  STREAM(dev, stream, s);
  CU(p_cuMemcpyDtoHAsync(dst, (CUdeviceptr)src, bytes, s));
  return 0;
}

int drc_memcpy_d2d_async(int dev, int stream, uint64_t dst, uint64_t src, size_t bytes) {
  USE(dev);
  STREAM(dev, stream, s);
  CU(p_cuMemcpyDtoDAsync((CUdeviceptr)dst, (CUdeviceptr)src, bytes, s));
  return 0;
}

int drc_memcpy_peer_async(int dst_dev, uint64_t dst, int src_dev, uint64_t src, size_t bytes,
                          int stream_dev, int stream) {
  USE(dst_dev);
  USE(src_dev);
  USE(stream_dev);
  STREAM(stream_dev, stream, s);
  CU(p_cuMemcpyPeerAsync((CUdeviceptr)dst, g_devs[dst_dev].ctx, (CUdeviceptr)src,
                         g_devs[src_dev].ctx, bytes, s));
  return 0;
}

int drc_enable_peer_access(int dev, int peer) {
  USE(peer);
  USE(dev);
  int can = 0;
  CU(p_cuDeviceCanAccessPeer(&can, g_devs[dev].dev, g_devs[peer].dev));
  if (!can) return fail("device %d cannot access peer %d", dev, peer);
  CUresult r = p_cuCtxEnablePeerAccess(g_devs[peer].ctx, 0);
  if (r != CUDA_SUCCESS && r != CUDA_ERROR_PEER_ACCESS_ALREADY_ENABLED)
    return cu_fail(r, "cuCtxEnablePeerAccess");
  return 0;
}

// ---- peer-visible allocations (sharded arrays).  cuMemAlloc memory -- unlike stream-ordered
// pool memory -- is covered by cuCtxEnablePeerAccess inside one process and can be exported to
// the other ranks' processes with the legacy IPC handles.
int drc_peer_alloc(int dev, size_t bytes, uint64_t* dptr) {
  USE(dev);
  CUdeviceptr p = 0;
  CU(p_cuMemAlloc(&p, bytes ? bytes : 1));
  *dptr = (uint64_t)p;
  return 0;
}

int drc_peer_free(int dev, uint64_t dptr) {
  USE(dev);
  CU(p_cuMemFree((CUdeviceptr)dptr));
  return 0;
}

int drc_ipc_get_handle(int dev, uint64_t dptr, void* handle64) {
  USE(dev);
  static_assert(sizeof(CUipcMemHandle) == DRC_IPC_HANDLE_BYTES, "CUipcMemHandle is 64 bytes");
  CU(p_cuIpcGetMemHandle((CUipcMemHandle*)handle64, (CUdeviceptr)dptr));
  return 0;
}

int drc_ipc_open_handle(int dev, const void* handle64, uint64_t* dptr) {
  USE(dev);
  CUipcMemHandle h;
  memcpy(&h, handle64, sizeof h);
  CUdeviceptr p = 0;
  CU(p_cuIpcOpenMemHandle(&p, h, CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS));
  *dptr = (uint64_t)p;
  return 0;
}

int drc_ipc_close_handle(int dev, uint64_t dptr) {
  USE(dev);
  CU(p_cuIpcCloseMemHandle((CUdeviceptr)dptr));
  return 0;
}

int drc_host_alloc(size_t bytes, void** hptr) {
  USE(g_cur_dev);
  CU(p_cuMemHostAlloc(hptr, bytes ? bytes : 1, CU_MEMHOSTALLOC_PORTABLE));
  return 0;
}

int drc_host_free(void* hptr) {
  USE(g_cur_dev);
  CU(p_cuMemFreeHost(hptr));
  return 0;
}

int drc_host_register(void* hptr, size_t bytes) {
  USE(g_cur_dev);
  CU(p_cuMemHostRegister(hptr, bytes, CU_MEMHOSTREGISTER_PORTABLE));
  return 0;
}

int drc_host_unregister(void* hptr) {
  USE(g_cur_dev);
  CU(p_cuMemHostUnregister(hptr));
  return 0;
}

// ---------------------------------------------------------------- compile + load
// NVRTC is bound with dlopen by ABSOLUTE path, newest toolkit first.  A plain -lnvrtc binds to
// whichever libnvrtc.so.12 the process loaded first -- PyTorch's wheel preloads its own 12.8
// copy, whose ptxas generates measurably worse SASS for the packed f32x2 kernels (24 extra
// negation FADDs per Black-Scholes vector, profiles/r1_bs_v8*) than the image's 12.9 toolkit.
#define NVRTC_FNS(X) X(nvrtcCreateProgram) X(nvrtcCompileProgram) X(nvrtcGetProgramLogSize) \
  X(nvrtcGetProgramLog) X(nvrtcDestroyProgram) X(nvrtcGetCUBINSize) X(nvrtcGetCUBIN)        \
  X(nvrtcGetErrorString) X(nvrtcVersion)
#define X(fn) static decltype(&fn) q_##fn = nullptr;
NVRTC_FNS(X)
#undef X
static std::once_flag g_nvrtc_once;
static std::string g_nvrtc_path, g_nvrtc_err;

static void load_nvrtc() {
  std::vector<std::string> cand;
  if (const char* e = getenv("DRC_NVRTC")) cand.push_back(e);
  if (const char* e = getenv("CUDA_HOME")) cand.push_back(std::string(e) + "/lib64/libnvrtc.so.12");
  cand.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
  cand.push_back("/usr/local/cuda/lib64/libnvrtc.so");
  cand.push_back("libnvrtc.so.12");
  cand.push_back("libnvrtc.so");
  void* h = nullptr;
  for (const auto& c : cand) {
    h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (h) { g_nvrtc_path = c; break; }
  }
  if (!h) { g_nvrtc_err = "libnvrtc not found (set DRC_NVRTC or CUDA_HOME)"; return; }
#define X(fn) q_##fn = (decltype(&fn))dlsym(h, #fn); if (!q_##fn) g_nvrtc_err = "missing NVRTC symbol " #fn;
  NVRTC_FNS(X)
#undef X
}
static int need_nvrtc() {
  std::call_once(g_nvrtc_once, load_nvrtc);
  if (!g_nvrtc_err.empty()) return fail("%s", g_nvrtc_err.c_str());
  return 0;
}
#define nvrtcCreateProgram q_nvrtcCreateProgram
#define nvrtcCompileProgram q_nvrtcCompileProgram
#define nvrtcGetProgramLogSize q_nvrtcGetProgramLogSize
#define nvrtcGetProgramLog q_nvrtcGetProgramLog
#define nvrtcDestroyProgram q_nvrtcDestroyProgram
#define nvrtcGetCUBINSize q_nvrtcGetCUBINSize
#define nvrtcGetCUBIN q_nvrtcGetCUBIN
#define nvrtcGetErrorString q_nvrtcGetErrorString

// version * 1000 + minor * 10 of the NVRTC in use and the path it was loaded from
int drc_nvrtc_version(int* major, int* minor, const char** path) {
  if (need_nvrtc()) return 1;
  int ma = 0, mi = 0;
  q_nvrtcVersion(&ma, &mi);
  if (major) *major = ma;
  if (minor) *minor = mi;
  if (path) *path = g_nvrtc_path.c_str();
  return 0;
}

int drc_compile(const char* source, const char* name, const char* const* options,
                int num_options, void** cubin, size_t* cubin_len, char** log) {
  *cubin = nullptr;
  *cubin_len = 0;
  if (log) *log = nullptr;
  if (need_nvrtc()) return 1;
  nvrtcProgram prog;
  nvrtcResult r = nvrtcCreateProgram(&prog, source, name, 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) return fail("nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
  r = nvrtcCompileProgram(prog, num_options, options);
  size_t log_len = 0;
  nvrtcGetProgramLogSize(prog, &log_len);
  std::string text(log_len ? log_len : 1, '\0');
  if (log_len > 1) nvrtcGetProgramLog(prog, &text[0]);
  if (log && log_len > 1) {
    *log = (char*)malloc(log_len + 1);
    memcpy(*log, text.data(), log_len);
    (*log)[log_len] = 0;
  }
  if (r != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    return fail("nvrtcCompileProgram(%s): %s\n%s", name, nvrtcGetErrorString(r), text.c_str());
  }
  size_t n = 0;
  r = nvrtcGetCUBINSize(prog, &n);
  if (r != NVRTC_SUCCESS || n == 0) {
    nvrtcDestroyProgram(&prog);
    return fail("nvrtcGetCUBINSize: %s (pass --gpu-architecture=sm_100a, not compute_)",
                nvrtcGetErrorString(r));
  }
  void* blob = malloc(n);
  r = nvrtcGetCUBIN(prog, (char*)blob);
  nvrtcDestroyProgram(&prog);
  if (r != NVRTC_SUCCESS) {
    free(blob);
    return fail("nvrtcGetCUBIN: %s", nvrtcGetErrorString(r));
  }
  *cubin = blob;
  *cubin_len = n;
  return 0;
}

int drc_free_blob(void* blob) {
  free(blob);
  return 0;
}

int drc_module_load(int dev, const void* cubin, size_t, uint64_t* module) {
  USE(dev);
  CUmodule m;
  CU(p_cuModuleLoadData(&m, cubin));
  *module = (uint64_t)m;
  return 0;
}

int drc_module_unload(int dev, uint64_t module) {
  USE(dev);
  CU(p_cuModuleUnload((CUmodule)module));
  return 0;
}

int drc_module_get_function(int dev, uint64_t module, const char* entry, uint64_t* func) {
  USE(dev);
  CUfunction f;
  CU(p_cuModuleGetFunction(&f, (CUmodule)module, entry));
  *func = (uint64_t)f;
  return 0;
}

int drc_func_set_max_dynamic_smem(int dev, uint64_t func, int bytes) {
  USE(dev);
  CU(p_cuFuncSetAttribute((CUfunction)func, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, bytes));
  return 0;
}

int drc_func_attrs(int dev, uint64_t func, int* num_regs, int* static_smem, int* local_bytes,
                   int* max_threads) {
  USE(dev);
  CUfunction f = (CUfunction)func;
  CU(p_cuFuncGetAttribute(num_regs, CU_FUNC_ATTRIBUTE_NUM_REGS, f));
  CU(p_cuFuncGetAttribute(static_smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, f));
  CU(p_cuFuncGetAttribute(local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f));
  CU(p_cuFuncGetAttribute(max_threads, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, f));
  return 0;
}

int drc_occupancy(int dev, uint64_t func, int block_threads, size_t dyn_smem, int* blocks_per_sm) {
  USE(dev);
  CU(p_cuOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, (CUfunction)func,
                                                   block_threads, dyn_smem));
  return 0;
}

// ---------------------------------------------------------------- launch
static int launch_impl(int dev, int stream, uint64_t func, uint32_t gx, uint32_t gy, uint32_t gz,
                       uint32_t bx, uint32_t by, uint32_t bz, uint32_t dyn_smem,
                       uint32_t cluster_x, void** args) {
  USE(dev);
  STREAM(dev, stream, s);
  if (cluster_x > 1) {
    CUlaunchConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDimX = gx; cfg.gridDimY = gy; cfg.gridDimZ = gz;
    cfg.blockDimX = bx; cfg.blockDimY = by; cfg.blockDimZ = bz;
    cfg.sharedMemBytes = dyn_smem;
    cfg.hStream = s;
    CUlaunchAttribute attr;
    memset(&attr, 0, sizeof attr);
    attr.id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
    attr.value.clusterDim.x = cluster_x;
    attr.value.clusterDim.y = 1;
    attr.value.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    CU(p_cuLaunchKernelEx(&cfg, (CUfunction)func, args, nullptr));
  } else {
    CU(p_cuLaunchKernel((CUfunction)func, gx, gy, gz, bx, by, bz, dyn_smem, s, args, nullptr));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int drc_launch(int dev, int stream, uint64_t func, const uint32_t grid[3],
               const uint32_t block[3], uint32_t dyn_smem, void** args, int) {
  return launch_impl(dev, stream, func, grid[0], grid[1], grid[2], block[0], block[1], block[2],
                     dyn_smem, 1, args);
}

int drc_launch_packed(int dev, int stream, uint64_t func, uint32_t gx, uint32_t gy, uint32_t gz,
                      uint32_t bx, uint32_t by, uint32_t bz, uint32_t dyn_smem,
                      uint32_t cluster_x, const void* blob, const uint32_t* offsets,
                      int num_args) {
  void* argv[64];
  if (num_args > 64) return fail("too many kernel arguments (%d > 64)", num_args);
  for (int i = 0; i < num_args; ++i) argv[i] = (char*)blob + offsets[i];
  return launch_impl(dev, stream, func, gx, gy, gz, bx, by, bz, dyn_smem, cluster_x, argv);
}

uint64_t drc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int drc_tensormap_encode(int dev, void* out128, int dtype_code, uint32_t rank, uint64_t gptr,
                         const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle, int l2_promotion) {
  USE(dev);
  CUtensorMapDataType dt;
  switch (dtype_code) {
    case 0: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    case 1: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64; break;
    case 2: dt = CU_TENSOR_MAP_DATA_TYPE_INT32; break;
    case 3: dt = CU_TENSOR_MAP_DATA_TYPE_INT64; break;
    case 4: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    default: return fail("bad dtype code %d", dtype_code);
  }
  if (rank < 1 || rank > 5) return fail("tensor map rank %u out of range", rank);
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];   // stride of dim i+1, in bytes
  }
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
  CU(p_cuTensorMapEncodeTiled((CUtensorMap*)out128, dt, rank, (void*)gptr, gdim, gstr, bdim, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle,
                              (CUtensorMapL2promotion)l2_promotion,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  return 0;
}

// ---------------------------------------------------------------- ordering / timing
int drc_stream_sync(int dev, int stream) {
  USE(dev);
  STREAM(dev, stream, s);
  CU(p_cuStreamSynchronize(s));
  return 0;
}

int drc_device_sync(int dev) {
  USE(dev);
  CU(p_cuCtxSynchronize());
  return 0;
}

int drc_event_create(int dev, uint64_t* event) {
  USE(dev);
  CUevent e;
  CU(p_cuEventCreate(&e, CU_EVENT_DEFAULT));
  *event = (uint64_t)e;
  return 0;
}

int drc_event_destroy(int dev, uint64_t event) {
  USE(dev);
  CU(p_cuEventDestroy((CUevent)event));
  return 0;
}

int drc_event_record(int dev, int stream, uint64_t event) {
  USE(dev);
  STREAM(dev, stream, s);
  CU(p_cuEventRecord((CUevent)event, s));
  return 0;
}

int drc_event_sync(int dev, uint64_t event) {
  USE(dev);
  CU(p_cuEventSynchronize((CUevent)event));
  return 0;
}

int drc_event_elapsed_ms(int dev, uint64_t start, uint64_t stop, float* ms) {
  USE(dev);
  CU(p_cuEventElapsedTime(ms, (CUevent)start, (CUevent)stop));
  return 0;
}

int drc_stream_wait_event(int dev, int stream, uint64_t event) {
  USE(dev);
  STREAM(dev, stream, s);
  CU(p_cuStreamWaitEvent(s, (CUevent)event, 0));
  return 0;
}

int drc_stream_handle(int dev, int stream, uint64_t* custream) {
  USE(dev);
  STREAM(dev, stream, s);
  *custream = (uint64_t)s;
  return 0;
}

// ---------------------------------------------------------------- NCCL
int drc_nccl_available(void) { return load_nccl() == 0 ? 1 : 0; }

int drc_nccl_get_unique_id(void* id128) {
  if (int e = load_nccl()) return e;
  static_assert(sizeof(ncclUniqueId) == DRC_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  NC(p_ncclGetUniqueId((ncclUniqueId*)id128));
  return 0;
}

int drc_nccl_init_rank(int dev, int nranks, int rank, const void* id128, uint64_t* comm) {
  if (int e = load_nccl()) return e;
  USE(dev);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t c;
  NC(p_ncclCommInitRank(&c, nranks, id, rank));
  *comm = (uint64_t)c;
  return 0;
}

int drc_nccl_init_all(int ndev, const int* devs, uint64_t* comms) {
  if (int e = load_nccl()) return e;
  std::vector<ncclComm_t> cs(ndev);
  for (int i = 0; i < ndev; ++i) USE(devs[i]);
  NC(p_ncclCommInitAll(cs.data(), ndev, devs));
  for (int i = 0; i < ndev; ++i) comms[i] = (uint64_t)cs[i];
  return 0;
}

int drc_nccl_destroy(uint64_t comm) {
  if (int e = load_nccl()) return e;
  NC(p_ncclCommDestroy((ncclComm_t)comm));
  return 0;
}

int drc_nccl_allreduce(uint64_t comm, int dev, int stream, uint64_t sendbuf, uint64_t recvbuf,
                       size_t count, int dtype_code, int op) {
  if (int e = load_nccl()) return e;
  USE(dev);
  STREAM(dev, stream, s);
  ncclDataType_t t;
  if (int e = nccl_dtype(dtype_code, &t)) return e;
  static const ncclRedOp_t ops[4] = {ncclSum, ncclProd, ncclMax, ncclMin};
  if (op < 0 || op > 3) return fail("bad reduction op %d", op);
  NC(p_ncclAllReduce((const void*)sendbuf, (void*)recvbuf, count, t, ops[op], (ncclComm_t)comm,
                     (cudaStream_t)s));
  return 0;
}

int drc_nccl_sendrecv(uint64_t comm, int dev, int stream, uint64_t sendbuf, size_t send_bytes,
                      int send_peer, uint64_t recvbuf, size_t recv_bytes, int recv_peer) {
  if (int e = load_nccl()) return e;
  USE(dev);
  STREAM(dev, stream, s);
  NC(p_ncclGroupStart());
  if (send_peer >= 0 && send_bytes)
    NC(p_ncclSend((const void*)sendbuf, send_bytes, ncclUint8, send_peer, (ncclComm_t)comm,
                  (cudaStream_t)s));
  if (recv_peer >= 0 && recv_bytes)
    NC(p_ncclRecv((void*)recvbuf, recv_bytes, ncclUint8, recv_peer, (ncclComm_t)comm,
                  (cudaStream_t)s));
  NC(p_ncclGroupEnd());
  return 0;
}

int drc_nccl_allgather(uint64_t comm, int dev, int stream, uint64_t sendbuf, uint64_t recvbuf,
                       size_t bytes_per_rank) {
  if (int e = load_nccl()) return e;
  USE(dev);
  STREAM(dev, stream, s);
  NC(p_ncclAllGather((const void*)sendbuf, (void*)recvbuf, bytes_per_rank, ncclUint8,
                     (ncclComm_t)comm, (cudaStream_t)s));
  return 0;
}

int drc_nccl_group_start(void) {
  if (int e = load_nccl()) return e;
  NC(p_ncclGroupStart());
  return 0;
}

int drc_nccl_group_end(void) {
  if (int e = load_nccl()) return e;
  NC(p_ncclGroupEnd());
  return 0;
}

// ---------------------------------------------------------------- cuFFT
namespace {
typedef int cufftHandle_t;
typedef int (*fn_cufftPlan1d)(cufftHandle_t*, int, int, int);
typedef int (*fn_cufftSetStream)(cufftHandle_t, void*);
typedef int (*fn_cufftExec)(cufftHandle_t, void*, void*, int);
typedef int (*fn_cufftDestroy)(cufftHandle_t);
void* g_libcufft = nullptr;
fn_cufftPlan1d p_cufftPlan1d = nullptr;
fn_cufftSetStream p_cufftSetStream = nullptr;
fn_cufftExec p_cufftExecC2C = nullptr, p_cufftExecZ2Z = nullptr;
struct FftPlan { int dev, n, batch, is_double; cufftHandle_t h; };
std::vector<FftPlan> g_plans;

int load_cufft() {
  if (g_libcufft) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_libcufft) return 0;
  void* h = dlopen("libcufft.so.11", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libcufft.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail("libcufft not found: %s", dlerror());
  p_cufftPlan1d = (fn_cufftPlan1d)dlsym(h, "cufftPlan1d");
  p_cufftSetStream = (fn_cufftSetStream)dlsym(h, "cufftSetStream");
  p_cufftExecC2C = (fn_cufftExec)dlsym(h, "cufftExecC2C");
  p_cufftExecZ2Z = (fn_cufftExec)dlsym(h, "cufftExecZ2Z");
  if (!p_cufftPlan1d || !p_cufftSetStream || !p_cufftExecC2C || !p_cufftExecZ2Z)
    return fail("libcufft lacks a required symbol");
  g_libcufft = h;
  return 0;
}
}  // namespace

int drc_fft_c2c_1d(int dev, int stream, uint64_t in, uint64_t out, int n, int batch,
                   int is_double, int inverse) {
  if (int e = load_cufft()) return e;
  USE(dev);
  STREAM(dev, stream, s);
  cufftHandle_t h = -1;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& p : g_plans)
      if (p.dev == dev && p.n == n && p.batch == batch && p.is_double == is_double) h = p.h;
    if (h < 0) {
      int r = p_cufftPlan1d(&h, n, is_double ? 0x69 : 0x29, batch);     // CUFFT_Z2Z : CUFFT_C2C
      if (r != 0) return fail("cufftPlan1d(n=%d, batch=%d) failed: %d", n, batch, r);
      g_plans.push_back({dev, n, batch, is_double, h});
    }
  }
  int r = p_cufftSetStream(h, (void*)s);
  if (r != 0) return fail("cufftSetStream failed: %d", r);
  r = (is_double ? p_cufftExecZ2Z : p_cufftExecC2C)(h, (void*)in, (void*)out, inverse ? 1 : -1);
  if (r != 0) return fail("cufftExec failed: %d", r);
  return 0;
}

}  // extern "C"
