// runtime.cu -- contexts, memory and transfer entry points of the C ABI, plus the shared
// error / scratch-memory plumbing.  Replaces the reference's cudaMalloc+cudaMemcpy conversion
// functions (converter/converter_order_two_cuda.cu:11-105, converter_order_one_cuda.cu:10-43),
// CUDAPeerToPeer (converter/converter_cuda.cu:12-21), CUDAContext validation
// (context/cuda_context_cuda.cu:9-15) and CUDADeleter (utils/utils_cuda.cuh:6-9); unlike the
// reference every CUDA return code is checked.
#include <mutex>

#include "common.cuh"

namespace sb200 {

static thread_local char g_err[1024] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int64_t &launch_counter() { return g_launches; }

const DeviceInfo &device_info(int device) {
  static std::mutex mu;
  static DeviceInfo infos[64];
  static bool have[64] = {false};
  std::lock_guard<std::mutex> lk(mu);
  SB_REQUIRE(device >= 0 && device < 64, SB200_ERR_BAD_DEVICE, "device %d out of range", device);
  if (!have[device]) {
    cudaDeviceProp p;
    SB_CUDA(cudaGetDeviceProperties(&p, device));
    infos[device].sm_count = p.multiProcessorCount;
    infos[device].max_smem_optin = (int)p.sharedMemPerBlockOptin;
    infos[device].l2_bytes = (int64_t)p.l2CacheSize;
    have[device] = true;
  }
  return infos[device];
}

static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {nullptr};

cudaMemPool_t scratch_pool(int device) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  SB_REQUIRE(device >= 0 && device < 64, SB200_ERR_BAD_DEVICE, "device %d out of range", device);
  if (!g_pools[device]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t pool;
    SB_CUDA(cudaMemPoolCreate(&pool, &props));
    // keep freed scratch blocks in OUR pool instead of returning them to the driver on every
    // stream synchronisation; sb200_trim gives them back
    uint64_t thresh = UINT64_MAX;
    SB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    g_pools[device] = pool;
  }
  return g_pools[device];
}

Workspace::Workspace(int device, cudaStream_t stream) : device_(device), stream_(stream) {
  device_info(device);
  pool_ = scratch_pool(device);
}
Workspace::~Workspace() {
  for (void *p : ptrs_) cudaFreeAsync(p, stream_);
}
void *Workspace::alloc_bytes(size_t bytes) {
  void *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMallocFromPoolAsync(&p, bytes, pool_, stream_);
  if (e != cudaSuccess) {
    set_error("scratch allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    throw Error{SB200_ERR_ALLOC};
  }
  ptrs_.push_back(p);
  return p;
}

}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_abi_version(void) { return SB200_ABI_VERSION; }
const char *sb200_last_error(void) { return g_err; }
int64_t sb200_launch_count(void) { return g_launches; }
void sb200_reset_launch_count(void) { g_launches = 0; }

int sb200_device_count(int *h_count) {
  if (!h_count) return SB200_ERR_BAD_ARG;
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess) {
    *h_count = 0;
    set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    return SB200_ERR_CUDA;
  }
  *h_count = cnt;
  return SB200_OK;
}

int sb200_can_access_peer(int device, int peer_device, int *h_can) {
  return guarded(device, [&] {
    SB_REQUIRE(h_can, SB200_ERR_BAD_ARG, "h_can is null");
    if (device == peer_device) {
      *h_can = 1;
      return;
    }
    SB_CUDA(cudaDeviceCanAccessPeer(h_can, device, peer_device));
  });
}

int sb200_enable_peer_access(int device, int peer_device) {
  return guarded(device, [&] {
    if (device == peer_device) return;
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
      return;
    }
    SB_CUDA(e);
  });
}

int sb200_malloc(int device, size_t bytes, void **h_out_ptr) {
  return guarded(device, [&] {
    SB_REQUIRE(h_out_ptr, SB200_ERR_BAD_ARG, "h_out_ptr is null");
    *h_out_ptr = nullptr;
    cudaError_t e = cudaMalloc(h_out_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      throw Error{SB200_ERR_ALLOC};
    }
  });
}

int sb200_free(int device, void *ptr) {
  return guarded(device, [&] {
    if (ptr) SB_CUDA(cudaFree(ptr));
  });
}

int sb200_malloc_host(size_t bytes, void **h_out_ptr) {
  if (!h_out_ptr) return SB200_ERR_BAD_ARG;
  cudaError_t e = cudaMallocHost(h_out_ptr, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return SB200_ERR_ALLOC;
  }
  return SB200_OK;
}

int sb200_free_host(void *h_ptr) {
  if (!h_ptr) return SB200_OK;
  cudaError_t e = cudaFreeHost(h_ptr);
  if (e != cudaSuccess) {
    set_error("cudaFreeHost failed: %s", cudaGetErrorString(e));
    return SB200_ERR_CUDA;
  }
  return SB200_OK;
}

int sb200_memcpy_h2d(int device, void *dst, const void *h_src, size_t bytes, void *stream) {
  return guarded(device, [&] {
    if (bytes == 0) return;
    SB_REQUIRE(dst && h_src, SB200_ERR_BAD_ARG, "null pointer in memcpy_h2d");
    SB_CUDA(cudaMemcpyAsync(dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  });
}

int sb200_memcpy_d2h(int device, void *h_dst, const void *src, size_t bytes, void *stream) {
  return guarded(device, [&] {
    if (bytes == 0) return;
    SB_REQUIRE(h_dst && src, SB200_ERR_BAD_ARG, "null pointer in memcpy_d2h");
    SB_CUDA(cudaMemcpyAsync(h_dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  });
}

int sb200_memcpy_d2d(int dst_device, void *dst, int src_device, const void *src, size_t bytes,
                     void *stream) {
  return guarded(dst_device, [&] {
    if (bytes == 0) return;
    SB_REQUIRE(dst && src, SB200_ERR_BAD_ARG, "null pointer in memcpy_d2d");
    if (dst_device == src_device)
      SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    else
      SB_CUDA(cudaMemcpyPeerAsync(dst, dst_device, src, src_device, bytes, (cudaStream_t)stream));
  });
}

int sb200_trim(int device) {
  return guarded(device, [&] {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (device >= 0 && device < 64 && g_pools[device]) {
      SB_CUDA(cudaDeviceSynchronize());
      SB_CUDA(cudaMemPoolTrimTo(g_pools[device], 0));
    }
  });
}

int sb200_stream_synchronize(int device, void *stream) {
  return guarded(device, [&] { SB_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); });
}

}  // extern "C"
