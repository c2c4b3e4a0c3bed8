// common.cuh -- shared plumbing of libsb200: error handling, dtype dispatch, stream-ordered
// scratch memory, launch accounting and the warp/block primitives the kernels are built from.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../include/sb200.h"

namespace sb200 {

constexpr int kWarp = 32;

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int64_t &launch_counter();

struct Error {
  int code;
};

#define SB_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      sb200::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,               \
                       cudaGetErrorString(e__));                                        \
      throw sb200::Error{e__ == cudaErrorMemoryAllocation ? SB200_ERR_ALLOC             \
                                                          : SB200_ERR_CUDA};            \
    }                                                                                   \
  } while (0)

#define SB_REQUIRE(cond, code, ...)       \
  do {                                    \
    if (!(cond)) {                        \
      sb200::set_error(__VA_ARGS__);      \
      throw sb200::Error{code};           \
    }                                     \
  } while (0)

// Every kernel launch goes through this so that launches are counted and checked.
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
  do {                                                                         \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                \
    sb200::launch_counter()++;                                                 \
    SB_CUDA(cudaGetLastError());                                               \
  } while (0)

// Selects `device` for the lifetime of the object and puts the caller's current device back
// afterwards (also when the body throws): an entry point never changes the calling thread's
// device, so a call on cuda:1 made while cuda:0 is current leaves cuda:0 current.
class DeviceScope {
 public:
  explicit DeviceScope(int device) {
    if (cudaGetDevice(&prev_) != cudaSuccess) prev_ = -1;
    if (prev_ != device) {
      cudaError_t e = cudaSetDevice(device);
      if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
        throw Error{SB200_ERR_CUDA};
      }
      switched_ = true;
    }
  }
  ~DeviceScope() {
    if (switched_ && prev_ >= 0) cudaSetDevice(prev_);
  }
  DeviceScope(const DeviceScope &) = delete;
  DeviceScope &operator=(const DeviceScope &) = delete;

 private:
  int prev_ = -1;
  bool switched_ = false;
};

// Wraps the body of an extern "C" entry point: selects the device (restoring the caller's on
// the way out), converts exceptions to codes.
template <typename Fn>
int guarded(int device, Fn &&fn) {
  try {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess) {
      set_error("cudaGetDeviceCount failed: %s (libsb200 has no CPU fallback)",
                cudaGetErrorString(e));
      return SB200_ERR_CUDA;
    }
    if (device < 0 || device >= cnt) {
      set_error("device %d out of range (device count %d)", device, cnt);
      return SB200_ERR_BAD_DEVICE;
    }
    DeviceScope scope(device);
    fn();
    return SB200_OK;
  } catch (const Error &err) {
    return err.code;
  } catch (const std::bad_alloc &) {
    set_error("host allocation failed");
    return SB200_ERR_ALLOC;
  } catch (...) {
    set_error("unexpected exception");
    return SB200_ERR_INTERNAL;
  }
}

// ---------------------------------------------------------------- device properties
struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
  int64_t l2_bytes;
};
const DeviceInfo &device_info(int device);

// ---------------------------------------------------------------- scratch memory
// Stream-ordered scratch from a PRIVATE memory pool per device (cudaMallocFromPoolAsync; the
// release threshold of that pool is raised so that repeated calls re-use the same blocks
// without touching the driver).  The device's default pool -- shared with every other library
// in the process -- is left alone; sb200_trim(device) hands the cached blocks back.
cudaMemPool_t scratch_pool(int device);
class Workspace {
 public:
  Workspace(int device, cudaStream_t stream);
  ~Workspace();
  template <typename T>
  T *alloc(size_t count) {
    return static_cast<T *>(alloc_bytes(count * sizeof(T)));
  }
  void *alloc_bytes(size_t bytes);
  cudaStream_t stream() const { return stream_; }
  int device() const { return device_; }

 private:
  int device_;
  cudaStream_t stream_;
  cudaMemPool_t pool_;
  std::vector<void *> ptrs_;
};

// ---------------------------------------------------------------- dtype dispatch
inline int dtype_size(int dt) {
  switch (dt) {
    case SB200_I32:
    case SB200_U32:
    case SB200_F32:
      return 4;
    case SB200_I64:
    case SB200_U64:
    case SB200_F64:
      return 8;
    default:
      return 0;
  }
}
inline bool is_int_dtype(int dt) {
  return dt == SB200_I32 || dt == SB200_U32 || dt == SB200_I64 || dt == SB200_U64;
}

struct NoVal {};  // ValueType = void / vals == nullptr
template <typename V>
constexpr bool has_val = !std::is_same_v<V, NoVal>;

// Calls f(I{}, N{}, V{}) with I in {int32,int64}, N in {int32,int64} (sizeof N >= sizeof I),
// V in {NoVal, uint32, uint64}.  Index values are < 2^31 / 2^62, so unsigned reference
// types are handled by the signed kernels of the same width.
template <typename Fn>
void dispatch_inv(int id_type, int nnz_type, int val_type, bool has_vals, Fn &&f) {
  SB_REQUIRE(is_int_dtype(id_type), SB200_ERR_BAD_DTYPE, "id_type %d is not an integer dtype",
             id_type);
  SB_REQUIRE(is_int_dtype(nnz_type), SB200_ERR_BAD_DTYPE, "nnz_type %d is not an integer dtype",
             nnz_type);
  int ib = dtype_size(id_type), nb = dtype_size(nnz_type);
  int vb = has_vals ? dtype_size(val_type) : 0;
  SB_REQUIRE(!has_vals || vb != 0, SB200_ERR_BAD_DTYPE, "val_type %d unsupported", val_type);
  SB_REQUIRE(nb >= ib, SB200_ERR_BAD_DTYPE,
             "nnz_type narrower than id_type is not supported (id %d B, nnz %d B)", ib, nb);
  auto with_v = [&](auto I_, auto N_) {
    if (vb == 0)
      f(I_, N_, NoVal{});
    else if (vb == 4)
      f(I_, N_, uint32_t{});
    else
      f(I_, N_, uint64_t{});
  };
  if (ib == 4 && nb == 4)
    with_v(int32_t{}, int32_t{});
  else if (ib == 4 && nb == 8)
    with_v(int32_t{}, int64_t{});
  else
    with_v(int64_t{}, int64_t{});
}

template <typename Fn>
void dispatch_id(int id_type, Fn &&f) {
  SB_REQUIRE(is_int_dtype(id_type), SB200_ERR_BAD_DTYPE, "id_type %d is not an integer dtype",
             id_type);
  if (dtype_size(id_type) == 4)
    f(int32_t{});
  else
    f(int64_t{});
}

// ---------------------------------------------------------------- device helpers
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, v, o);
    if ((int)lane_id() >= o) v += t;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_reduce_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_reduce_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_reduce_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// Block-wide exclusive scan of one value per thread.  `warp_sums` is shared scratch of at
// least 33 T.  Returns the exclusive prefix; *total (optional) receives the block sum.
// Contains __syncthreads(): must be called by all threads of the block.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *warp_sums, T *total = nullptr) {
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  T inc = warp_inclusive_scan(v);
  __syncthreads();  // protect warp_sums from a previous use
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    T s = lane < nw ? warp_sums[lane] : T(0);
    T si = warp_inclusive_scan(s);
    warp_sums[lane] = si - s;
    if (lane == 31) warp_sums[32] = si;
  }
  __syncthreads();
  if (total) *total = warp_sums[32];
  return warp_sums[wid] + inc - v;
}

// Values travel through the kernels as bit patterns (uint32_t / uint64_t).  Where the reference
// COMPARES values -- std::less<std::pair<IDType, ValueType>> on duplicate ids, csr.cc:147 -- the
// pattern is mapped to an unsigned key that orders like the real type `vkind` (an SB200_* code):
// floats: negative values reversed below the positive ones; signed integers: sign bit flipped.
template <typename V>
__device__ __forceinline__ V val_order_key(V bits, int vkind) {
  constexpr V sign = V(1) << (sizeof(V) * 8 - 1);
  if (vkind == SB200_F32 || vkind == SB200_F64) return (bits & sign) ? V(~bits) : V(bits | sign);
  if (vkind == SB200_I32 || vkind == SB200_I64) return V(bits ^ sign);
  return bits;
}

// Duplicate ids inside a segment (outside the reference's own input contract, but its result is
// still defined): the reference sorts EVERY segment by (id, value) as soon as ANY segment is
// unsorted, and otherwise leaves all of them untouched (csr.cc:99-157).  The fast kernels order
// ties arbitrarily, flag the segments that hold duplicates and report whether any segment was
// unsorted in source order; dup_fix_kernel (segsort.cuh) then puts the flagged segments right.
struct DupCtx {
  unsigned char *seg_flag;  // [n_seg], zeroed: 1 = the segment holds duplicate ids
  unsigned *any_dup;        // zeroed: some segment is flagged
  unsigned *unsorted;       // zeroed: some segment had an inversion in source order
                            // (null: the caller knows the sort happens)
  __device__ __forceinline__ void flag(int64_t seg) const {
    if (seg_flag) {
      seg_flag[seg] = 1;
      *reinterpret_cast<volatile unsigned *>(any_dup) = 1u;
    }
  }
  // call with a block-uniform predicate per thread; contains __syncthreads_or
  __device__ __forceinline__ void report_unsorted(bool mine) const {
    const int any = __syncthreads_or(mine ? 1 : 0);
    if (unsorted && any && threadIdx.x == 0 &&
        *reinterpret_cast<volatile unsigned *>(unsorted) == 0u)
      *reinterpret_cast<volatile unsigned *>(unsorted) = 1u;
  }
};

// streaming (read-once / write-once) accesses: keep them out of L1
template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
  return __ldcs(p);
}
template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v) {
  __stcs(p, v);
}

}  // namespace sb200
