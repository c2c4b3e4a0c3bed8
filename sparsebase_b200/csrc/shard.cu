// shard.cu -- device-side helpers of the multi-GPU (row-block sharded) path.  None of these
// has a counterpart in the reference (it has no distributed code); they
// are the per-GPU pieces that sparsebase_b200/sharded.py stitches together with
// torch.distributed collectives.
//
//   sb200_exclusive_scan        out[0..n] = exclusive prefix sums (row_ptr from lengths)
//   sb200_rank_keys             rank of every key among n DISTINCT keys (groups a rank's rows
//                               by their destination in the sharded Permute2D)
//   sb200_max_degree            max_i (row_ptr[i+1] - row_ptr[i])
//   sb200_degree_histogram      hist[d] = #rows of the block with degree d
//   sb200_degree_rank_combine   out[i] = local_rank[i] + offset[degree(i)]: turns the block-local
//                               DegreeReorder rank into the global one
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

template <typename I>
__global__ void iota_kernel(I *p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (I)i;
}

template <typename I>
__global__ void scatter_rank_kernel(const I *__restrict__ sorted_idx, int64_t n,
                                    I *__restrict__ rank) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) rank[sorted_idx[p]] = (I)p;
}

// Persistent CTAs with a private histogram of the small degrees in shared memory: a power-law
// block holds tens of millions of rows of degree 1..8, and one global atomic per warp and
// degree still serialised on a handful of addresses (1.75 ms on the low-degree block of
// R-MAT-26 at 4 GPUs).  Large degrees are rare and go straight to global memory.
constexpr int kHistSmall = 2048;
template <typename N>
__global__ void __launch_bounds__(256)
    degree_histogram_kernel(const N *__restrict__ row_ptr, int64_t n, int64_t nbins,
                            unsigned long long *__restrict__ hist) {
  __shared__ unsigned small[kHistSmall];
  for (int k = threadIdx.x; k < kHistSmall; k += blockDim.x) small[k] = 0;
  __syncthreads();
  const int64_t span = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += span) {
    const int64_t i = base + threadIdx.x;
    int64_t d = -1;
    if (i < n) d = (int64_t)(row_ptr[i + 1] - row_ptr[i]);
    const bool ok = d >= 0 && d < nbins;
    // neighbouring rows often share a degree: one atomic per distinct degree in the warp
    const unsigned peers = __match_any_sync(0xffffffffu, ok ? d : -1);
    if (ok && (int)lane_id() == __ffs(peers) - 1) {
      if (d < kHistSmall)
        atomicAdd(&small[d], (unsigned)__popc(peers));
      else
        atomicAdd(&hist[d], (unsigned long long)__popc(peers));
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kHistSmall && k < nbins; k += blockDim.x)
    if (small[k]) atomicAdd(&hist[k], (unsigned long long)small[k]);
}

template <typename I, typename N>
__global__ void degree_rank_combine_kernel(const N *__restrict__ row_ptr,
                                           const I *__restrict__ local_rank,
                                           const int64_t *__restrict__ offset, int64_t n,
                                           int64_t flip_from, I *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t d = (int64_t)(row_ptr[i + 1] - row_ptr[i]);
  const int64_t g = (int64_t)local_rank[i] + offset[d];
  out[i] = (I)(flip_from >= 0 ? flip_from - g : g);
}

template <typename N>
__global__ void max_degree2_kernel(const N *__restrict__ xadj, int64_t n,
                                   unsigned long long *__restrict__ out) {
  unsigned long long m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long d = (unsigned long long)(xadj[i + 1] - xadj[i]);
    m = d > m ? d : m;
  }
  m = warp_reduce_max(m);
  if (lane_id() == 0 && m) atomicMax(out, m);
}

}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_exclusive_scan(int device, int64_t n, const void *in, void *out, int dtype,
                         void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && out && (n == 0 || in), SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(is_int_dtype(dtype), SB200_ERR_BAD_DTYPE, "dtype %d is not an integer dtype", dtype);
    Workspace ws(device, (cudaStream_t)stream);
    if (dtype_size(dtype) == 4)
      exclusive_scan<int32_t>(ws, LoadFn<int32_t>{(const int32_t *)in}, (int32_t *)out, n);
    else
      exclusive_scan<int64_t>(ws, LoadFn<int64_t>{(const int64_t *)in}, (int64_t *)out, n);
  });
}

int sb200_rank_keys(int device, int64_t n, const void *keys, int64_t key_bound, void *out_rank,
                    int id_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && key_bound >= 0 && (n == 0 || (keys && out_rank)), SB200_ERR_BAD_ARG,
               "bad argument");
    if (n == 0) return;
    Workspace ws(device, (cudaStream_t)stream);
    cudaStream_t st = ws.stream();
    dispatch_id(id_type, [&](auto I_) {
      using I = decltype(I_);
      using UI = typename std::make_unsigned<I>::type;
      UI *idx = ws.alloc<UI>(n), *idx_out = ws.alloc<UI>(n), *idx_tmp = ws.alloc<UI>(n);
      UI *k_out = ws.alloc<UI>(n), *k_tmp = ws.alloc<UI>(n);
      SB_LAUNCH((iota_kernel<UI>), (unsigned)ceil_div(n, 256), 256, 0, st, idx, n);
      std::vector<RsBitRange> ranges = {{0, bits_for((uint64_t)(key_bound > 0 ? key_bound - 1 : 0))}};
      radix_sort<UI, UI, NoVal>(ws, {(UI *)keys, idx, nullptr}, {k_out, idx_out, nullptr},
                                {k_tmp, idx_tmp, nullptr}, n, ranges);
      SB_LAUNCH((scatter_rank_kernel<I>), (unsigned)ceil_div(n, 256), 256, 0, st,
                (const I *)idx_out, n, (I *)out_rank);
    });
  });
}

int sb200_max_degree(int device, int64_t n, const void *row_ptr, int nnz_type, int64_t *h_out,
                     void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && h_out && (n == 0 || row_ptr), SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(is_int_dtype(nnz_type), SB200_ERR_BAD_DTYPE, "bad nnz_type");
    *h_out = 0;
    if (n == 0) return;
    Workspace ws(device, (cudaStream_t)stream);
    cudaStream_t st = ws.stream();
    unsigned long long *md = ws.alloc<unsigned long long>(1);
    SB_CUDA(cudaMemsetAsync(md, 0, sizeof(*md), st));
    const int grid = device_info(device).sm_count * 8;
    if (dtype_size(nnz_type) == 4)
      SB_LAUNCH((max_degree2_kernel<int32_t>), grid, 256, 0, st, (const int32_t *)row_ptr, n, md);
    else
      SB_LAUNCH((max_degree2_kernel<int64_t>), grid, 256, 0, st, (const int64_t *)row_ptr, n, md);
    unsigned long long h = 0;
    SB_CUDA(cudaMemcpyAsync(&h, md, sizeof(h), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    *h_out = (int64_t)h;
  });
}

int sb200_degree_histogram(int device, int64_t n, const void *row_ptr, int nnz_type,
                           int64_t nbins, void *out_hist, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && nbins >= 0 && out_hist && (n == 0 || row_ptr), SB200_ERR_BAD_ARG,
               "bad argument");
    SB_REQUIRE(is_int_dtype(nnz_type), SB200_ERR_BAD_DTYPE, "bad nnz_type");
    cudaStream_t st = (cudaStream_t)stream;
    SB_CUDA(cudaMemsetAsync(out_hist, 0, nbins * sizeof(unsigned long long), st));
    if (n == 0) return;
    const int64_t want = ceil_div(n, 256), cap = (int64_t)device_info(device).sm_count * 8;
    const unsigned hist_grid = (unsigned)(want < cap ? want : cap);
    if (dtype_size(nnz_type) == 4)
      SB_LAUNCH((degree_histogram_kernel<int32_t>), hist_grid, 256, 0, st,
                (const int32_t *)row_ptr, n, nbins, (unsigned long long *)out_hist);
    else
      SB_LAUNCH((degree_histogram_kernel<int64_t>), hist_grid, 256, 0, st,
                (const int64_t *)row_ptr, n, nbins, (unsigned long long *)out_hist);
  });
}

int sb200_degree_rank_combine(int device, int64_t n, const void *row_ptr, const void *local_rank,
                              const void *offset, int64_t flip_from, void *out, int id_type,
                              int nnz_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && (n == 0 || (row_ptr && local_rank && offset && out)), SB200_ERR_BAD_ARG,
               "bad argument");
    if (n == 0) return;
    dispatch_inv(id_type, nnz_type, SB200_VOID, false, [&](auto I_, auto N_, auto) {
      using I = decltype(I_);
      using N = decltype(N_);
      SB_LAUNCH((degree_rank_combine_kernel<I, N>), (unsigned)ceil_div(n, 256), 256, 0,
                (cudaStream_t)stream, (const N *)row_ptr, (const I *)local_rank,
                (const int64_t *)offset, n, flip_from, (I *)out);
    });
  });
}

}  // extern "C"
