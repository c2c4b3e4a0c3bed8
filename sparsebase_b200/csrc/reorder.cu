// reorder.cu -- degree features, DegreeReorder, Permute1D/2D, InversePermutation, row sharding.
//
//   sb200_degrees             feature/degrees.cc:93-105
//   sb200_degree_distribution feature/degree_distribution.cc:146-162
//   sb200_degree_reorder      reorder/degree_reorder.cc:22-62
//   sb200_permute2d           permute/permute_order_two.cc:21-79 (+ CSR ctor sort, csr.cc:99-157)
//   sb200_permute1d           permute/permute_order_one.cc:17-37
//   sb200_inverse_permutation bases/reorder_base.h:662-671
//   sb200_partition_rows      (new) nnz-balanced row blocks for the multi-GPU path
#include <limits>

#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "segsort.cuh"

namespace sb200 {

constexpr int kMapBlock = 256;
inline unsigned map_grid(int64_t n, int per_thread = 1) {
  int64_t b = ceil_div(n, (int64_t)kMapBlock * per_thread);
  return (unsigned)(b > 0 ? b : 1);
}

// ---------------------------------------------------------------- degree features
template <typename I, typename N>
__global__ void degrees_kernel(const N *__restrict__ row_ptr, int64_t n, I *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) st_stream(out + i, (I)(row_ptr[i + 1] - row_ptr[i]));
}

// dist[i] = (rows[i+1]-rows[i]) / (FeatureType)num_edges  -- the N-typed difference is
// converted to F and divided with IEEE round-to-nearest (no fast-math in this library).
template <typename N, typename F>
__global__ void degree_distribution_kernel(const N *__restrict__ row_ptr, int64_t n, N num_edges,
                                           F *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const N d = row_ptr[i + 1] - row_ptr[i];
    F q;
    if constexpr (std::is_same_v<F, float>)
      q = __fdiv_rn((float)d, (float)num_edges);
    else
      q = __ddiv_rn((double)d, (double)num_edges);
    st_stream(out + i, q);
  }
}

// ---------------------------------------------------------------- permutations
template <typename I>
__global__ void inverse_permutation_kernel(const I *__restrict__ perm, int64_t n,
                                           I *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[perm[i]] = (I)i;
}

// out[order[i]] = vals[i]: identical to the reference's two steps inv[order[i]] = i;
// out[j] = vals[inv[j]] for a permutation `order`, with one pass instead of two.
template <typename I, typename V>
__global__ void permute1d_kernel(const V *__restrict__ vals, const I *__restrict__ order,
                                 int64_t n, V *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[order[i]] = ld_stream(vals + i);
}

// ---------------------------------------------------------------- DegreeReorder
// Rank of vertex u in the order (degree ascending, id DESCENDING) -- the reference fills each
// degree bucket from its end (degree_reorder.cc:42-46).  Keys = degrees, laid out so that the
// stable sort sees ids in descending order: slot p holds vertex n-1-p.
template <typename I, typename N>
__global__ void degree_keys_kernel(const N *__restrict__ row_ptr, int64_t n,
                                   uint32_t *__restrict__ keys32, uint64_t *__restrict__ keys64,
                                   I *__restrict__ ids, unsigned long long *__restrict__ max_deg) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long d = 0;
  if (p < n) {
    const int64_t u = n - 1 - p;
    d = (unsigned long long)(row_ptr[u + 1] - row_ptr[u]);
    if (keys32) keys32[p] = (uint32_t)d;
    if (keys64) keys64[p] = (uint64_t)d;
    ids[p] = (I)u;
  }
  d = warp_reduce_max(d);
  if (lane_id() == 0 && d > 0) atomicMax(max_deg, d);
}

// inv[sorted[pos]] = ascending ? pos : n-1-pos   (degree_reorder.cc:47-57)
template <typename I>
__global__ void degree_rank_kernel(const I *__restrict__ sorted, int64_t n, int ascending,
                                   I *__restrict__ inv) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) inv[sorted[p]] = (I)(ascending ? p : n - 1 - p);
}

template <typename I, typename N>
void degree_reorder_impl(Workspace &ws, int64_t n, const N *row_ptr, bool ascending, I *out_inv) {
  if (n <= 0) return;
  cudaStream_t st = ws.stream();
  using UI = typename std::make_unsigned<I>::type;
  const bool wide = sizeof(N) == 8;  // degrees may exceed 32 bits only with 64-bit nnz
  uint32_t *k32 = wide ? nullptr : ws.alloc<uint32_t>(n);
  uint64_t *k64 = wide ? ws.alloc<uint64_t>(n) : nullptr;
  UI *ids = ws.alloc<UI>(n);
  unsigned long long *max_deg = ws.alloc<unsigned long long>(1);
  SB_CUDA(cudaMemsetAsync(max_deg, 0, sizeof(unsigned long long), st));
  SB_LAUNCH((degree_keys_kernel<I, N>), map_grid(n), kMapBlock, 0, st, row_ptr, n, k32, k64,
            (I *)ids, max_deg);
  unsigned long long h_max = 0;
  SB_CUDA(cudaMemcpyAsync(&h_max, max_deg, sizeof(h_max), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  std::vector<RsBitRange> ranges = {{0, bits_for(h_max)}};
  const int P = n > 1 ? rs_num_passes(ranges) : 0;
  UI *ids_out = ws.alloc<UI>(n);
  UI *ids_tmp = P > 1 ? ws.alloc<UI>(n) : nullptr;
  if (wide) {
    uint64_t *ko = ws.alloc<uint64_t>(n), *kt = P > 1 ? ws.alloc<uint64_t>(n) : nullptr;
    radix_sort<uint64_t, UI, NoVal>(ws, {k64, ids, nullptr}, {ko, ids_out, nullptr},
                                    {kt, ids_tmp, nullptr}, n, ranges);
  } else {
    uint32_t *ko = ws.alloc<uint32_t>(n), *kt = P > 1 ? ws.alloc<uint32_t>(n) : nullptr;
    radix_sort<uint32_t, UI, NoVal>(ws, {k32, ids, nullptr}, {ko, ids_out, nullptr},
                                    {kt, ids_tmp, nullptr}, n, ranges);
  }
  SB_LAUNCH((degree_rank_kernel<I>), map_grid(n), kMapBlock, 0, st, (const I *)ids_out, n,
            ascending ? 1 : 0, out_inv);
}

// ---------------------------------------------------------------- Permute2D
// Old row i becomes new row j = row_order[i].  One coalesced pass over xadj scatters, per new
// row, the source offset of its first entry and its length; the scan and the gather kernels
// then read both arrays coalesced (no dependent irow -> xadj gathers inside the hot kernels).
// The same pass finds the longest row, which selects the gather kernel.
template <typename I, typename N>
__global__ void permute_prepare_kernel(const N *__restrict__ xadj, const I *__restrict__ row_order,
                                       int64_t n, int64_t *__restrict__ src_base,
                                       N *__restrict__ new_len,
                                       unsigned long long *__restrict__ max_len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long len = 0;
  if (i < n) {
    const int64_t j = row_order ? (int64_t)row_order[i] : i;
    const N b = xadj[i], e = xadj[i + 1];
    src_base[j] = (int64_t)b;
    new_len[j] = e - b;
    len = (unsigned long long)(e - b);
  }
  // one atomic per CTA, and only when it would raise the maximum (same-address atomics
  // serialise in L2: one per warp cost 0.25 ms at 16.7 M rows)
  __shared__ unsigned long long s_max[kMapBlock / 32];
  len = warp_reduce_max(len);
  if (lane_id() == 0) s_max[threadIdx.x >> 5] = len;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long mx = 0;
    for (int w = 0; w < kMapBlock / 32; w++) mx = s_max[w] > mx ? s_max[w] : mx;
    if (mx > *reinterpret_cast<volatile unsigned long long *>(max_len)) atomicMax(max_len, mx);
  }
}

template <typename I, typename N, typename V>
struct GatherLoader {
  const int64_t *src_base;  // per new row: offset of the old row's first entry
  const I *adj;
  const V *vals;
  const I *col_order;  // old col -> new col, or null
  __device__ int64_t seg_base(int64_t r) const { return src_base[r]; }
  // default cache policy on purpose: a gathered row shares its 32-byte sectors with the rows
  // next to it in the SOURCE, which are gathered a little later (evict-first loads made HBM
  // deliver those sectors twice)
  __device__ I raw_key(int64_t p) const { return __ldg(adj + p); }
  __device__ I map_key(I c) const { return col_order ? col_order[c] : c; }
  __device__ V val(int64_t p) const { return __ldg(vals + p); }
};

// ---- matrices whose rows all have <= kShortRow entries (stencils, meshes).  One warp owns
//      32 consecutive NEW rows per step.  Their entries are fetched in concatenated order
//      (lane s reads the s-th entry of the 32-row batch: consecutive lanes read consecutive
//      addresses inside a source row), staged in a 2 KB per-warp slice of shared memory,
//      sorted by the lane that owns the row with a fixed compare-exchange network in
//      registers, and written back as ONE contiguous, fully coalesced run -- the 32 rows are
//      adjacent in the output.  No block barriers; three dependent loads per step (row records
//      -> entries -> renumbered columns). ----
constexpr int kShortRow = 8;
constexpr int kSrBlock = 256;

template <typename I, typename V>
__device__ __forceinline__ void short_cex(I &ka, V &va, I &kb, V &vb) {
  bool swap = kb < ka;
  if constexpr (has_val<V>) swap = swap || (kb == ka && vb < va);
  if (swap) {
    const I tk = ka;
    ka = kb;
    kb = tk;
    if constexpr (has_val<V>) {
      const V tv = va;
      va = vb;
      vb = tv;
    }
  }
}

template <typename I, typename N, typename V>
__global__ void __launch_bounds__(kSrBlock)
    permute_short_rows_kernel(const int64_t *__restrict__ src_base,
                              const N *__restrict__ out_ptr, const I *__restrict__ adj,
                              const V *__restrict__ vals, const I *__restrict__ col_order,
                              int64_t n, I *__restrict__ out_col, V *__restrict__ out_vals) {
  using VR = typename std::conditional<has_val<V>, V, char>::type;
  __shared__ I stage_k[kSrBlock / 32][32 * kShortRow];
  __shared__ VR stage_v[kSrBlock / 32][has_val<V> ? 32 * kShortRow : 1];
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  I *sk = stage_k[wid];
  [[maybe_unused]] VR *sv = stage_v[wid];
  const int64_t nbatches = (n + 31) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * kSrBlock) >> 5;
  for (int64_t bt = (((int64_t)blockIdx.x * kSrBlock) >> 5) + wid; bt < nbatches; bt += wstride) {
    const int64_t j = (bt << 5) + lane;
    int64_t ob = 0, p = 0;
    unsigned len = 0;
    if (j < n) {
      ob = (int64_t)out_ptr[j];
      len = (unsigned)((int64_t)out_ptr[j + 1] - ob);
      p = src_base[j];
    }
    const int64_t ob0 = __shfl_sync(0xffffffffu, ob, 0);
    const unsigned incl = warp_inclusive_scan(len);
    const unsigned excl = incl - len;  // == ob - ob0 for the rows that exist
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    // ---- fetch in concatenated order ----
    for (unsigned base = 0; base < total; base += 32) {
      const unsigned sidx = base + lane;
      unsigned owner = 0;  // number of lanes whose inclusive end <= sidx
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const unsigned val = __shfl_sync(0xffffffffu, incl, (owner + step - 1) & 31);
        if (val <= sidx) owner += step;
      }
      owner &= 31;
      const int64_t p_o = __shfl_sync(0xffffffffu, p, owner);
      const unsigned ex_o = __shfl_sync(0xffffffffu, excl, owner);
      if (sidx < total) {
        const int64_t src = p_o + (int64_t)(sidx - ex_o);
        I c = __ldg(adj + src);
        if constexpr (has_val<V>) sv[sidx] = __ldg(vals + src);
        if (col_order) c = col_order[c];
        sk[sidx] = c;
      }
    }
    __syncwarp();
    // ---- the owning lane sorts its row in registers ----
    I k[kShortRow];
    [[maybe_unused]] VR v[kShortRow];
#pragma unroll
    for (int u = 0; u < kShortRow; u++) {
      k[u] = std::numeric_limits<I>::max();  // padding sorts to the end
      if constexpr (has_val<V>) v[u] = V(0);
      if ((unsigned)u < len) {
        k[u] = sk[excl + u];
        if constexpr (has_val<V>) v[u] = sv[excl + u];
      }
    }
    // Batcher odd-even merge sort for 8 keys (19 compare-exchanges)
#define SB_CEX(a, b)                                          \
  if constexpr (has_val<V>)                                   \
    short_cex<I, V>(k[a], v[a], k[b], v[b]);                  \
  else {                                                      \
    NoVal nv1, nv2;                                           \
    short_cex<I, NoVal>(k[a], nv1, k[b], nv2);                \
  }
    SB_CEX(0, 1) SB_CEX(2, 3) SB_CEX(4, 5) SB_CEX(6, 7)
    SB_CEX(0, 2) SB_CEX(1, 3) SB_CEX(4, 6) SB_CEX(5, 7)
    SB_CEX(1, 2) SB_CEX(5, 6)
    SB_CEX(0, 4) SB_CEX(1, 5) SB_CEX(2, 6) SB_CEX(3, 7)
    SB_CEX(2, 4) SB_CEX(3, 5)
    SB_CEX(1, 2) SB_CEX(3, 4) SB_CEX(5, 6)
#undef SB_CEX
#pragma unroll
    for (int u = 0; u < kShortRow; u++) {
      if ((unsigned)u < len) {
        sk[excl + u] = k[u];
        if constexpr (has_val<V>) sv[excl + u] = v[u];
      }
    }
    __syncwarp();
    // ---- one contiguous run of the output ----
    for (unsigned sidx = lane; sidx < total; sidx += 32) {
      st_stream(out_col + ob0 + sidx, sk[sidx]);
      if constexpr (has_val<V>) st_stream(out_vals + ob0 + sidx, (V)sv[sidx]);
    }
    __syncwarp();
  }
}

template <typename I, typename N, typename V>
void permute2d_impl(Workspace &ws, int64_t n, int64_t m, int64_t nnz, const N *xadj,
                    const I *adj, const V *vals, const I *row_order, const I *col_order,
                    N *out_row_ptr, I *out_col, V *out_vals) {
  cudaStream_t st = ws.stream();
  int64_t *src_base = ws.alloc<int64_t>(n + 1);
  N *new_len = ws.alloc<N>(n + 1);
  unsigned long long *max_len = ws.alloc<unsigned long long>(1);
  SB_CUDA(cudaMemsetAsync(max_len, 0, sizeof(unsigned long long), st));
  if (n > 0)
    SB_LAUNCH((permute_prepare_kernel<I, N>), map_grid(n), kMapBlock, 0, st, xadj, row_order, n,
              src_base, new_len, max_len);
  exclusive_scan<N>(ws, LoadFn<N>{new_len}, out_row_ptr, n);
  if (n <= 0 || nnz <= 0) return;
  unsigned long long h_max = 0;
  SB_CUDA(cudaMemcpyAsync(&h_max, max_len, sizeof(h_max), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (h_max <= (unsigned long long)kShortRow) {
    // one 32-row batch per warp, CTAs in row order: neighbouring rows are in flight at the
    // same time, so the sectors they share (20-byte rows in 32-byte sectors, nearby col_order
    // entries) are fetched from HBM once.  (A grid-stride loop over a capped grid spreads the
    // resident warps over distant row ranges and doubled the DRAM read traffic.)
    SB_LAUNCH((permute_short_rows_kernel<I, N, V>), (unsigned)ceil_div(n, (int64_t)kSrBlock), kSrBlock,
              0, st, (const int64_t *)src_base, (const N *)out_row_ptr, adj, vals, col_order, n,
              out_col, out_vals);
    return;
  }
  GatherLoader<I, N, V> ld{src_base, adj, vals, col_order};
  segmented_sort<I, N, V>(ws, ld, (const N *)out_row_ptr, n, m, nnz, out_col, out_vals);
}

// ---------------------------------------------------------------- row sharding
template <typename N>
__global__ void partition_rows_kernel(const N *__restrict__ row_ptr, int64_t n, int64_t nnz,
                                      int parts, int64_t *__restrict__ bounds) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > parts) return;
  if (k == 0) {
    bounds[0] = 0;
    return;
  }
  if (k == parts) {
    bounds[k] = n;
    return;
  }
  const int64_t target = (int64_t)(((__int128)nnz * k) / parts);
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)row_ptr[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  bounds[k] = lo;
}

}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_degrees(int device, int64_t n, const void *row_ptr, void *out_degrees, int id_type,
                  int nnz_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && (n == 0 || (row_ptr && out_degrees)), SB200_ERR_BAD_ARG, "bad argument");
    if (n == 0) return;
    dispatch_inv(id_type, nnz_type, SB200_VOID, false, [&](auto I_, auto N_, auto) {
      using I = decltype(I_);
      using N = decltype(N_);
      SB_LAUNCH((degrees_kernel<I, N>), map_grid(n), kMapBlock, 0, (cudaStream_t)stream,
                (const N *)row_ptr, n, (I *)out_degrees);
    });
  });
}

int sb200_degree_distribution(int device, int64_t n, int64_t nnz, const void *row_ptr,
                              void *out_dist, int nnz_type, int feature_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && (n == 0 || (row_ptr && out_dist)), SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(feature_type == SB200_F32 || feature_type == SB200_F64, SB200_ERR_BAD_DTYPE,
               "feature_type must be F32 or F64");
    SB_REQUIRE(is_int_dtype(nnz_type), SB200_ERR_BAD_DTYPE, "bad nnz_type");
    if (n == 0) return;
    cudaStream_t st = (cudaStream_t)stream;
    auto run = [&](auto N_, auto F_) {
      using N = decltype(N_);
      using F = decltype(F_);
      SB_LAUNCH((degree_distribution_kernel<N, F>), map_grid(n), kMapBlock, 0, st,
                (const N *)row_ptr, n, (N)nnz, (F *)out_dist);
    };
    if (dtype_size(nnz_type) == 4) {
      if (feature_type == SB200_F32)
        run(int32_t{}, float{});
      else
        run(int32_t{}, double{});
    } else {
      if (feature_type == SB200_F32)
        run(int64_t{}, float{});
      else
        run(int64_t{}, double{});
    }
  });
}

int sb200_degree_reorder(int device, int64_t n, const void *row_ptr, int ascending,
                         void *out_inv, int id_type, int nnz_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && (n == 0 || (row_ptr && out_inv)), SB200_ERR_BAD_ARG, "bad argument");
    Workspace ws(device, (cudaStream_t)stream);
    dispatch_inv(id_type, nnz_type, SB200_VOID, false, [&](auto I_, auto N_, auto) {
      using I = decltype(I_);
      using N = decltype(N_);
      degree_reorder_impl<I, N>(ws, n, (const N *)row_ptr, ascending != 0, (I *)out_inv);
    });
  });
}

int sb200_permute2d(int device, int64_t n, int64_t m, int64_t nnz, const void *row_ptr,
                    const void *col, const void *vals, const void *row_order,
                    const void *col_order, void *out_row_ptr, void *out_col, void *out_vals,
                    int id_type, int nnz_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && m >= 0 && nnz >= 0 && row_ptr && out_row_ptr, SB200_ERR_BAD_ARG,
               "bad argument");
    SB_REQUIRE(nnz == 0 || (col && out_col), SB200_ERR_BAD_ARG, "null array");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      permute2d_impl<I, N, V>(ws, n, m, nnz, (const N *)row_ptr, (const I *)col,
                              (const V *)vals, (const I *)row_order, (const I *)col_order,
                              (N *)out_row_ptr, (I *)out_col, (V *)out_vals);
    });
  });
}

int sb200_permute1d(int device, int64_t len, const void *vals, const void *order, void *out,
                    int id_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(len >= 0 && (len == 0 || (vals && order && out)), SB200_ERR_BAD_ARG,
               "bad argument");
    if (len == 0) return;
    const int vb = dtype_size(val_type);
    SB_REQUIRE(vb == 4 || vb == 8, SB200_ERR_BAD_DTYPE, "val_type %d unsupported", val_type);
    cudaStream_t st = (cudaStream_t)stream;
    dispatch_id(id_type, [&](auto I_) {
      using I = decltype(I_);
      if (vb == 4)
        SB_LAUNCH((permute1d_kernel<I, uint32_t>), map_grid(len), kMapBlock, 0, st,
                  (const uint32_t *)vals, (const I *)order, len, (uint32_t *)out);
      else
        SB_LAUNCH((permute1d_kernel<I, uint64_t>), map_grid(len), kMapBlock, 0, st,
                  (const uint64_t *)vals, (const I *)order, len, (uint64_t *)out);
    });
  });
}

int sb200_inverse_permutation(int device, int64_t len, const void *perm, void *out, int id_type,
                              void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(len >= 0 && (len == 0 || (perm && out)), SB200_ERR_BAD_ARG, "bad argument");
    if (len == 0) return;
    dispatch_id(id_type, [&](auto I_) {
      using I = decltype(I_);
      SB_LAUNCH((inverse_permutation_kernel<I>), map_grid(len), kMapBlock, 0,
                (cudaStream_t)stream, (const I *)perm, len, (I *)out);
    });
  });
}

int sb200_partition_rows(int device, int64_t n, int64_t nnz, const void *row_ptr, int nnz_type,
                         int parts, int64_t *h_bounds, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && parts >= 1 && parts <= 1024 && row_ptr && h_bounds, SB200_ERR_BAD_ARG,
               "bad argument");
    SB_REQUIRE(is_int_dtype(nnz_type), SB200_ERR_BAD_DTYPE, "bad nnz_type");
    Workspace ws(device, (cudaStream_t)stream);
    cudaStream_t st = ws.stream();
    int64_t *bounds = ws.alloc<int64_t>(parts + 1);
    if (dtype_size(nnz_type) == 4)
      SB_LAUNCH((partition_rows_kernel<int32_t>), (unsigned)ceil_div(parts + 1, 128), 128, 0, st,
                (const int32_t *)row_ptr, n, nnz, parts, bounds);
    else
      SB_LAUNCH((partition_rows_kernel<int64_t>), (unsigned)ceil_div(parts + 1, 128), 128, 0, st,
                (const int64_t *)row_ptr, n, nnz, parts, bounds);
    SB_CUDA(cudaMemcpyAsync(h_bounds, bounds, (parts + 1) * sizeof(int64_t),
                            cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
