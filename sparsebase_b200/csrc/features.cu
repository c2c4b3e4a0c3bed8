// features.cu -- the callers on either side of the path.
//
//   sb200_edges_to_coo      io/edge_list_reader.cc:28-151 (EdgeListReader::ReadCOO on an
//                           in-memory edge list: self-edge removal, undirected expansion,
//                           (row, col) sort, unique) -- the step right before the COO constructor
//   sb200_degree_features   feature/degrees_degree_distribution.cc:147-166 (degrees +
//                           distribution in one pass), feature/min_max_avg_degree.cc:168-191,
//                           feature/bandwidth.cc:92-111, feature/profile.cc:92-106 -- the quality
//                           metrics of a reordering, one pass over row_ptr and one over col
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

// ------------------------------------------------------------------------------------
// edge list -> sorted, de-duplicated COO
// ------------------------------------------------------------------------------------
// n = max u + 1, m = max v + 1 over the KEPT edges (edge_list_reader.cc:46-47)
template <typename I>
__global__ void __launch_bounds__(256)
    edges_extent_kernel(const I *__restrict__ u, const I *__restrict__ v, int64_t n_edges,
                        int remove_self, unsigned long long *__restrict__ out2) {
  unsigned long long mu = 0, mv = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges;
       i += (int64_t)gridDim.x * blockDim.x) {
    const I a = ld_stream(u + i), b = ld_stream(v + i);
    if (a != b || !remove_self) {
      mu = (unsigned long long)a + 1 > mu ? (unsigned long long)a + 1 : mu;
      mv = (unsigned long long)b + 1 > mv ? (unsigned long long)b + 1 : mv;
    }
  }
  mu = warp_reduce_max(mu);
  mv = warp_reduce_max(mv);
  if (lane_id() == 0) {
    if (mu) atomicMax(out2, mu);
    if (mv) atomicMax(out2 + 1, mv);
  }
}

// key = row << col_bits | col; the reverse edge right behind every kept edge when undirected
// (edge_list_reader.cc:43-44); dropped self edges carry the bit above the key bits, so they
// sort behind every real key and never split a run of equal keys
template <typename I, typename V>
__global__ void __launch_bounds__(256)
    edges_pack_kernel(const I *__restrict__ u, const I *__restrict__ v, const V *__restrict__ w,
                      int64_t n_edges, int remove_self, int undirected, int col_bits,
                      uint64_t dropped, uint64_t *__restrict__ keys, V *__restrict__ vals) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges;
       i += (int64_t)gridDim.x * blockDim.x) {
    const I a = ld_stream(u + i), b = ld_stream(v + i);
    const bool keep = a != b || !remove_self;
    const uint64_t k1 = keep ? ((uint64_t)a << col_bits) | (uint64_t)b : dropped;
    if (undirected) {
      const uint64_t k2 = keep ? ((uint64_t)b << col_bits) | (uint64_t)a : dropped;
      // the two entries of an edge are adjacent: one 16-byte store
      reinterpret_cast<ulonglong2 *>(keys)[i] = make_ulonglong2(k1, k2);
      if constexpr (has_val<V>) {
        if (vals) {
          const V x = ld_stream(w + i);
          vals[2 * i] = x;
          vals[2 * i + 1] = x;
        }
      }
    } else {
      keys[i] = k1;
      if constexpr (has_val<V>) {
        if (vals) vals[i] = ld_stream(w + i);
      }
    }
  }
}

// 1 at the first entry of every run of equal real keys (every real entry when duplicates stay)
struct EdgeHeadFn {
  const uint64_t *keys;
  uint64_t dropped;
  int remove_duplicates;
  __device__ int64_t operator()(int64_t i) const {
    const uint64_t k = keys[i];
    if (k >= dropped) return 0;
    return (!remove_duplicates || i == 0 || keys[i - 1] != k) ? 1 : 0;
  }
};

template <typename I, typename V>
__global__ void __launch_bounds__(256)
    edges_emit_kernel(const uint64_t *__restrict__ keys, const V *__restrict__ vals,
                      const int64_t *__restrict__ pos, int64_t cnt, int col_bits,
                      I *__restrict__ out_row, I *__restrict__ out_col, V *__restrict__ out_vals) {
  const uint64_t mask = (1ull << col_bits) - 1ull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = pos[i];
    if (pos[i + 1] != p) {  // a head: goes to slot p
      const uint64_t k = keys[i];
      out_row[p] = (I)(k >> col_bits);
      out_col[p] = (I)(k & mask);
      if constexpr (has_val<V>) {
        if (out_vals) out_vals[p] = vals[i];
      }
    }
  }
}

template <typename I, typename V>
void edges_to_coo_impl(Workspace &ws, int64_t n_edges, const I *u, const I *v, const V *w,
                       bool remove_duplicates, bool remove_self, bool undirected, bool square,
                       I *out_row, I *out_col, V *out_vals, int64_t *h_out3) {
  cudaStream_t st = ws.stream();
  h_out3[0] = h_out3[1] = h_out3[2] = 0;
  if (n_edges <= 0) return;
  const int grid = device_info(ws.device()).sm_count * 8;
  unsigned long long *ext = ws.alloc<unsigned long long>(2);
  SB_CUDA(cudaMemsetAsync(ext, 0, 2 * sizeof(unsigned long long), st));
  SB_LAUNCH((edges_extent_kernel<I>), grid, 256, 0, st, u, v, n_edges, remove_self ? 1 : 0, ext);
  unsigned long long h_ext[2] = {0, 0};
  SB_CUDA(cudaMemcpyAsync(h_ext, ext, sizeof(h_ext), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  int64_t n = (int64_t)h_ext[0], m = (int64_t)h_ext[1];
  if (square || undirected) {  // edge_list_reader.cc:51-54
    n = n > m ? n : m;
    m = n;
  }
  h_out3[0] = n;
  h_out3[1] = m;
  // the reverse edges swap the roles of u and v: both fields must hold either id
  const int64_t bound = undirected ? (n > m ? n : m) : 0;
  const int col_bits = bits_for((uint64_t)((undirected ? bound : m) > 0 ? (undirected ? bound : m) - 1 : 0));
  const int row_bits = bits_for((uint64_t)((undirected ? bound : n) > 0 ? (undirected ? bound : n) - 1 : 0));
  SB_REQUIRE(col_bits + row_bits <= 62, SB200_ERR_BAD_ARG,
             "(row,col) key needs %d bits; at most 62 supported", col_bits + row_bits);
  const int key_bits = col_bits + row_bits > 0 ? col_bits + row_bits : 1;
  const uint64_t dropped = 1ull << key_bits;
  const int64_t cnt = n_edges * (undirected ? 2 : 1);
  uint64_t *k0 = ws.alloc<uint64_t>(cnt + 1), *k1 = ws.alloc<uint64_t>(cnt + 1),
           *k2 = ws.alloc<uint64_t>(cnt + 1);
  V *v0 = nullptr, *v1 = nullptr, *v2 = nullptr;
  const bool hv = has_val<V> && w != nullptr && out_vals != nullptr;
  if constexpr (has_val<V>) {
    if (hv) {
      v0 = ws.alloc<V>(cnt);
      v1 = ws.alloc<V>(cnt);
      v2 = ws.alloc<V>(cnt);
    }
  }
  SB_LAUNCH((edges_pack_kernel<I, V>), grid, 256, 0, st, u, v, hv ? w : (const V *)nullptr,
            n_edges, remove_self ? 1 : 0, undirected ? 1 : 0, col_bits, dropped, k0, v0);
  // stable: among equal (row, col) the first in input order stays first (the survivor of unique)
  std::vector<RsBitRange> ranges = {{0, key_bits + 1}};
  const V *sorted_v = nullptr;
  if constexpr (has_val<V>) {
    if (hv) {
      radix_sort<uint64_t, V, NoVal>(ws, {k0, v0, nullptr}, {k1, v1, nullptr}, {k2, v2, nullptr},
                                     cnt, ranges);
      sorted_v = v1;
    } else {
      radix_sort<uint64_t, NoVal, NoVal>(ws, {k0, nullptr, nullptr}, {k1, nullptr, nullptr},
                                         {k2, nullptr, nullptr}, cnt, ranges);
    }
  } else {
    radix_sort<uint64_t, NoVal, NoVal>(ws, {k0, nullptr, nullptr}, {k1, nullptr, nullptr},
                                       {k2, nullptr, nullptr}, cnt, ranges);
  }
  int64_t *pos = ws.alloc<int64_t>(cnt + 1);
  exclusive_scan<int64_t>(ws, EdgeHeadFn{k1, dropped, remove_duplicates ? 1 : 0}, pos, cnt);
  SB_LAUNCH((edges_emit_kernel<I, V>), grid, 256, 0, st, (const uint64_t *)k1, sorted_v,
            (const int64_t *)pos, cnt, col_bits, out_row, out_col, hv ? out_vals : (V *)nullptr);
  int64_t nnz = 0;
  SB_CUDA(cudaMemcpyAsync(&nnz, pos + cnt, sizeof(nnz), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  h_out3[2] = nnz;
}

// ------------------------------------------------------------------------------------
// fused degree features + bandwidth / profile
// ------------------------------------------------------------------------------------
// out4 (device): {min degree, max degree, bandwidth, profile}
constexpr int kFtPer = 4;

template <typename I, typename N, typename F>
__global__ void __launch_bounds__(256)
    degree_features_rows_kernel(const N *__restrict__ row_ptr, int64_t n, N num_edges,
                                I *__restrict__ out_deg, F *__restrict__ out_dist,
                                unsigned long long *__restrict__ out4) {
  unsigned long long mn = ~0ull, mx = 0;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kFtPer;
  if (i0 < n) {
    N p[kFtPer + 1];
#pragma unroll
    for (int k = 0; k <= kFtPer; k++) p[k] = i0 + k <= n ? row_ptr[i0 + k] : N(0);
#pragma unroll
    for (int k = 0; k < kFtPer; k++) {
      if (i0 + k < n) {
        const N d = p[k + 1] - p[k];
        if (out_deg) st_stream(out_deg + i0 + k, (I)d);
        if (out_dist) {
          F q;
          if constexpr (std::is_same_v<F, float>)
            q = __fdiv_rn((float)d, (float)num_edges);
          else
            q = __ddiv_rn((double)d, (double)num_edges);
          st_stream(out_dist + i0 + k, q);
        }
        const unsigned long long ud = (unsigned long long)d;
        mn = ud < mn ? ud : mn;
        mx = ud > mx ? ud : mx;
      }
    }
  }
  mn = warp_reduce_min(mn);
  mx = warp_reduce_max(mx);
  __shared__ unsigned long long s_mn[8], s_mx[8];
  const unsigned wid = threadIdx.x >> 5;
  if (lane_id() == 0) {
    s_mn[wid] = mn;
    s_mx[wid] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) {
      mn = s_mn[w] < mn ? s_mn[w] : mn;
      mx = s_mx[w] > mx ? s_mx[w] : mx;
    }
    if (mn != ~0ull) atomicMin(out4, mn);
    atomicMax(out4 + 1, mx);
  }
}

// One warp per 32 consecutive rows, walking their concatenated entries 32 at a time (the rows
// are adjacent in memory, so the reads are coalesced whatever the row lengths).
//   bandwidth = max |i - j| + 1 (bandwidth.cc:101-107), profile = sum_i (i - min(i, min_k col))
template <typename I, typename N>
__global__ void __launch_bounds__(256)
    bandwidth_profile_kernel(const N *__restrict__ row_ptr, const I *__restrict__ col, int64_t n,
                             unsigned long long *__restrict__ out4) {
  __shared__ long long s_min[8][32];
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long bw = 0, prof = 0;
  for (int64_t g = warp * 32; g < n; g += nwarps * 32) {
    const int64_t r = g + lane;
    int64_t xs = 0;
    unsigned long long d = 0;
    if (r < n) {
      xs = (int64_t)row_ptr[r];
      d = (unsigned long long)((int64_t)row_ptr[r + 1] - xs);
    }
    s_min[wid][lane] = r;
    __syncwarp();
    const unsigned long long incl = warp_inclusive_scan(d);
    const unsigned long long excl = incl - d;
    const unsigned long long tot = __shfl_sync(0xffffffffu, incl, 31);
    for (unsigned long long base = 0; base < tot; base += 32) {
      const unsigned long long s = base + lane;
      unsigned lo = 0;  // number of lanes whose inclusive end <= s == owner lane
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const unsigned long long val = __shfl_sync(0xffffffffu, incl, (lo + step - 1) & 31);
        if (val <= s) lo += step;
      }
      const unsigned j = lo & 31;
      const int64_t xs_j = __shfl_sync(0xffffffffu, xs, j);
      const unsigned long long excl_j = __shfl_sync(0xffffffffu, excl, j);
      if (s < tot) {
        const long long c = (long long)ld_stream(col + xs_j + (int64_t)(s - excl_j));
        const long long i = g + j;
        const unsigned long long w = (unsigned long long)(i >= c ? i - c + 1 : c - i + 1);
        bw = w > bw ? w : bw;
        if (c < i) atomicMin(&s_min[wid][j], c);
      }
    }
    __syncwarp();
    if (r < n) prof += (unsigned long long)(r - s_min[wid][lane]);
    __syncwarp();
  }
  bw = warp_reduce_max(bw);
  prof = warp_reduce_sum(prof);
  if (lane == 0) {
    if (bw) atomicMax(out4 + 2, bw);
    if (prof) atomicAdd(out4 + 3, prof);
  }
}

// ------------------------------------------------------------------ BOBAReorder
// reorder/boba_reorder.cc:35-137: the COO is sorted by (col, row); a vertex is placed by its
// first appearance in the row array of that list, then -- if it never appears there -- by its
// first appearance in the column array, then by id (vertices without entries).  The parallel
// variant of the reference states it as a key: key[v] = min index of v in rows ++ cols
// (:107-118), never-seen vertices keep 2 * nnz, and vertices are ranked by (key, id) (:120-127);
// the sequential variant (:73-105) produces the same order.
template <typename SI, typename UI>
__global__ void boba_first_kernel(const UI *__restrict__ rows_sorted,
                                  const UI *__restrict__ cols_sorted, int64_t nnz,
                                  SI *__restrict__ key) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
       p += (int64_t)gridDim.x * blockDim.x) {
    const UI r = rows_sorted[p], c = cols_sorted[p];
    // (consecutive entries of one column are consecutive here: one atomic per run and lane
    // would do, but a row id repeats only across columns)
    atomicMin(&key[r], (SI)p);
    if (p == 0 || cols_sorted[p - 1] != c) atomicMin(&key[c], (SI)(nnz + p));
  }
}
template <typename SI>
__global__ void boba_fill_kernel(SI *__restrict__ key, int64_t nodes, SI value) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nodes) key[i] = value;
}

// ------------------------------------------------------------------ ReorderHeatmap
// reorder/reorder_heatmap.cc:43-120: density[bu][bv] counts the nonzeros whose permuted
// coordinates (order_r[i], order_c[col]) fall into block (bu, bv) of a b x b grid with
// bsize = n / b rows AND columns per block (the last block takes the remainder).
// Work is dealt out by ENTRIES (4096 per warp), as in the row push of the sharded Permute2D: a
// power-law matrix keeps its heavy rows next to each other.  Lane l resolves row l of the current
// batch of 32 rows (its block row, its clipped extent), then the warp walks the batch's
// concatenated entries 32 at a time.  Counts go to a CTA-private histogram in shared memory when
// the grid has at most kHeatShared cells, else straight to global memory.
constexpr int kHeatShared = 4096;
constexpr int64_t kHeatChunk = 4096;
template <typename I, typename N>
__global__ void __launch_bounds__(256)
    heatmap_count_kernel(const N *__restrict__ row_ptr, const I *__restrict__ col,
                         const I *__restrict__ order_r, const I *__restrict__ order_c, int64_t n,
                         int64_t nnz, unsigned long long bsize, int b,
                         unsigned long long *__restrict__ density) {
  __shared__ unsigned sh[kHeatShared];
  const int cells = b * b;
  const bool priv = cells <= kHeatShared;
  if (priv) {
    for (int k = threadIdx.x; k < cells; k += blockDim.x) sh[k] = 0;
    __syncthreads();
  }
  const unsigned lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nchunks = (nnz + kHeatChunk - 1) / kHeatChunk;
  for (int64_t w = warp; w < nchunks; w += nwarps) {
    const int64_t e0 = w * kHeatChunk;
    const int64_t e1 = e0 + kHeatChunk < nnz ? e0 + kHeatChunk : nnz;
    int64_t lo = 0, hi = n;  // last row with row_ptr[row] <= e0
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)row_ptr[mid] <= e0)
        lo = mid;
      else
        hi = mid;
    }
    for (int64_t g = lo; g < n; g += 32) {
      const int64_t i = g + lane;
      int64_t bgn = 0;
      unsigned d = 0, bu = 0;
      int64_t next_begin = nnz;
      if (i < n) {
        const int64_t rb = (int64_t)row_ptr[i], re = (int64_t)row_ptr[i + 1];
        next_begin = re;
        bgn = rb > e0 ? rb : e0;
        const int64_t e = re < e1 ? re : e1;
        if (e > bgn) {
          d = (unsigned)(e - bgn);
          const unsigned long long u = (unsigned long long)(order_r ? order_r[i] : (I)i);
          const unsigned long long q = u / bsize;
          bu = (unsigned)(q >= (unsigned long long)b ? b - 1 : q);
        }
      }
      const unsigned incl = warp_inclusive_scan(d);
      const unsigned excl = incl - d;
      const unsigned tot = __shfl_sync(0xffffffffu, incl, 31);
      for (unsigned base = 0; base < tot; base += 32) {
        const unsigned s = base + lane;
        unsigned own = 0;  // number of lanes whose inclusive end <= s == owner lane
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const unsigned val = __shfl_sync(0xffffffffu, incl, (own + step - 1) & 31);
          if (val <= s) own += step;
        }
        const unsigned j = own & 31;
        const int64_t bgn_j = __shfl_sync(0xffffffffu, bgn, j);
        const unsigned excl_j = __shfl_sync(0xffffffffu, excl, j);
        const unsigned bu_j = __shfl_sync(0xffffffffu, bu, j);
        if (s < tot) {
          const I c = ld_stream(col + bgn_j + (s - excl_j));
          const unsigned long long v = (unsigned long long)(order_c ? __ldg(order_c + c) : c);
          const unsigned long long q = v / bsize;
          const unsigned bv = (unsigned)(q >= (unsigned long long)b ? b - 1 : q);
          const unsigned cell = bu_j * (unsigned)b + bv;
          if (priv)
            atomicAdd(&sh[cell], 1u);
          else
            atomicAdd(&density[cell], 1ull);
        }
      }
      if (__shfl_sync(0xffffffffu, next_begin, 31) >= e1) break;
    }
  }
  if (priv) {
    __syncthreads();
    for (int k = threadIdx.x; k < cells; k += blockDim.x)
      if (sh[k]) atomicAdd(&density[k], (unsigned long long)sh[k]);
  }
}
// heat = density / (row_ptr[n] + .0f): a FLOAT division whatever FloatType is (:112)
template <typename N, typename F>
__global__ void heatmap_finish_kernel(const unsigned long long *__restrict__ density,
                                      const N *__restrict__ row_ptr, int64_t n, int cells,
                                      F *__restrict__ heat) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cells) return;
  const float total = (float)row_ptr[n];
  heat[k] = (F)__fdiv_rn((float)(N)density[k], total);
}

}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_edges_to_coo(int device, int64_t n_edges, const void *u, const void *v, const void *w,
                       int remove_duplicates, int remove_self_edges, int read_undirected,
                       int square, void *out_row, void *out_col, void *out_vals,
                       int64_t *h_out3, int id_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n_edges >= 0 && h_out3, SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(n_edges == 0 || (u && v && out_row && out_col), SB200_ERR_BAD_ARG, "null array");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = w != nullptr && out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, SB200_I64, val_type, hv, [&](auto I_, auto, auto V_) {
      using I = decltype(I_);
      using V = decltype(V_);
      edges_to_coo_impl<I, V>(ws, n_edges, (const I *)u, (const I *)v, (const V *)w,
                              remove_duplicates != 0, remove_self_edges != 0,
                              read_undirected != 0, square != 0, (I *)out_row, (I *)out_col,
                              (V *)out_vals, h_out3);
    });
  });
}

int sb200_degree_features(int device, int64_t n, int64_t nnz, const void *row_ptr,
                          const void *col, void *out_degrees, void *out_dist,
                          int64_t *h_out_scalars, double *h_out_avg, int id_type, int nnz_type,
                          int feature_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && nnz >= 0 && h_out_scalars && (n == 0 || row_ptr), SB200_ERR_BAD_ARG,
               "bad argument");
    SB_REQUIRE(feature_type == SB200_F32 || feature_type == SB200_F64, SB200_ERR_BAD_DTYPE,
               "feature_type must be F32 or F64");
    h_out_scalars[0] = h_out_scalars[1] = h_out_scalars[2] = h_out_scalars[3] = 0;
    if (h_out_avg) *h_out_avg = 0.0;
    if (n == 0) return;
    Workspace ws(device, (cudaStream_t)stream);
    cudaStream_t st = ws.stream();
    unsigned long long *out4 = ws.alloc<unsigned long long>(4);
    const unsigned long long init[4] = {~0ull, 0ull, 0ull, 0ull};
    SB_CUDA(cudaMemcpyAsync(out4, init, sizeof(init), cudaMemcpyHostToDevice, st));
    int64_t h_first = 0, h_last = 0;
    dispatch_inv(id_type, nnz_type, SB200_VOID, false, [&](auto I_, auto N_, auto) {
      using I = decltype(I_);
      using N = decltype(N_);
      const unsigned grid = (unsigned)ceil_div(n, (int64_t)256 * kFtPer);
      if (feature_type == SB200_F32)
        SB_LAUNCH((degree_features_rows_kernel<I, N, float>), grid, 256, 0, st,
                  (const N *)row_ptr, n, (N)nnz, (I *)out_degrees, (float *)out_dist, out4);
      else
        SB_LAUNCH((degree_features_rows_kernel<I, N, double>), grid, 256, 0, st,
                  (const N *)row_ptr, n, (N)nnz, (I *)out_degrees, (double *)out_dist, out4);
      if (col && nnz > 0) {
        const int64_t groups = ceil_div(n, 32);
        const int64_t cap = (int64_t)device_info(device).sm_count * 16;
        const int64_t blocks = ceil_div(groups, 8);
        SB_LAUNCH((bandwidth_profile_kernel<I, N>), (unsigned)(blocks < cap ? blocks : cap), 256,
                  0, st, (const N *)row_ptr, (const I *)col, n, out4);
      }
      N ends[2] = {0, 0};
      SB_CUDA(cudaMemcpyAsync(&ends[0], (const N *)row_ptr, sizeof(N), cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaMemcpyAsync(&ends[1], (const N *)row_ptr + n, sizeof(N), cudaMemcpyDeviceToHost,
                              st));
      unsigned long long h4[4];
      SB_CUDA(cudaMemcpyAsync(h4, out4, sizeof(h4), cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaStreamSynchronize(st));
      h_first = (int64_t)ends[0];
      h_last = (int64_t)ends[1];
      for (int k = 0; k < 4; k++) h_out_scalars[k] = (int64_t)h4[k];
    });
    // avg_degree.cc:134-135: degree_sum / (FeatureType) num_vertices, in FeatureType
    if (h_out_avg) {
      if (feature_type == SB200_F32)
        *h_out_avg = (double)((float)(h_last - h_first) / (float)n);
      else
        *h_out_avg = (double)(h_last - h_first) / (double)n;
    }
  });
}

int sb200_reorder_heatmap(int device, int64_t n, int64_t m, int64_t nnz, const void *row_ptr,
                          const void *col, const void *order_r, const void *order_c,
                          int num_parts, void *out_heat, int id_type, int nnz_type,
                          int feature_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && m >= 0 && nnz >= 0 && out_heat && (n == 0 || row_ptr) &&
                   (nnz == 0 || col),
               SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(feature_type == SB200_F32 || feature_type == SB200_F64, SB200_ERR_BAD_DTYPE,
               "feature_type must be F32 or F64");
    // reorder_heatmap.cc:52-56
    SB_REQUIRE(num_parts >= 1 && num_parts <= n && num_parts <= m, SB200_ERR_BAD_ARG,
               "Cannot generate heatmap for matrix when num_parts > number of rows or columns");
    SB_REQUIRE(num_parts <= 46340, SB200_ERR_BAD_ARG, "num_parts * num_parts must fit in an int");
    Workspace ws(device, (cudaStream_t)stream);
    cudaStream_t st = ws.stream();
    const int cells = num_parts * num_parts;
    unsigned long long *density = ws.alloc<unsigned long long>(cells);
    SB_CUDA(cudaMemsetAsync(density, 0, (size_t)cells * sizeof(unsigned long long), st));
    dispatch_inv(id_type, nnz_type, SB200_VOID, false, [&](auto I_, auto N_, auto) {
      using I = decltype(I_);
      using N = decltype(N_);
      if (nnz > 0) {
        const int64_t warps = ceil_div(nnz, kHeatChunk);
        const int64_t cap = (int64_t)device_info(device).sm_count * 8;
        const int64_t blocks = ceil_div(warps, 8);
        SB_LAUNCH((heatmap_count_kernel<I, N>), (unsigned)(blocks < cap ? blocks : cap), 256, 0, st,
                  (const N *)row_ptr, (const I *)col, (const I *)order_r, (const I *)order_c, n,
                  nnz, (unsigned long long)(n / num_parts), num_parts, density);
      }
      if (feature_type == SB200_F32)
        SB_LAUNCH((heatmap_finish_kernel<N, float>), (unsigned)ceil_div(cells, 256), 256, 0, st,
                  (const unsigned long long *)density, (const N *)row_ptr, n, cells,
                  (float *)out_heat);
      else
        SB_LAUNCH((heatmap_finish_kernel<N, double>), (unsigned)ceil_div(cells, 256), 256, 0, st,
                  (const unsigned long long *)density, (const N *)row_ptr, n, cells,
                  (double *)out_heat);
    });
  });
}

int sb200_boba_reorder(int device, int64_t n, int64_t m, int64_t nnz, const void *row,
                       const void *col, void *out_inv, int id_type, void *stream) {
  return guarded(device, [&] {
    const int64_t nodes = n > m ? n : m;
    SB_REQUIRE(n >= 0 && m >= 0 && nnz >= 0 && (nodes == 0 || out_inv) && (nnz == 0 || (row && col)),
               SB200_ERR_BAD_ARG, "bad argument");
    if (nodes == 0) return;
    Workspace ws(device, (cudaStream_t)stream);
    cudaStream_t st = ws.stream();
    dispatch_id(id_type, [&](auto I_) {
      using I = decltype(I_);
      using UI = typename std::make_unsigned<I>::type;
      using SI = typename std::conditional<sizeof(I) == 4, int, long long>::type;
      // boba_reorder.cc:44,113 computes nnzs * 2 in IDType
      SB_REQUIRE((uint64_t)nnz * 2 + 1 < ((uint64_t)1 << (8 * sizeof(I) - 1)), SB200_ERR_BAD_ARG,
                 "2 * nnz does not fit in IDType");
      SI *key = reinterpret_cast<SI *>(ws.alloc<I>(nodes));
      SB_LAUNCH((boba_fill_kernel<SI>), (unsigned)ceil_div(nodes, 256), 256, 0, st, key, nodes,
                (SI)(2 * nnz));
      if (nnz > 0) {
        UI *k_out = ws.alloc<UI>(nnz), *r_out = ws.alloc<UI>(nnz);
        UI *k_tmp = ws.alloc<UI>(nnz), *r_tmp = ws.alloc<UI>(nnz);
        // stable sort by column of the (row, col)-sorted list = the reference's (col, row) sort
        radix_sort<UI, UI, NoVal>(ws, {(UI *)col, (UI *)row, nullptr}, {k_out, r_out, nullptr},
                                  {k_tmp, r_tmp, nullptr}, nnz,
                                  {{0, bits_for((uint64_t)(nodes > 1 ? nodes - 1 : 1))}});
        const int64_t cap = (int64_t)device_info(device).sm_count * 16;
        const int64_t blocks = ceil_div(nnz, 256);
        SB_LAUNCH((boba_first_kernel<SI, UI>), (unsigned)(blocks < cap ? blocks : cap), 256, 0, st,
                  (const UI *)r_out, (const UI *)k_out, nnz, key);
      }
      const int rc = sb200_rank_keys(device, nodes, key, 2 * nnz + 1, out_inv, id_type, stream);
      if (rc != SB200_OK) throw Error{rc};
    });
  });
}

}  // extern "C"
