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
// Four consecutive rows per thread: the five row_ptr entries they need are one 16-byte load
// (when the array is 16-byte aligned) plus the first entry of the next thread, which comes by
// shuffle; the four results leave as one 16-byte streaming store (4-byte outputs).
constexpr int kDfPer = 4;

template <typename N>
__device__ __forceinline__ void df_load5(const N *__restrict__ row_ptr, int64_t n, int64_t i0,
                                         N (&p)[kDfPer + 1]) {
  const bool vec = sizeof(N) == 4 && i0 + kDfPer <= n &&
                   (reinterpret_cast<uintptr_t>(row_ptr) & 15) == 0;
  if (vec) {
    const uint4 q = *reinterpret_cast<const uint4 *>(row_ptr + i0);
    p[0] = (N)q.x;
    p[1] = (N)q.y;
    p[2] = (N)q.z;
    p[3] = (N)q.w;
  } else {
#pragma unroll
    for (int k = 0; k < kDfPer; k++) p[k] = i0 + k <= n ? row_ptr[i0 + k] : N(0);
  }
  // entry i0 + 4 = the next thread's first entry (lane 31 and the ragged end read it)
  const N nxt = __shfl_down_sync(0xffffffffu, p[0], 1);
  p[kDfPer] = nxt;
  if (lane_id() == 31 || i0 + 2 * kDfPer > n) p[kDfPer] = i0 + kDfPer <= n ? row_ptr[i0 + kDfPer] : N(0);
}

template <typename T>
__device__ __forceinline__ void df_store4(T *__restrict__ out, int64_t n, int64_t i0,
                                          const T (&v)[kDfPer]) {
  if (sizeof(T) == 4 && i0 + kDfPer <= n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    uint4 q;
    q.x = *reinterpret_cast<const unsigned *>(&v[0]);
    q.y = *reinterpret_cast<const unsigned *>(&v[1]);
    q.z = *reinterpret_cast<const unsigned *>(&v[2]);
    q.w = *reinterpret_cast<const unsigned *>(&v[3]);
    __stcs(reinterpret_cast<uint4 *>(out + i0), q);
  } else {
#pragma unroll
    for (int k = 0; k < kDfPer; k++)
      if (i0 + k < n) st_stream(out + i0 + k, v[k]);
  }
}

template <typename I, typename N>
__global__ void __launch_bounds__(kMapBlock)
    degrees_kernel(const N *__restrict__ row_ptr, int64_t n, I *__restrict__ out) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kDfPer;
  N p[kDfPer + 1];
  df_load5<N>(row_ptr, n, i0 < n ? i0 : n, p);  // whole warps take part in the shuffle
  if (i0 >= n) return;
  I v[kDfPer];
#pragma unroll
  for (int k = 0; k < kDfPer; k++) v[k] = (I)(p[k + 1] - p[k]);
  df_store4<I>(out, n, i0, v);
}

// dist[i] = (rows[i+1]-rows[i]) / (FeatureType)num_edges  -- the N-typed difference is
// converted to F and divided with IEEE round-to-nearest (no fast-math in this library).
template <typename N, typename F>
__global__ void __launch_bounds__(kMapBlock)
    degree_distribution_kernel(const N *__restrict__ row_ptr, int64_t n, N num_edges,
                               F *__restrict__ out) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kDfPer;
  N p[kDfPer + 1];
  df_load5<N>(row_ptr, n, i0 < n ? i0 : n, p);
  if (i0 >= n) return;
  F v[kDfPer];
#pragma unroll
  for (int k = 0; k < kDfPer; k++) {
    const N d = p[k + 1] - p[k];
    if constexpr (std::is_same_v<F, float>)
      v[k] = __fdiv_rn((float)d, (float)num_edges);
    else
      v[k] = __ddiv_rn((double)d, (double)num_edges);
  }
  df_store4<F>(out, n, i0, v);
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
// degree bucket from its end (degree_reorder.cc:42-46).  Slot p of the sort input holds vertex
// n-1-p, so that a STABLE sort by degree sees ids in descending order.
//
// Almost every vertex of a sparse matrix has a small degree, so the sort is one stable 256-way
// partition on the digit min(degree, 255), computed straight from row_ptr (no key / id arrays
// are materialised):
//   upsweep    per-chunk digit histograms                        reads row_ptr once
//   spine scan (scan.cuh)
//   downsweep  stable rank of every slot inside its bin; bins 0..254 hold one degree each, so
//              that rank is final and inv[n-1-p] is written directly (coalesced: ids are
//              monotone in p).  Slots of bin 255 (degree >= 255) go, in stable order, to a
//              compact (degree, id) list                          reads row_ptr, writes inv
//   tail       the compact list (a tiny fraction of n) is sorted by degree with the general
//              radix sort and ranked behind the first 255 bins.
// HBM traffic for a low-degree matrix: 2 * (n+1) * N + n * I  vs  (n+1) * N + n * I compulsory.
constexpr int kDgBlock = 256;
constexpr int kDgWarps = kDgBlock / 32;
constexpr int kDgIpt = 8;
constexpr int kDgTile = kDgBlock * kDgIpt;
constexpr int kDgBins = 256;

struct DgChunking {
  int64_t n, tiles;
  int nchunks;
  __host__ __device__ int64_t tile_begin(int c) const { return tiles * c / nchunks; }
  __host__ __device__ int64_t tile_end(int c) const { return tiles * (c + 1) / nchunks; }
};

template <typename N>
__device__ __forceinline__ unsigned long long dg_degree(const N *__restrict__ row_ptr, int64_t n,
                                                        int64_t p) {
  const int64_t u = n - 1 - p;
  return (unsigned long long)(row_ptr[u + 1] - row_ptr[u]);
}

template <typename N>
__global__ void __launch_bounds__(kDgBlock)
    degree_upsweep_kernel(const N *__restrict__ row_ptr, DgChunking ch,
                          int64_t *__restrict__ spine, unsigned long long *__restrict__ max_deg) {
  __shared__ unsigned hist[kDgWarps][kDgBins];
  __shared__ unsigned long long s_max[kDgWarps];
  const unsigned wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kDgWarps * kDgBins; i += kDgBlock) (&hist[0][0])[i] = 0;
  __syncthreads();
  const int c = blockIdx.x;
  const int64_t begin = ch.tile_begin(c) * kDgTile;
  int64_t end = ch.tile_end(c) * kDgTile;
  if (end > ch.n) end = ch.n;
  unsigned long long mx = 0;
  for (int64_t base = begin; base < end; base += (int64_t)kDgBlock * 4) {
    unsigned long long d[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int64_t p = base + k * kDgBlock + threadIdx.x;
      d[k] = p < end ? dg_degree(row_ptr, ch.n, p) : 0ull;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int64_t p = base + k * kDgBlock + threadIdx.x;
      if (p < end) {
        mx = d[k] > mx ? d[k] : mx;
        // warp-aggregated: one shared-memory atomic per distinct digit among the lanes
        const unsigned dig = d[k] < 255ull ? (unsigned)d[k] : 255u;
        const unsigned peers = __match_any_sync(__activemask(), dig);
        if ((int)lane_id() == __ffs(peers) - 1) atomicAdd(&hist[wid][dig], (unsigned)__popc(peers));
      }
    }
  }
  mx = warp_reduce_max(mx);
  if (lane_id() == 0) s_max[wid] = mx;
  __syncthreads();
  for (int b = threadIdx.x; b < kDgBins; b += kDgBlock) {
    unsigned sum = 0;
#pragma unroll
    for (int w = 0; w < kDgWarps; w++) sum += hist[w][b];
    spine[(int64_t)b * ch.nchunks + c] = sum;
  }
  if (threadIdx.x == 0) {  // one atomic per CTA, and only when it raises the maximum
    mx = 0;
    for (int w = 0; w < kDgWarps; w++) mx = s_max[w] > mx ? s_max[w] : mx;
    if (mx > *reinterpret_cast<volatile unsigned long long *>(max_deg)) atomicMax(max_deg, mx);
  }
}

// inv[id] = ascending ? pos : n-1-pos   (degree_reorder.cc:47-57)
template <typename I, typename N, typename HK>
__global__ void __launch_bounds__(kDgBlock)
    degree_downsweep_kernel(const N *__restrict__ row_ptr, DgChunking ch,
                            const int64_t *__restrict__ spine, int ascending,
                            I *__restrict__ inv, HK *__restrict__ high_key,
                            I *__restrict__ high_id) {
  __shared__ unsigned cnt[kDgWarps][kDgBins];
  __shared__ int64_t bin_off[kDgBins];
  __shared__ int64_t high_base;
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const int c = blockIdx.x;
  const int64_t n = ch.n;
  bin_off[threadIdx.x] = spine[(int64_t)threadIdx.x * ch.nchunks + c];  // kDgBlock == kDgBins
  if (threadIdx.x == 0) high_base = spine[(int64_t)255 * ch.nchunks];
  // the degrees of the next tile are requested before the current one is ranked
  // 32-bit degrees when row_ptr is 32-bit: half the registers of the two tiles in flight
  using DT = typename std::conditional<sizeof(N) == 4, unsigned, unsigned long long>::type;
  DT nd[kDgIpt];
  {
    const int64_t wb = ch.tile_begin(c) * kDgTile + (int64_t)wid * (kDgIpt * 32);
#pragma unroll
    for (int r = 0; r < kDgIpt; r++) {
      const int64_t p = wb + r * 32 + lane;
      nd[r] = (ch.tile_begin(c) < ch.tile_end(c) && p < n) ? (DT)dg_degree(row_ptr, n, p) : DT(0);
    }
  }
  for (int64_t tile = ch.tile_begin(c); tile < ch.tile_end(c); tile++) {
    const int64_t warp_base = tile * kDgTile + (int64_t)wid * (kDgIpt * 32);
    for (int i = threadIdx.x; i < kDgWarps * kDgBins; i += kDgBlock) (&cnt[0][0])[i] = 0;
    DT d[kDgIpt];
#pragma unroll
    for (int r = 0; r < kDgIpt; r++) d[r] = nd[r];
    if (tile + 1 < ch.tile_end(c)) {
#pragma unroll
      for (int r = 0; r < kDgIpt; r++) {
        const int64_t p = warp_base + kDgTile + r * 32 + lane;
        nd[r] = p < n ? (DT)dg_degree(row_ptr, n, p) : DT(0);
      }
    }
    __syncthreads();
    // stable rank inside the warp: rounds in order, lanes in order
    unsigned lp[kDgIpt];
#pragma unroll
    for (int r = 0; r < kDgIpt; r++) {
      const bool valid = warp_base + r * 32 + lane < n;
      const unsigned dig = valid ? (d[r] < DT(255) ? (unsigned)d[r] : 255u) : 0xffffffffu;
      const unsigned peers = __match_any_sync(0xffffffffu, dig);
      const int leader = __ffs(peers) - 1;
      unsigned base = 0;
      if ((int)lane == leader && valid) {
        base = cnt[wid][dig];
        cnt[wid][dig] = base + __popc(peers);
      }
      base = __shfl_sync(0xffffffffu, base, leader);
      lp[r] = base + __popc(peers & lanemask_lt());
      __syncwarp();
    }
    __syncthreads();
    {  // per digit: exclusive scan over the warps, advance the chunk's running bin offsets
      const int b = threadIdx.x;
      unsigned run = 0;
#pragma unroll
      for (int w = 0; w < kDgWarps; w++) {
        const unsigned t = cnt[w][b];
        cnt[w][b] = run;
        run += t;
      }
      // cnt[w][b] += old bin offset would need 64 bits; keep them apart
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kDgIpt; r++) {
        const int64_t p = warp_base + r * 32 + lane;
        if (p < n) {
          const unsigned dig = d[r] < DT(255) ? (unsigned)d[r] : 255u;
          const int64_t pos = bin_off[dig] + cnt[wid][dig] + lp[r];
          const int64_t u = n - 1 - p;
          if (dig < 255u) {
            inv[u] = (I)(ascending ? pos : n - 1 - pos);
          } else {
            high_key[pos - high_base] = (HK)d[r];
            high_id[pos - high_base] = (I)u;
          }
        }
      }
      __syncthreads();
      bin_off[b] += run;
    }
  }
}

template <typename I>
__global__ void degree_rank_kernel(const I *__restrict__ sorted, int64_t cnt, int64_t base,
                                   int64_t n, int ascending, I *__restrict__ inv) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < cnt) inv[sorted[q]] = (I)(ascending ? base + q : n - 1 - (base + q));
}

template <typename I, typename N>
void degree_reorder_impl(Workspace &ws, int64_t n, const N *row_ptr, bool ascending, I *out_inv) {
  if (n <= 0) return;
  cudaStream_t st = ws.stream();
  using UI = typename std::make_unsigned<I>::type;
  using HK = typename std::conditional<sizeof(N) == 8, uint64_t, uint32_t>::type;
  const DeviceInfo &di = device_info(ws.device());
  DgChunking ch;
  ch.n = n;
  ch.tiles = ceil_div(n, kDgTile);
  const int64_t max_chunks = (int64_t)di.sm_count * 8;
  ch.nchunks = (int)(ch.tiles < max_chunks ? ch.tiles : max_chunks);
  const int64_t spine_len = (int64_t)kDgBins * ch.nchunks;
  int64_t *spine_in = ws.alloc<int64_t>(spine_len + 1);
  int64_t *spine = ws.alloc<int64_t>(spine_len + 1);
  unsigned long long *max_deg = ws.alloc<unsigned long long>(1);
  SB_CUDA(cudaMemsetAsync(max_deg, 0, sizeof(unsigned long long), st));
  SB_LAUNCH((degree_upsweep_kernel<N>), ch.nchunks, kDgBlock, 0, st, row_ptr, ch, spine_in,
            max_deg);
  exclusive_scan<int64_t>(ws, LoadFn<int64_t>{spine_in}, spine, spine_len);
  // the vertices of bin 255 (degree >= 255): how many, and how wide their degrees are
  unsigned long long h_max = 0;
  int64_t h_base = 0;
  SB_CUDA(cudaMemcpyAsync(&h_max, max_deg, sizeof(h_max), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&h_base, spine + (int64_t)255 * ch.nchunks, sizeof(h_base),
                          cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const int64_t n_high = n - h_base;
  HK *hk = n_high > 0 ? ws.alloc<HK>(n_high) : nullptr;
  UI *hid = n_high > 0 ? ws.alloc<UI>(n_high) : nullptr;
  SB_LAUNCH((degree_downsweep_kernel<I, N, HK>), ch.nchunks, kDgBlock, 0, st, row_ptr, ch,
            (const int64_t *)spine, ascending ? 1 : 0, out_inv, hk, (I *)hid);
  if (n_high <= 0) return;
  const UI *sorted = hid;
  if (h_max > 255ull && n_high > 1) {
    // stable sort of the tail by degree; degrees below 255 do not occur here, so the low 8 bits
    // still matter (255 vs 256 ...) and the full width is sorted
    std::vector<RsBitRange> ranges = {{0, bits_for(h_max)}};
    const int P = rs_num_passes(ranges);
    HK *ko = ws.alloc<HK>(n_high), *kt = P > 1 ? ws.alloc<HK>(n_high) : nullptr;
    UI *io = ws.alloc<UI>(n_high), *it = P > 1 ? ws.alloc<UI>(n_high) : nullptr;
    radix_sort<HK, UI, NoVal>(ws, {hk, hid, nullptr}, {ko, io, nullptr}, {kt, it, nullptr}, n_high,
                              ranges);
    sorted = io;
  }
  SB_LAUNCH((degree_rank_kernel<I>), map_grid(n_high), kMapBlock, 0, st, (const I *)sorted, n_high,
            h_base, n, ascending ? 1 : 0, out_inv);
}

// ---------------------------------------------------------------- Permute2D
// Old row i becomes new row j = row_order[i].  One coalesced pass over xadj scatters, per new
// row, the source offset of its first entry and its length; the scan and the gather kernels
// then read both arrays coalesced (no dependent irow -> xadj gathers inside the hot kernels).
// The same pass finds the longest row, which selects the gather kernel.
// (source offset, length) of a new row, written and read as one word
template <typename N>
struct alignas(2 * sizeof(N)) RowRec {
  N base, len;
};
template <typename N>
struct RowLenFn {
  const RowRec<N> *rec;
  __device__ N operator()(int64_t i) const { return rec[i].len; }
};
constexpr int kPrepRows = 4;  // rows per thread: independent row_order loads in flight
template <typename I, typename N>
__global__ void __launch_bounds__(kMapBlock)
    permute_prepare_kernel(const N *__restrict__ xadj, const I *__restrict__ row_order, int64_t n,
                           RowRec<N> *__restrict__ rec,
                           unsigned long long *__restrict__ max_len, int64_t near_rows) {
  // max_len[0]: longest row.  max_len[1]: number of old rows i whose successor i+1 lands within
  // near_rows new rows of it -- when most do, rows that are neighbours in the SOURCE are gathered
  // close in time and share sectors in L2; when few do, the gathered rows are read with
  // evict-first loads (GatherLoader::stream).
  const int64_t i0 = (int64_t)blockIdx.x * (kMapBlock * kPrepRows) + threadIdx.x;
  int64_t j[kPrepRows];
  N b[kPrepRows], e[kPrepRows];
#pragma unroll
  for (int u = 0; u < kPrepRows; u++) {
    const int64_t i = i0 + u * kMapBlock;
    if (i < n) {
      j[u] = row_order ? (int64_t)row_order[i] : i;
      b[u] = xadj[i];
      e[u] = xadj[i + 1];
    }
  }
  unsigned long long len = 0;
  unsigned near = 0;
#pragma unroll
  for (int u = 0; u < kPrepRows; u++) {
    const int64_t i = i0 + u * kMapBlock;
    const int64_t jn = __shfl_down_sync(0xffffffffu, i < n ? j[u] : (int64_t)0, 1);
    if (i + 1 < n && lane_id() < 31) {
      const int64_t d = jn > j[u] ? jn - j[u] : j[u] - jn;
      near += d <= near_rows ? 1u : 0u;
    }
    if (i < n) {
      rec[j[u]] = RowRec<N>{b[u], (N)(e[u] - b[u])};  // one scattered store per row
      const unsigned long long l = (unsigned long long)(e[u] - b[u]);
      len = l > len ? l : len;
    }
  }
  // one atomic per CTA, and only when it would raise the maximum (same-address atomics
  // serialise in L2: one per warp cost 0.25 ms at 16.7 M rows)
  __shared__ unsigned long long s_max;
  __shared__ unsigned s_near;
  if (threadIdx.x == 0) {
    s_max = 0;
    s_near = 0;
  }
  __syncthreads();
  near = __reduce_add_sync(0xffffffffu, near);
  if (lane_id() == 0 && near > 0) atomicAdd(&s_near, near);
  if (__all_sync(0xffffffffu, len < (1ull << 32)))
    len = __reduce_max_sync(0xffffffffu, (unsigned)len);
  else
    len = warp_reduce_max(len);
  if (lane_id() == 0 && len > 0) atomicMax(&s_max, len);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long mx = s_max;
    if (mx > *reinterpret_cast<volatile unsigned long long *>(max_len)) atomicMax(max_len, mx);
    if (s_near > 0) atomicAdd(max_len + 1, (unsigned long long)s_near);
  }
}

template <typename I, typename N, typename V>
struct GatherLoader {
  const RowRec<N> *rec;  // per new row: offset of the old row's first entry (+ its length)
  const I *adj;
  const V *vals;
  const I *col_order;  // old col -> new col, or null
  bool stream;         // evict-first loads for the gathered rows
  __device__ int64_t seg_base(int64_t r) const { return (int64_t)rec[r].base; }
  // default cache policy on purpose: a gathered row shares its 32-byte sectors with the rows
  // next to it in the SOURCE, which are gathered a little later (evict-first loads made HBM
  // deliver those sectors twice)
  // (`stream` is set when the permutation scatters source neighbours far apart: nothing is
  // shared then, and evict-first keeps the gathered rows from displacing the col_order table)
  __device__ I raw_key(int64_t p) const { return stream ? __ldcs(adj + p) : __ldg(adj + p); }
  __device__ I map_key(I c) const { return col_order ? __ldg(col_order + c) : c; }
  __device__ V val(int64_t p) const { return stream ? __ldcs(vals + p) : __ldg(vals + p); }
};

// ---- matrices whose rows all have <= kShortRow entries (stencils, meshes).  One warp owns
//      32 consecutive NEW rows per step.  Their entries are fetched in concatenated order
//      (lane s reads the s-th entry of the 32-row batch: consecutive lanes read consecutive
//      addresses inside a source row), all loads of the batch before the first dependent
//      col_order gather, staged in a 2 KB per-warp slice of shared memory, ranked inside their
//      row by enumeration (<= 8 compares) and written straight to their sorted place -- the 32
//      rows are adjacent in the output, so the warp writes ONE contiguous run.  No block
//      barriers; three dependent loads per step (row records -> entries -> renumbered
//      columns). ----
constexpr int kShortRow = 8;
constexpr int kSrBlock = 256;

template <typename I, typename N, typename V, int MINB>
__global__ void __launch_bounds__(kSrBlock, MINB)
    permute_short_rows_kernel(const RowRec<N> *__restrict__ rec, const N *__restrict__ out_ptr,
                              const I *__restrict__ adj, const V *__restrict__ vals,
                              const I *__restrict__ col_order, int64_t n,
                              I *__restrict__ out_col, V *__restrict__ out_vals, DupCtx dc) {
  using VR = typename std::conditional<has_val<V>, V, char>::type;
  __shared__ I stage_k[kSrBlock / 32][32 * kShortRow];
  __shared__ VR stage_v[kSrBlock / 32][has_val<V> ? 32 * kShortRow : 1];
  __shared__ unsigned char stage_own[kSrBlock / 32][32 * kShortRow];
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  I *sk = stage_k[wid];
  [[maybe_unused]] VR *sv = stage_v[wid];
  unsigned char *own = stage_own[wid];
  const int64_t nbatches = (n + 31) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * kSrBlock) >> 5;
  bool unsorted = false;  // some row has an inversion in source order (csr.cc:99-116)
  for (int64_t bt = (((int64_t)blockIdx.x * kSrBlock) >> 5) + wid; bt < nbatches; bt += wstride) {
    const int64_t j = (bt << 5) + lane;
    N ob = 0, p = 0;
    unsigned len = 0;
    if (j < n) {
      const RowRec<N> rr = rec[j];
      ob = out_ptr[j];
      len = (unsigned)rr.len;
      p = rr.base;
    }
    const N ob0 = __shfl_sync(0xffffffffu, ob, 0);
    const unsigned incl = warp_inclusive_scan(len);
    const unsigned excl = incl - len;  // == ob - ob0 for the rows that exist
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    // every row tells its slots of the concatenated batch who owns them
#pragma unroll
    for (int u = 0; u < kShortRow; u++)
      if ((unsigned)u < len) own[excl + u] = (unsigned char)lane;
    __syncwarp();
    // ---- fetch in concatenated order: all the loads of the batch are issued before the
    //      first dependent renumbering gather, all the gathers before the first use ----
    I c[kShortRow];
    [[maybe_unused]] VR cv[kShortRow];
#pragma unroll
    for (int u = 0; u < kShortRow; u++) {
      const unsigned sidx = u * 32 + lane;
      const unsigned owner = sidx < total ? own[sidx] : 0u;
      const N p_o = __shfl_sync(0xffffffffu, p, owner);
      const unsigned ex_o = __shfl_sync(0xffffffffu, excl, owner);
      if (sidx < total) {
        const N src = p_o + (N)(sidx - ex_o);
        c[u] = __ldg(adj + src);
        if constexpr (has_val<V>) cv[u] = __ldg(vals + src);
      }
    }
    if (col_order) {
#pragma unroll
      for (int u = 0; u < kShortRow; u++)
        if (u * 32 + lane < total) c[u] = __ldg(col_order + c[u]);
    }
#pragma unroll
    for (int u = 0; u < kShortRow; u++) {
      const unsigned sidx = u * 32 + lane;
      if (sidx < total) {
        sk[sidx] = c[u];
        if constexpr (has_val<V>) sv[sidx] = cv[u];
      }
    }
    __syncwarp();
    // ---- rank of every entry inside its row (rows have <= kShortRow entries, all staged in
    //      this warp's slice); the entry goes straight to its sorted place.  The 32 rows are
    //      adjacent in the output, so the warp still writes one contiguous run. ----
#pragma unroll
    for (int u = 0; u < kShortRow; u++) {
      const unsigned sidx = u * 32 + lane;
      const unsigned owner = sidx < total ? own[sidx] : 0u;
      const unsigned ex_o = __shfl_sync(0xffffffffu, excl, owner);
      const unsigned len_o = __shfl_sync(0xffffffffu, len, owner);
      if (sidx < total) {
        const I k = c[u];
        // unique column ids (every valid input): the rank is the number of smaller ids.  The
        // loop also counts the ids <= k; more than one means duplicates: those are ranked in
        // source order and the row is flagged for the reference's tie rule (DupCtx).
        unsigned lt = 0, le = 0;
#pragma unroll
        for (int t = 0; t < kShortRow; t++) {
          if ((unsigned)t < len_o) {
            const I kj = sk[ex_o + t];
            lt += kj < k ? 1u : 0u;
            le += kj <= k ? 1u : 0u;
          }
        }
        // in source order the entry sits at sidx - ex_o; the row is non-decreasing iff every
        // entry lies inside the range of its equals
        const unsigned at = sidx - ex_o;
        if (at < lt || at >= le) unsorted = true;
        unsigned rank = lt;
        if (le - lt > 1u) {
          for (unsigned jj = ex_o; jj < sidx; jj++) rank += sk[jj] == k ? 1u : 0u;
          if (rank == lt + 1u) dc.flag((bt << 5) + owner);
        }
        st_stream(out_col + ob0 + ex_o + rank, k);
        if constexpr (has_val<V>) st_stream(out_vals + ob0 + ex_o + rank, (V)cv[u]);
      }
    }
    __syncwarp();
  }
  dc.report_unsorted(unsorted);
}

// ---- matrices whose longest row has <= 64 entries (random graphs of moderate degree, banded
//      matrices).  Same warp-owned scheme as the short-row kernel, without block barriers and
//      without the window bookkeeping of ss_tile_kernel: one warp owns R consecutive NEW rows
//      (R = 16 for rows <= 32, R = 8 for rows <= 64, so a batch has <= 512 entries), fetches
//      their entries in concatenated order four 32-slot rounds at a time (row gathers, then the
//      dependent col_order gathers, then the staging stores), ranks every entry inside its row
//      by enumeration in the warp's shared-memory slice and writes the batch as one contiguous
//      run.  ncu on C3 had ss_tile_kernel at 7 ms where a barrier-free gather of the same
//      access mix (profiles/ubench/gather.cu) takes 1.6 ms. ----
constexpr int kMrCap = 512;
// warps per CTA so that the static shared memory stays under 48 KB for 8-byte ids / values
template <typename I, typename V>
constexpr int mr_block() {
  return (sizeof(I) + (has_val<V> ? sizeof(V) : 0)) <= 8 ? 256 : 128;
}

template <typename I, typename N, typename V, int R>
__global__ void __launch_bounds__((mr_block<I, V>()))
    permute_mid_rows_kernel(const RowRec<N> *__restrict__ rec, const N *__restrict__ out_ptr,
                            const I *__restrict__ adj, const V *__restrict__ vals,
                            const I *__restrict__ col_order, int64_t n, bool stream,
                            I *__restrict__ out_col, V *__restrict__ out_vals, DupCtx dc) {
  using VR = typename std::conditional<has_val<V>, V, char>::type;
  constexpr int kMrBlock = mr_block<I, V>();
  bool unsorted = false;  // some row has an inversion in source order (csr.cc:99-116)
  __shared__ I stage_k[kMrBlock / 32][kMrCap];
  __shared__ VR stage_v[kMrBlock / 32][has_val<V> ? kMrCap : 1];
  __shared__ unsigned char stage_own[kMrBlock / 32][kMrCap];
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  I *sk = stage_k[wid];
  [[maybe_unused]] VR *sv = stage_v[wid];
  unsigned char *own = stage_own[wid];
  const int64_t nbatches = (n + R - 1) / R;
  const int64_t wstride = ((int64_t)gridDim.x * kMrBlock) >> 5;
  for (int64_t bt = (((int64_t)blockIdx.x * kMrBlock) >> 5) + wid; bt < nbatches; bt += wstride) {
    const int64_t j = bt * R + lane;
    N ob = 0, p = 0;
    unsigned len = 0;
    if (lane < (unsigned)R && j < n) {
      const RowRec<N> rr = rec[j];
      ob = out_ptr[j];
      len = (unsigned)rr.len;
      p = rr.base;
    }
    const N ob0 = __shfl_sync(0xffffffffu, ob, 0);
    const unsigned incl = warp_inclusive_scan(len);
    const unsigned excl = incl - len;
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);  // <= R * (512 / R)
    for (unsigned u = 0; u < len; u++) own[excl + u] = (unsigned char)lane;
    __syncwarp();
    // ---- fetch, four rounds in flight ----
    for (unsigned base = 0; base < total; base += 128) {
      I c[4];
      [[maybe_unused]] VR cv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const unsigned sidx = base + u * 32 + lane;
        const unsigned owner = sidx < total ? own[sidx] : 0u;
        const N p_o = __shfl_sync(0xffffffffu, p, owner);
        const unsigned ex_o = __shfl_sync(0xffffffffu, excl, owner);
        if (sidx < total) {
          const N src = p_o + (N)(sidx - ex_o);
          c[u] = stream ? __ldcs(adj + src) : __ldg(adj + src);
          if constexpr (has_val<V>) cv[u] = stream ? __ldcs(vals + src) : __ldg(vals + src);
        }
      }
      if (col_order) {
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (base + u * 32 + lane < total) c[u] = __ldg(col_order + c[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const unsigned sidx = base + u * 32 + lane;
        if (sidx < total) {
          sk[sidx] = c[u];
          if constexpr (has_val<V>) sv[sidx] = cv[u];
        }
      }
    }
    __syncwarp();
    // ---- rank inside the row, write to the sorted place ----
    for (unsigned base = 0; base < total; base += 32) {
      const unsigned sidx = base + lane;
      const unsigned owner = sidx < total ? own[sidx] : 0u;
      const unsigned ex_o = __shfl_sync(0xffffffffu, excl, owner);
      const unsigned len_o = __shfl_sync(0xffffffffu, len, owner);
      if (sidx < total) {
        const I k = sk[sidx];
        unsigned lt = 0, le = 0;  // see permute_short_rows_kernel
        const I *rowk = sk + ex_o;
        unsigned t = 0;
        for (; t + 4 <= len_o; t += 4) {
          const I k0 = rowk[t], k1 = rowk[t + 1], k2 = rowk[t + 2], k3 = rowk[t + 3];
          lt += (k0 < k ? 1u : 0u) + (k1 < k ? 1u : 0u) + (k2 < k ? 1u : 0u) + (k3 < k ? 1u : 0u);
          le += (k0 <= k ? 1u : 0u) + (k1 <= k ? 1u : 0u) + (k2 <= k ? 1u : 0u) +
                (k3 <= k ? 1u : 0u);
        }
        for (; t < len_o; t++) {
          const I kj = rowk[t];
          lt += kj < k ? 1u : 0u;
          le += kj <= k ? 1u : 0u;
        }
        const unsigned at = sidx - ex_o;  // see permute_short_rows_kernel
        if (at < lt || at >= le) unsorted = true;
        unsigned rank = lt;
        if (le - lt > 1u) {  // duplicate ids: source order here, the tie rule in dup_fix_kernel
          for (unsigned jj = ex_o; jj < sidx; jj++) rank += sk[jj] == k ? 1u : 0u;
          if (rank == lt + 1u) dc.flag(bt * R + owner);
        }
        st_stream(out_col + ob0 + ex_o + rank, k);
        if constexpr (has_val<V>) st_stream(out_vals + ob0 + ex_o + rank, (V)sv[sidx]);
      }
    }
    __syncwarp();
  }
  dc.report_unsorted(unsorted);
}

template <typename I, typename N, typename V>
void permute2d_impl(Workspace &ws, int64_t n, int64_t m, int64_t nnz, const N *xadj,
                    const I *adj, const V *vals, const I *row_order, const I *col_order,
                    N *out_row_ptr, I *out_col, V *out_vals, int vkind) {
  cudaStream_t st = ws.stream();
  RowRec<N> *rec = ws.alloc<RowRec<N>>(n + 1);
  unsigned long long *max_len = ws.alloc<unsigned long long>(2);
  SB_CUDA(cudaMemsetAsync(max_len, 0, 2 * sizeof(unsigned long long), st));
  // "near": within 32 MB worth of gathered rows (a quarter of the L2)
  const int64_t row_bytes = n > 0 ? (nnz / n + 1) * (int64_t)(sizeof(I) + (has_val<V> ? sizeof(V) : 0)) : 1;
  const int64_t near_rows = (32ll << 20) / row_bytes;
  if (n > 0)
    SB_LAUNCH((permute_prepare_kernel<I, N>), map_grid(n, kPrepRows), kMapBlock, 0, st, xadj,
              row_order, n, rec, max_len, near_rows);
  exclusive_scan<N>(ws, RowLenFn<N>{rec}, out_row_ptr, n);
  if (n <= 0 || nnz <= 0) return;
  unsigned long long h_stats[2] = {0, 0};
  SB_CUDA(cudaMemcpyAsync(h_stats, max_len, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const unsigned long long h_max = h_stats[0];
  // duplicate column ids (rare): flagged by the gather kernels, settled by dup_fix_kernel
  DupCtx dc = {nullptr, nullptr, nullptr};
  if constexpr (has_val<V>) {
    if (h_max <= 64ull) dc = make_dup_ctx(ws, n, true);
  }
  GatherLoader<I, N, V> fix_ld{rec, adj, vals, col_order, false};
  if (h_max <= (unsigned long long)kShortRow) {
    // one 32-row batch per warp, CTAs in row order: neighbouring rows are in flight at the
    // same time, so the sectors they share (20-byte rows in 32-byte sectors, nearby col_order
    // entries) are fetched from HBM once.  (A grid-stride loop over a capped grid spreads the
    // resident warps over distant row ranges and doubled the DRAM read traffic.)
    const unsigned grid = (unsigned)ceil_div(n, (int64_t)kSrBlock);
    static const int minb = [] {
      const char *e = getenv("SB200_SR_MINB");  // tuning: resident CTAs the kernel is compiled for
      return e ? atoi(e) : 6;
    }();
    // wide ids / values do not fit 40 registers
    constexpr bool narrow = sizeof(I) == 4 && sizeof(N) == 4 && (!has_val<V> || sizeof(V) == 4);
    if (minb == 5 || !narrow)
      SB_LAUNCH((permute_short_rows_kernel<I, N, V, 5>), grid, kSrBlock, 0, st,
                (const RowRec<N> *)rec, (const N *)out_row_ptr, adj, vals, col_order, n, out_col,
                out_vals, dc);
    else  // 40 registers: 48 resident warps per SM
      SB_LAUNCH((permute_short_rows_kernel<I, N, V, (narrow ? 6 : 5)>), grid, kSrBlock, 0, st,
                (const RowRec<N> *)rec, (const N *)out_row_ptr, adj, vals, col_order, n, out_col,
                out_vals, dc);
    launch_dup_fix<I, N, V>(ws, fix_ld, (const N *)out_row_ptr, n, dc, false, out_col, out_vals,
                            vkind);
    return;
  }
  static const int mid_env = [] {
    const char *e = getenv("SB200_P2D_MID");  // tuning: 0 disables the mid-row kernel
    return e ? atoi(e) : 1;
  }();
  static const int stream_env = [] {
    const char *e = getenv("SB200_P2D_STREAM");  // tuning: 0 = never, 1 = always
    return e ? atoi(e) : -1;
  }();
  const bool local = row_order == nullptr || 2 * h_stats[1] >= (unsigned long long)n;
  const bool stream = stream_env >= 0 ? stream_env != 0 : !local;
  if (h_max <= 64ull && mid_env) {
    constexpr int kMrBlock = mr_block<I, V>();
    if (h_max <= 32ull) {
      const unsigned grid = (unsigned)ceil_div(ceil_div(n, 16), kMrBlock / 32);
      SB_LAUNCH((permute_mid_rows_kernel<I, N, V, 16>), grid, kMrBlock, 0, st,
                (const RowRec<N> *)rec, (const N *)out_row_ptr, adj, vals, col_order, n, stream,
                out_col, out_vals, dc);
    } else {
      const unsigned grid = (unsigned)ceil_div(ceil_div(n, 8), kMrBlock / 32);
      SB_LAUNCH((permute_mid_rows_kernel<I, N, V, 8>), grid, kMrBlock, 0, st,
                (const RowRec<N> *)rec, (const N *)out_row_ptr, adj, vals, col_order, n, stream,
                out_col, out_vals, dc);
    }
    launch_dup_fix<I, N, V>(ws, fix_ld, (const N *)out_row_ptr, n, dc, false, out_col, out_vals,
                            vkind);
    return;
  }
  GatherLoader<I, N, V> ld{rec, adj, vals, col_order, stream};
  segmented_sort<I, N, V>(ws, ld, (const N *)out_row_ptr, n, m, nnz, out_col, out_vals, vkind,
                          false);
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
      SB_LAUNCH((degrees_kernel<I, N>), map_grid(n, kDfPer), kMapBlock, 0, (cudaStream_t)stream,
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
      SB_LAUNCH((degree_distribution_kernel<N, F>), map_grid(n, kDfPer), kMapBlock, 0, st,
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
                              (N *)out_row_ptr, (I *)out_col, (V *)out_vals, val_type);
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
