// convert.cu -- format constructors and order-two conversions of the C ABI.
//
//   sb200_coo_sort         format/coo.cc:96-157
//   sb200_compressed_sort  format/csr.cc:99-157, format/csc.cc:99-157
//   sb200_coo_to_csr       converter/converter_order_two.cc:162-212
//   sb200_csr_to_coo       converter/converter_order_two.cc:71-118
//   sb200_coo_to_csc       converter/converter_order_two.cc:20-70
//   sb200_csr_to_csc       converter/converter_order_two.cc:119-128
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "segsort.cuh"

namespace sb200 {

// =====================================================================================
// Sorted-index stream -> pointer array, fused with the col/val copy.
//
// For a stream idx[0..nnz) that is non-decreasing, the reference's "histogram, inclusive scan,
// shift right" (converter_order_two.cc:180-192) equals: ptr[r] = first position whose index is
// >= r.  Each thread looks at one boundary (idx[i-1], idx[i]) and writes the ptr entries in
// (idx[i-1], idx[i]]; the copies of the other arrays ride along in the same pass, so the whole
// COO->CSR conversion is ONE kernel reading 3 arrays and writing 2 (+ row_ptr).
// flags[0] |= 1 if the stream has an inversion (then ptr is rebuilt by the histogram path);
// flags[1] |= 1 if a (idx equal, sec decreasing) pair exists, i.e. some segment of the
// compressed result is unsorted (the CSR-constructor check of csr.cc:99-116).
// =====================================================================================
constexpr int kBfBlock = 256;
constexpr int kBfIpt = 4;

// A run of kGapMax or more empty segments (isolated vertices, or the columns a row block of
// a sharded matrix never touches) is not filled by the one thread that sees the boundary: it
// is appended to a list and filled afterwards by gap_fill_kernel with the whole grid.  (One
// thread filling 8.4 M trailing entries cost 30 ms in the row-block CSR->CSC.)
constexpr int kGapMax = 1024;        // shorter runs are filled in place (a power-law matrix has
                                     // millions of short runs: deferring them all through one
                                     // list counter cost 26 ms on R-MAT-26)
constexpr int kGapWarpMax = 32768;   // longer gaps are filled by all CTAs together
struct GapRec {
  int64_t lo, hi;  // ptr[lo..hi] = value
  int64_t value;
};
struct GapList {
  GapRec *rec;
  unsigned *count;
  unsigned cap;  // exact bound for a sorted stream; a stream with inversions may overrun it,
                 // and its ptr is rebuilt by the histogram path anyway
};
template <typename N>
__device__ __forceinline__ void fill_or_defer(N *__restrict__ ptr, int64_t lo, int64_t hi,
                                              int64_t value, const GapList &gl) {
  if (hi - lo >= kGapMax) {
    const unsigned k = atomicAdd(gl.count, 1u);
    if (k < gl.cap) gl.rec[k] = GapRec{lo, hi, value};
  } else {
    for (int64_t r = lo; r <= hi; r++) ptr[r] = (N)value;
  }
}
template <typename N>
__global__ void __launch_bounds__(256) gap_fill_kernel(GapList gl, N *__restrict__ ptr) {
  const unsigned cnt = *gl.count < gl.cap ? *gl.count : gl.cap;
  const unsigned lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp; k < cnt; k += nwarps) {  // medium gaps: one warp each
    const GapRec g = gl.rec[k];
    if (g.hi - g.lo >= kGapWarpMax) continue;
    for (int64_t r = g.lo + lane; r <= g.hi; r += 32) ptr[r] = (N)g.value;
  }
  for (unsigned k = 0; k < cnt; k++) {  // long gaps (at most n / kGapWarpMax): whole grid
    const GapRec g = gl.rec[k];
    if (g.hi - g.lo < kGapWarpMax) continue;
    for (int64_t r = g.lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= g.hi;
         r += (int64_t)gridDim.x * blockDim.x)
      ptr[r] = (N)g.value;
  }
}

template <typename I, typename N, typename V>
__global__ void __launch_bounds__(kBfBlock)
    boundary_fill_copy_kernel(const I *__restrict__ idx, const I *__restrict__ sec,
                              const V *__restrict__ vals, int64_t nnz, int64_t n_seg,
                              I idx_base, N *__restrict__ ptr, I *__restrict__ out_sec,
                              V *__restrict__ out_vals, unsigned *__restrict__ flags,
                              int64_t start, GapList gl) {
  const int64_t base = start + ((int64_t)blockIdx.x * kBfBlock) * kBfIpt;
  bool inv = false, sec_unsorted = false;
  I my[kBfIpt], prev[kBfIpt], ms[kBfIpt], ps[kBfIpt];
#pragma unroll
  for (int k = 0; k < kBfIpt; k++) {
    const int64_t i = base + (int64_t)k * kBfBlock + threadIdx.x;
    if (i < nnz) {
      my[k] = ld_stream(idx + i) - idx_base;
      prev[k] = i > 0 ? idx[i - 1] - idx_base : I(0);
      if (sec) {
        ms[k] = ld_stream(sec + i);
        ps[k] = i > 0 ? sec[i - 1] : I(0);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kBfIpt; k++) {
    const int64_t i = base + (int64_t)k * kBfBlock + threadIdx.x;
    if (i < nnz) {
      if (my[k] < prev[k]) inv = true;
      if (sec) {
        // CSR ctor check: inside a segment, sec must be non-decreasing starting from 0
        if (i > 0 && my[k] == prev[k] && ms[k] < ps[k]) sec_unsorted = true;
        if (ms[k] < 0) sec_unsorted = true;
        if (out_sec) st_stream(out_sec + i, ms[k]);
      }
      if constexpr (has_val<V>) {
        if (out_vals) st_stream(out_vals + i, ld_stream(vals + i));
      }
      // ptr entries for segments (prev, my]; the first element also covers segment 0
      int64_t lo = i > 0 ? (int64_t)prev[k] + 1 : 0;
      int64_t hi = (int64_t)my[k];
      if (hi >= n_seg) hi = n_seg - 1;  // indices >= n_seg are not representable (see header)
      if (lo < 0) lo = 0;
      fill_or_defer<N>(ptr, lo, hi, i, gl);
      if (i == nnz - 1) fill_or_defer<N>(ptr, (int64_t)my[k] + 1 < 0 ? 0 : (int64_t)my[k] + 1, n_seg, nnz, gl);
    }
  }
  if (__any_sync(0xffffffffu, inv) && lane_id() == 0) atomicOr(&flags[0], 1u);
  if (__any_sync(0xffffffffu, sec_unsorted) && lane_id() == 0) atomicOr(&flags[1], 1u);
}

// ---- 128-bit variant: every thread owns kBvLoads * (16 / sizeof(I)) CONSECUTIVE elements and
//      moves them with 16-byte loads and stores, all loads of the thread issued before the first
//      use (48-96 bytes in flight per thread).  The element before a thread's first one comes
//      from the neighbouring lane by shuffle (lane 0 reads it from L2).  Needs 16-byte aligned
//      arrays; handles the first (nnz / E) * E elements, the scalar kernel finishes the tail. ----
constexpr int kBvBlock = 256;
constexpr int kBvLoads = 2;

template <typename I, typename N, typename V>
__global__ void __launch_bounds__(kBvBlock)
    boundary_fill_copy_vec_kernel(const I *__restrict__ idx, const I *__restrict__ sec,
                                  const V *__restrict__ vals, int64_t nnz, int64_t ngroups,
                                  int64_t n_seg, I idx_base, N *__restrict__ ptr,
                                  I *__restrict__ out_sec, V *__restrict__ out_vals,
                                  unsigned *__restrict__ flags, GapList gl) {
  constexpr int kPer = 16 / sizeof(I);      // elements per 16-byte word
  constexpr int E = kBvLoads * kPer;        // elements per thread
  using VR = typename std::conditional<has_val<V>, V, I>::type;
  constexpr int kVw = has_val<V> ? (E * (int)sizeof(VR)) / 16 : 1;  // 16-byte words of values
  const int64_t g = (int64_t)blockIdx.x * kBvBlock + threadIdx.x;
  const bool active = g < ngroups;
  const int64_t i0 = g * E;
  union W {
    uint4 q;
    I e[kPer];
  };
  W a[kBvLoads], b[kBvLoads];
  uint4 v[kVw];
  if (active) {
#pragma unroll
    for (int k = 0; k < kBvLoads; k++) a[k].q = __ldcs(reinterpret_cast<const uint4 *>(idx + i0) + k);
    if (sec) {
#pragma unroll
      for (int k = 0; k < kBvLoads; k++)
        b[k].q = __ldcs(reinterpret_cast<const uint4 *>(sec + i0) + k);
    }
    if constexpr (has_val<V>) {
      if (out_vals) {
#pragma unroll
        for (int k = 0; k < kVw; k++) v[k] = __ldcs(reinterpret_cast<const uint4 *>(vals + i0) + k);
      }
    }
  }
  // the element before this thread's first one
  I last_idx = active ? a[kBvLoads - 1].e[kPer - 1] : I(0);
  I last_sec = (active && sec) ? b[kBvLoads - 1].e[kPer - 1] : I(0);
  I prev_idx = __shfl_up_sync(0xffffffffu, last_idx, 1);
  I prev_sec = __shfl_up_sync(0xffffffffu, last_sec, 1);
  if (active && lane_id() == 0) {
    prev_idx = i0 > 0 ? idx[i0 - 1] : idx_base;  // (0 after the base is subtracted)
    prev_sec = (i0 > 0 && sec) ? sec[i0 - 1] : I(0);
  }
  bool inv = false, sec_unsorted = false;
  if (active) {
    if (sec && out_sec) {
#pragma unroll
      for (int k = 0; k < kBvLoads; k++) __stcs(reinterpret_cast<uint4 *>(out_sec + i0) + k, b[k].q);
    }
    if constexpr (has_val<V>) {
      if (out_vals) {
#pragma unroll
        for (int k = 0; k < kVw; k++) __stcs(reinterpret_cast<uint4 *>(out_vals + i0) + k, v[k]);
      }
    }
    I p = prev_idx - idx_base, ps = prev_sec;
#pragma unroll
    for (int k = 0; k < E; k++) {
      const int64_t i = i0 + k;
      const I my = a[k / kPer].e[k % kPer] - idx_base;
      if (my < p) inv = true;
      if (sec) {
        const I ms = b[k / kPer].e[k % kPer];
        if (i > 0 && my == p && ms < ps) sec_unsorted = true;
        if (ms < 0) sec_unsorted = true;
        ps = ms;
      }
      if (my != p || i == 0) {
        int64_t lo = i > 0 ? (int64_t)p + 1 : 0;
        int64_t hi = (int64_t)my;
        if (hi >= n_seg) hi = n_seg - 1;
        if (lo < 0) lo = 0;
        fill_or_defer<N>(ptr, lo, hi, i, gl);
      }
      if (i == nnz - 1) fill_or_defer<N>(ptr, (int64_t)my + 1 < 0 ? 0 : (int64_t)my + 1, n_seg, nnz, gl);
      p = my;
    }
  }
  if (__any_sync(0xffffffffu, inv) && lane_id() == 0) atomicOr(&flags[0], 1u);
  if (__any_sync(0xffffffffu, sec_unsorted) && lane_id() == 0) atomicOr(&flags[1], 1u);
}

template <typename N>
__global__ void fill_kernel(N *p, int64_t cnt, N v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

// ---- fallback for a stream with inversions (COO built with ignore_sort=true): the literal
//      row histogram with warp-aggregated atomics, then the scan ----
template <typename I>
__global__ void histogram_kernel(const I *__restrict__ idx, int64_t nnz, int64_t n_seg,
                                 I idx_base, unsigned long long *__restrict__ hist) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x) {
    const I r = idx[i] - idx_base;
    if ((int64_t)r >= n_seg || r < 0) continue;
    // warp-aggregated: one atomic per distinct row among the active lanes
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, r);
    if ((int)lane_id() == __ffs(peers) - 1)
      atomicAdd(&hist[r], (unsigned long long)__popc(peers));
  }
}

template <typename N>
struct HistFn {
  const unsigned long long *h;
  __device__ N operator()(int64_t i) const { return (N)h[i]; }
};

// Builds ptr[n_seg+1] from idx (+ optional fused copies), honouring unsorted input.
// Returns flags {stream had inversions, some compressed segment is unsorted}.
template <typename I, typename N, typename V>
void build_ptr_and_copy(Workspace &ws, const I *idx, const I *sec, const V *vals, int64_t nnz,
                        int64_t n_seg, N *ptr, I *out_sec, V *out_vals, unsigned h_flags[2],
                        I idx_base = 0) {
  cudaStream_t st = ws.stream();
  unsigned *flags = ws.alloc<unsigned>(2);
  SB_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(unsigned), st));
  if (nnz == 0) {
    SB_LAUNCH((fill_kernel<N>), 64, 256, 0, st, ptr, n_seg + 1, N(0));
    h_flags[0] = h_flags[1] = 0;
    return;
  }
  // every deferred gap covers more than kGapMax segments of [0, n_seg], so this bounds the list
  GapList gl;
  gl.cap = (unsigned)((n_seg + 1) / kGapMax + 2);
  gl.rec = ws.alloc<GapRec>(gl.cap);
  gl.count = ws.alloc<unsigned>(1);
  SB_CUDA(cudaMemsetAsync(gl.count, 0, sizeof(unsigned), st));
  // 16-byte path for the bulk, scalar path for the tail (or everything if unaligned)
  constexpr int kE = kBvLoads * (16 / (int)sizeof(I));
  auto aligned16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  int64_t done = 0;
  if (aligned16(idx) && aligned16(sec) && aligned16(vals) && aligned16(out_sec) &&
      aligned16(out_vals) && nnz >= kE) {
    const int64_t ngroups = nnz / kE;
    SB_LAUNCH((boundary_fill_copy_vec_kernel<I, N, V>), (unsigned)ceil_div(ngroups, kBvBlock),
              kBvBlock, 0, st, idx, sec, vals, nnz, ngroups, n_seg, idx_base, ptr, out_sec,
              out_vals, flags, gl);
    done = ngroups * kE;
  }
  if (done < nnz) {
    const int64_t per_block = (int64_t)kBfBlock * kBfIpt;
    SB_LAUNCH((boundary_fill_copy_kernel<I, N, V>), (unsigned)ceil_div(nnz - done, per_block),
              kBfBlock, 0, st, idx, sec, vals, nnz, n_seg, idx_base, ptr, out_sec, out_vals, flags,
              done, gl);
  }
  SB_LAUNCH((gap_fill_kernel<N>), device_info(ws.device()).sm_count * 4, 256, 0, st, gl, ptr);
  SB_CUDA(cudaMemcpyAsync(h_flags, flags, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (h_flags[0]) {
    unsigned long long *hist = ws.alloc<unsigned long long>(n_seg + 1);
    SB_CUDA(cudaMemsetAsync(hist, 0, (n_seg + 1) * sizeof(unsigned long long), st));
    SB_LAUNCH((histogram_kernel<I>), device_info(ws.device()).sm_count * 8, 256, 0, st, idx, nnz,
              n_seg, idx_base, hist);
    exclusive_scan<N>(ws, HistFn<N>{hist}, ptr, n_seg);
  }
}

// =====================================================================================
// CSR/CSC constructor: check + segmented sort
// =====================================================================================
template <typename I, typename N>
__global__ void segments_sorted_check_kernel(const N *__restrict__ ptr,
                                             const I *__restrict__ idx, int64_t n_seg,
                                             unsigned *__restrict__ flag) {
  // one warp per segment, lanes stride over its entries (coalesced)
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  bool bad = false;
  for (int64_t r = warp; r < n_seg; r += nwarps) {
    const int64_t b = ptr[r], e = ptr[r + 1];
    for (int64_t j = b + lane; j < e; j += 32) {
      const I cur = idx[j];
      const I prev = j > b ? idx[j - 1] : I(0);  // csr.cc:107 -- prev_value starts at 0
      if (cur < prev) bad = true;
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flag, 1u);
}

template <typename I, typename N, typename V>
struct InPlaceLoader {
  const N *ptr;
  const I *idx;
  const V *vals;
  __device__ int64_t seg_base(int64_t r) const { return (int64_t)ptr[r]; }
  __device__ I raw_key(int64_t p) const { return idx[p]; }
  __device__ I map_key(I c) const { return c; }
  __device__ V val(int64_t p) const { return vals[p]; }
};

template <typename I, typename N, typename V>
bool compressed_check(Workspace &ws, const N *ptr, const I *idx, int64_t n_seg) {
  cudaStream_t st = ws.stream();
  unsigned *flag = ws.alloc<unsigned>(1);
  SB_CUDA(cudaMemsetAsync(flag, 0, sizeof(unsigned), st));
  if (n_seg > 0) {
    int64_t blocks = ceil_div(n_seg * 32, 256);
    int64_t cap = (int64_t)device_info(ws.device()).sm_count * 32;
    SB_LAUNCH((segments_sorted_check_kernel<I, N>), (unsigned)(blocks < cap ? blocks : cap), 256,
              0, st, ptr, idx, n_seg, flag);
  }
  unsigned h = 0;
  SB_CUDA(cudaMemcpyAsync(&h, flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return h == 0;
}

template <typename I, typename N, typename V>
void compressed_sort_inplace(Workspace &ws, const N *ptr, I *idx, V *vals, int64_t n_seg,
                             int64_t n_idx, int64_t nnz, int vkind) {
  // the caller ran the constructor's check: some segment is unsorted, so EVERY segment is
  // sorted by (index, value) -- values compared in their real type `vkind`
  InPlaceLoader<I, N, V> ld{ptr, idx, vals};
  segmented_sort<I, N, V>(ws, ld, ptr, n_seg, n_idx, nnz, idx, vals, vkind, true);
}

// =====================================================================================
// COO constructor sort
// =====================================================================================
template <typename I>
__global__ void coo_sorted_check_kernel(const I *__restrict__ row, const I *__restrict__ col,
                                        int64_t nnz, unsigned *__restrict__ flag) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x) {
    const I r = row[i], c = col[i];
    const I pr = i > 0 ? row[i - 1] : I(0), pc = i > 0 ? col[i - 1] : I(0);  // coo.cc:97-98
    if (pr > r || (pr == r && pc > c)) bad = true;
  }
  if (__any_sync(0xffffffffu, bad) && lane_id() == 0) atomicOr(flag, 1u);
}

template <typename I>
__global__ void coo_pack_kernel(const I *__restrict__ row, const I *__restrict__ col, int64_t nnz,
                                int col_bits, uint64_t *__restrict__ keys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = ((uint64_t)row[i] << col_bits) | (uint64_t)col[i];
}
template <typename I>
__global__ void coo_unpack_kernel(const uint64_t *__restrict__ keys, int64_t nnz, int col_bits,
                                  I *__restrict__ row, I *__restrict__ col) {
  const uint64_t mask = col_bits >= 64 ? ~0ull : ((1ull << col_bits) - 1ull);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    row[i] = (I)(k >> col_bits);
    col[i] = (I)(k & mask);
  }
}

template <typename I, typename V>
bool coo_sort_impl(Workspace &ws, int64_t n, int64_t m, int64_t nnz, I *row, I *col, V *vals) {
  cudaStream_t st = ws.stream();
  if (nnz <= 1) return true;
  const int grid = device_info(ws.device()).sm_count * 8;
  unsigned *flag = ws.alloc<unsigned>(1);
  SB_CUDA(cudaMemsetAsync(flag, 0, sizeof(unsigned), st));
  SB_LAUNCH((coo_sorted_check_kernel<I>), grid, 256, 0, st, row, col, nnz, flag);
  unsigned h = 0;
  SB_CUDA(cudaMemcpyAsync(&h, flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (h == 0) return true;
  const int col_bits = bits_for((uint64_t)(m > 0 ? m - 1 : 0));
  const int row_bits = bits_for((uint64_t)(n > 0 ? n - 1 : 0));
  SB_REQUIRE(col_bits + row_bits <= 64, SB200_ERR_BAD_ARG,
             "(row,col) key needs %d bits; at most 64 supported", col_bits + row_bits);
  uint64_t *k0 = ws.alloc<uint64_t>(nnz), *k1 = ws.alloc<uint64_t>(nnz),
           *k2 = ws.alloc<uint64_t>(nnz);
  SB_LAUNCH((coo_pack_kernel<I>), grid, 256, 0, st, row, col, nnz, col_bits, k0);
  std::vector<RsBitRange> ranges = {{0, col_bits + row_bits}};
  V *v1 = nullptr, *v2 = nullptr;
  if constexpr (has_val<V>) {
    if (vals) {
      v1 = ws.alloc<V>(nnz);
      v2 = ws.alloc<V>(nnz);
    }
  }
  if constexpr (has_val<V>) {
    if (vals) {
      radix_sort<uint64_t, V, NoVal>(ws, {k0, vals, nullptr}, {k1, v1, nullptr},
                                     {k2, v2, nullptr}, nnz, ranges);
      SB_CUDA(cudaMemcpyAsync(vals, v1, nnz * sizeof(V), cudaMemcpyDeviceToDevice, st));
    } else {
      radix_sort<uint64_t, NoVal, NoVal>(ws, {k0, nullptr, nullptr}, {k1, nullptr, nullptr},
                                         {k2, nullptr, nullptr}, nnz, ranges);
    }
  } else {
    radix_sort<uint64_t, NoVal, NoVal>(ws, {k0, nullptr, nullptr}, {k1, nullptr, nullptr},
                                       {k2, nullptr, nullptr}, nnz, ranges);
  }
  SB_LAUNCH((coo_unpack_kernel<I>), grid, 256, 0, st, (const uint64_t *)k1, nnz, col_bits, row,
            col);
  return false;
}

// =====================================================================================
// CSR -> COO : expand row_ptr into a row index per nonzero (+ fused col/val copy)
// =====================================================================================
constexpr int kExBlock = 256;
constexpr int kExTile = 2048;

template <typename N>
__global__ void ex_tile_bounds_kernel(const N *__restrict__ ptr, int64_t n_seg, int64_t ntiles,
                                      int64_t *__restrict__ tile_seg) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  if (t == ntiles) {
    tile_seg[t] = n_seg;
    return;
  }
  const int64_t target = t * kExTile;
  int64_t lo = 0, hi = n_seg;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)ptr[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  tile_seg[t] = lo;
}

// Window t covers output positions [t*kExTile, (t+1)*kExTile).  Segments starting inside the
// window mark their first position with (r - r0 + 1); an inclusive max-scan spreads the mark;
// positions before the first mark belong to the segment that straddles the window start
// (r0 - 1).  Every thread owns 8 consecutive positions: marks are zeroed, scanned and turned
// into row ids in registers, and each output array is written with two 16-byte stores per
// thread (4-byte ids; the copies of col / vals move the same way).
template <typename T>
__device__ __forceinline__ void ex_copy8(const T *__restrict__ src, T *__restrict__ dst, int64_t w0,
                                         int q0, int count) {
  if (q0 + 8 <= count && ((reinterpret_cast<uintptr_t>(src + w0) |
                           reinterpret_cast<uintptr_t>(dst + w0)) & 15) == 0) {
    constexpr int kVec = 16 / sizeof(T);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + w0 + q0);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + w0 + q0);
#pragma unroll
    for (int k = 0; k < 8 / kVec; k++) __stcs(d4 + k, __ldcs(s4 + k));
  } else {
    for (int k = 0; k < 8 && q0 + k < count; k++)
      st_stream(dst + w0 + q0 + k, ld_stream(src + w0 + q0 + k));
  }
}

template <typename I, typename N, typename V>
__global__ void __launch_bounds__(kExBlock)
    expand_ptr_kernel(const N *__restrict__ ptr, const int64_t *__restrict__ tile_seg,
                      int64_t nnz, const I *__restrict__ col, const V *__restrict__ vals,
                      I *__restrict__ out_row, I *__restrict__ out_col,
                      V *__restrict__ out_vals, I row_base) {
  static_assert(kExTile == kExBlock * 8, "8 positions per thread");
  __shared__ __align__(16) unsigned mark[kExTile];
  __shared__ unsigned warp_max[kExBlock / 32];
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const int64_t t = blockIdx.x;
  const int64_t w0 = t * kExTile;
  const int count = (int)(nnz - w0 < kExTile ? nnz - w0 : kExTile);
  const int64_t r0 = tile_seg[t], r1 = tile_seg[t + 1];
  const int q0 = threadIdx.x * 8;
  uint4 *m4 = reinterpret_cast<uint4 *>(mark + q0);
  m4[0] = make_uint4(0u, 0u, 0u, 0u);
  m4[1] = make_uint4(0u, 0u, 0u, 0u);
  // the copies do not depend on the row ids: get them in flight first
  if (out_col) ex_copy8<I>(col, out_col, w0, q0, count);
  if constexpr (has_val<V>) {
    if (out_vals) ex_copy8<V>(vals, out_vals, w0, q0, count);
  }
  __syncthreads();
  // r1 - r0 (segments starting in the window, empty ones included) < 2^32: checked by the host
  for (int64_t r = r0 + threadIdx.x; r < r1; r += kExBlock) {
    const int64_t b = ptr[r], e = ptr[r + 1];
    if (e > b) mark[b - w0] = (unsigned)(r - r0 + 1);  // non-empty segments: distinct starts
  }
  __syncthreads();
  uint4 a = m4[0], c = m4[1];
  unsigned h[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
  unsigned m = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) m = h[k] > m ? h[k] : m;
  unsigned inc = m;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned tt = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)lane >= o) inc = tt > inc ? tt : inc;
  }
  if (lane == 31) warp_max[wid] = inc;
  __syncthreads();
  unsigned run = 0;  // 0 = the segment straddling the window start (r0 - 1)
  for (unsigned w = 0; w < wid; w++) run = warp_max[w] > run ? warp_max[w] : run;
  const unsigned prev = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane > 0) run = prev > run ? prev : run;
  I rows[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    run = h[k] > run ? h[k] : run;
    rows[k] = (I)(r0 - 1 + (int64_t)run) + row_base;
  }
  if (q0 + 8 <= count && (reinterpret_cast<uintptr_t>(out_row + w0) & 15) == 0 &&
      sizeof(I) == 4) {
    uint4 *d4 = reinterpret_cast<uint4 *>(out_row + w0 + q0);
    __stcs(d4, make_uint4((unsigned)rows[0], (unsigned)rows[1], (unsigned)rows[2],
                          (unsigned)rows[3]));
    __stcs(d4 + 1, make_uint4((unsigned)rows[4], (unsigned)rows[5], (unsigned)rows[6],
                              (unsigned)rows[7]));
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (q0 + k < count) st_stream(out_row + w0 + q0 + k, rows[k]);
  }
}

template <typename I, typename N, typename V>
void expand_ptr(Workspace &ws, const N *ptr, int64_t n_seg, int64_t nnz, const I *col,
                const V *vals, I *out_row, I *out_col, V *out_vals, I row_base = 0) {
  if (nnz <= 0) return;
  SB_REQUIRE(n_seg < (1ll << 32) - 1, SB200_ERR_BAD_ARG,
             "row expansion supports fewer than 2^32-1 segments (got %lld)", (long long)n_seg);
  cudaStream_t st = ws.stream();
  const int64_t ntiles = ceil_div(nnz, kExTile);
  int64_t *tile_seg = ws.alloc<int64_t>(ntiles + 1);
  SB_LAUNCH((ex_tile_bounds_kernel<N>), (unsigned)ceil_div(ntiles + 1, 256), 256, 0, st, ptr,
            n_seg, ntiles, tile_seg);
  SB_LAUNCH((expand_ptr_kernel<I, N, V>), (unsigned)ntiles, kExBlock, 0, st, ptr,
            (const int64_t *)tile_seg, nnz, col, vals, out_row, out_col, out_vals, row_base);
}

// =====================================================================================
// COO -> CSC core: stable sort of (row, val) by column + col_ptr from the sorted keys
// =====================================================================================
template <typename I, typename N, typename V>
void to_csc_core(Workspace &ws, int64_t n, int64_t m, int64_t nnz, const I *row, const I *col,
                 const V *vals, N *out_col_ptr, I *out_row, V *out_vals, int vkind,
                 bool rows_ascending = false) {
  // n = number of col_ptr segments: dims[0] for the reference layout (square assumption),
  // m for the row-block variant used by the multi-GPU path
  using UI = typename std::make_unsigned<I>::type;
  cudaStream_t st = ws.stream();
  SB_REQUIRE(m <= n, SB200_ERR_BAD_ARG,
             "CSC of an n x m matrix with m > n is not representable in the reference layout "
             "(col_ptr has n+1 entries, converter_order_two.cc:32); got n=%lld m=%lld",
             (long long)n, (long long)m);
  unsigned h_flags[2] = {0, 0};
  if (nnz == 0) {
    build_ptr_and_copy<I, N, NoVal>(ws, nullptr, nullptr, nullptr, 0, n, out_col_ptr, nullptr,
                                    nullptr, h_flags);
    return;
  }
  const int col_bits = bits_for((uint64_t)(m > 0 ? m - 1 : 0));
  std::vector<RsBitRange> ranges = {{0, col_bits}};
  const int P = nnz > 1 ? rs_num_passes(ranges) : 0;
  UI *k_out = ws.alloc<UI>(nnz);
  UI *k_tmp = P > 1 ? ws.alloc<UI>(nnz) : nullptr;
  UI *r_tmp = P > 1 ? ws.alloc<UI>(nnz) : nullptr;
  if constexpr (has_val<V>) {
    V *v_tmp = P > 1 ? ws.alloc<V>(nnz) : nullptr;
    radix_sort<UI, UI, V>(ws, {(UI *)col, (UI *)row, (V *)vals}, {k_out, (UI *)out_row, out_vals},
                          {k_tmp, r_tmp, v_tmp}, nnz, ranges);
  } else {
    radix_sort<UI, UI, NoVal>(ws, {(UI *)col, (UI *)row, nullptr},
                              {k_out, (UI *)out_row, nullptr}, {k_tmp, r_tmp, nullptr}, nnz,
                              ranges);
  }
  // col_ptr from the sorted column keys; the secondary stream (rows) gives the CSC-ctor check.
  // When the row stream was non-decreasing (row ids expanded from a row_ptr), a stable sort by
  // column leaves the rows of every column ascending by construction: nothing to check.
  build_ptr_and_copy<I, N, NoVal>(ws, (const I *)k_out,
                                  rows_ascending ? (const I *)nullptr : (const I *)out_row,
                                  nullptr, nnz, n, out_col_ptr, nullptr, nullptr, h_flags);
  if (h_flags[1]) {  // csc.cc:99-157: some column has unsorted rows -> sort every column
    if constexpr (has_val<V>)
      compressed_sort_inplace<I, N, V>(ws, out_col_ptr, out_row, out_vals, n, n, nnz, vkind);
    else
      compressed_sort_inplace<I, N, NoVal>(ws, out_col_ptr, out_row, nullptr, n, n, nnz, vkind);
  }
}

}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_coo_sort(int device, int64_t n, int64_t m, int64_t nnz, void *row, void *col,
                   void *vals, int id_type, int val_type, int *h_was_sorted, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && m >= 0 && nnz >= 0, SB200_ERR_BAD_ARG, "negative size");
    SB_REQUIRE(nnz == 0 || (row && col), SB200_ERR_BAD_ARG, "row/col is null");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, SB200_I64, val_type, hv, [&](auto I_, auto, auto V_) {
      using I = decltype(I_);
      using V = decltype(V_);
      bool sorted = coo_sort_impl<I, V>(ws, n, m, nnz, (I *)row, (I *)col, (V *)vals);
      if (h_was_sorted) *h_was_sorted = sorted ? 1 : 0;
    });
  });
}

int sb200_compressed_sort(int device, int64_t n_seg, int64_t n_idx, int64_t nnz, const void *ptr,
                          void *idx, void *vals, int id_type, int nnz_type, int val_type,
                          int *h_was_sorted, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n_seg >= 0 && nnz >= 0 && ptr, SB200_ERR_BAD_ARG, "bad size or null ptr");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      bool sorted = compressed_check<I, N, V>(ws, (const N *)ptr, (const I *)idx, n_seg);
      if (h_was_sorted) *h_was_sorted = sorted ? 1 : 0;
      if (!sorted)
        compressed_sort_inplace<I, N, V>(ws, (const N *)ptr, (I *)idx, (V *)vals, n_seg, n_idx,
                                         nnz, val_type);
    });
  });
}

int sb200_coo_to_csr(int device, int64_t n, int64_t m, int64_t nnz, const void *row,
                     const void *col, const void *vals, void *out_row_ptr, void *out_col,
                     void *out_vals, int id_type, int nnz_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && m >= 0 && nnz >= 0 && out_row_ptr, SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(nnz == 0 || (row && col && out_col), SB200_ERR_BAD_ARG, "null array");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      unsigned flags[2];
      build_ptr_and_copy<I, N, V>(ws, (const I *)row, (const I *)col, (const V *)vals, nnz, n,
                                  (N *)out_row_ptr, (I *)out_col, (V *)out_vals, flags);
      // CSR constructor (csr.cc:99-157).  With a row-sorted stream flags[1] is exactly its
      // check; otherwise run the check on the compressed result.
      bool need_sort = flags[1] != 0;
      if (flags[0])
        need_sort = !compressed_check<I, N, V>(ws, (const N *)out_row_ptr, (const I *)out_col, n);
      if (need_sort)
        compressed_sort_inplace<I, N, V>(ws, (const N *)out_row_ptr, (I *)out_col,
                                         (V *)out_vals, n, m, nnz, val_type);
    });
  });
}

int sb200_csr_to_coo(int device, int64_t n, int64_t m, int64_t nnz, const void *row_ptr,
                     const void *col, const void *vals, void *out_row, void *out_col,
                     void *out_vals, int id_type, int nnz_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && nnz >= 0 && row_ptr, SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(nnz == 0 || (col && out_row && out_col), SB200_ERR_BAD_ARG, "null array");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      expand_ptr<I, N, V>(ws, (const N *)row_ptr, n, nnz, (const I *)col, (const V *)vals,
                          (I *)out_row, (I *)out_col, (V *)out_vals);
      // COO constructor (coo.cc:96-157): a CSR whose rows are unsorted yields an unsorted COO
      coo_sort_impl<I, V>(ws, n, m, nnz, (I *)out_row, (I *)out_col, (V *)out_vals);
    });
  });
}

int sb200_coo_to_csc(int device, int64_t n, int64_t m, int64_t nnz, const void *row,
                     const void *col, const void *vals, void *out_col_ptr, void *out_row,
                     void *out_vals, int id_type, int nnz_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && m >= 0 && nnz >= 0 && out_col_ptr, SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(nnz == 0 || (row && col && out_row), SB200_ERR_BAD_ARG, "null array");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      to_csc_core<I, N, V>(ws, n, m, nnz, (const I *)row, (const I *)col, (const V *)vals,
                           (N *)out_col_ptr, (I *)out_row, (V *)out_vals, val_type);
    });
  });
}

int sb200_csr_to_csc(int device, int64_t n, int64_t m, int64_t nnz, const void *row_ptr,
                     const void *col, const void *vals, void *out_col_ptr, void *out_row,
                     void *out_vals, int id_type, int nnz_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && m >= 0 && nnz >= 0 && row_ptr && out_col_ptr, SB200_ERR_BAD_ARG,
               "bad argument");
    SB_REQUIRE(nnz == 0 || (col && out_row), SB200_ERR_BAD_ARG, "null array");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      // CsrCoo: only the row expansion is materialised (col/vals are read in place).  The
      // intermediate COO constructor's sort is a no-op for the result: a stable sort by
      // column of the row-major stream is independent of the column order inside a row.
      I *rows = ws.alloc<I>(nnz);
      expand_ptr<I, N, NoVal>(ws, (const N *)row_ptr, n, nnz, (const I *)nullptr,
                              (const NoVal *)nullptr, rows, (I *)nullptr, (NoVal *)nullptr);
      to_csc_core<I, N, V>(ws, n, m, nnz, rows, (const I *)col, (const V *)vals,
                           (N *)out_col_ptr, (I *)out_row, (V *)out_vals, val_type, true);
    });
  });
}

// ---- row-block variants for the multi-GPU path ----

int sb200_coo_to_csr_block(int device, int64_t row_lo, int64_t n_local, int64_t m, int64_t nnz,
                           const void *row, const void *col, const void *vals,
                           void *out_row_ptr, void *out_col, void *out_vals, int id_type,
                           int nnz_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(row_lo >= 0 && n_local >= 0 && m >= 0 && nnz >= 0 && out_row_ptr,
               SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(nnz == 0 || (row && col && out_col), SB200_ERR_BAD_ARG, "null array");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      unsigned flags[2];
      build_ptr_and_copy<I, N, V>(ws, (const I *)row, (const I *)col, (const V *)vals, nnz,
                                  n_local, (N *)out_row_ptr, (I *)out_col, (V *)out_vals, flags,
                                  (I)row_lo);
      bool need_sort = flags[1] != 0;
      if (flags[0])
        need_sort =
            !compressed_check<I, N, V>(ws, (const N *)out_row_ptr, (const I *)out_col, n_local);
      if (need_sort)
        compressed_sort_inplace<I, N, V>(ws, (const N *)out_row_ptr, (I *)out_col,
                                         (V *)out_vals, n_local, m, nnz, val_type);
    });
  });
}

int sb200_csr_to_csc_block(int device, int64_t row_lo, int64_t n_local, int64_t m, int64_t nnz,
                           const void *row_ptr, const void *col, const void *vals,
                           void *out_col_ptr, void *out_row, void *out_vals, int id_type,
                           int nnz_type, int val_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(row_lo >= 0 && n_local >= 0 && m >= 0 && nnz >= 0 && row_ptr && out_col_ptr,
               SB200_ERR_BAD_ARG, "bad argument");
    SB_REQUIRE(nnz == 0 || (col && out_row), SB200_ERR_BAD_ARG, "null array");
    Workspace ws(device, (cudaStream_t)stream);
    const bool hv = vals != nullptr && out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      I *rows = ws.alloc<I>(nnz);
      expand_ptr<I, N, NoVal>(ws, (const N *)row_ptr, n_local, nnz, (const I *)nullptr,
                              (const NoVal *)nullptr, rows, (I *)nullptr, (NoVal *)nullptr,
                              (I)row_lo);
      // col_ptr gets m+1 entries here (one per global column)
      to_csc_core<I, N, V>(ws, m, m, nnz, rows, (const I *)col, (const V *)vals,
                           (N *)out_col_ptr, (I *)out_row, (V *)out_vals, val_type, true);
    });
  });
}

}  // extern "C"
