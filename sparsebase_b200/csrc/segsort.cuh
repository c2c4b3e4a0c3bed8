// segsort.cuh -- fused gather / renumber / per-segment sort, one pass over HBM.
//
// Implements the arithmetic of the CSR / CSC constructor's per-row sort
// (format/csr.cc:123-157, format/csc.cc:123-157) and, with a gathering loader, the whole of
// PermuteOrderTwoCSR + that constructor sort (permute/permute_order_two.cc:64-77): each CTA
// owns the segments that START inside one 2048-entry window of the OUTPUT layout, pulls their
// entries through the loader into shared memory (for Permute2D: old row located through the
// inverted row order, column ids renumbered through col_order), sorts every segment on chip
// and writes it to its final place.  Every nonzero is read once and written once.
//
//   segment length <= 32   : rank-by-enumeration, one thread per entry
//   33 .. kSsLong          : bitonic network run by one warp in shared memory
//   > kSsLong              : appended to a list; sorted afterwards by the global radix sort
//                            on the composite key (list rank, index)        (segsort_long)
//
// Ties (duplicate indices inside a segment) are outside the parity contract; they
// are ordered deterministically: original order (<=32) or by the value's bit pattern (>32).
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

constexpr int kSsBlock = 256;
constexpr int kSsTile = 1024;  // window of output positions per CTA
constexpr int kSsLong = 1024;  // longest segment sorted on chip (must be >= kSsTile)
constexpr int kSsCap = kSsTile + kSsLong;
constexpr int kSsEnum = 32;
constexpr int kSsMaxMid = kSsCap / (kSsEnum + 1) + 1;

// first segment r with ptr[r] >= t*kSsTile, for every window t (tile_seg[ntiles] = n_seg)
template <typename N>
__global__ void ss_tile_bounds_kernel(const N *__restrict__ ptr, int64_t n_seg, int64_t ntiles,
                                      int64_t *__restrict__ tile_seg) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  if (t == ntiles) {
    tile_seg[t] = n_seg;
    return;
  }
  const int64_t target = t * kSsTile;
  int64_t lo = 0, hi = n_seg;  // lower_bound over ptr[0..n_seg)
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)ptr[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  tile_seg[t] = lo;
}

template <typename I, typename V>
struct SsSmem {
  I key[kSsCap];
  typename std::conditional<has_val<V>, V, char>::type val[has_val<V> ? kSsCap : 1];
  unsigned lrow[kSsCap];          // segment number (relative to the window's first) per entry
  unsigned short sbeg[kSsCap];    // local start / end of the entry's segment
  unsigned short send[kSsCap];
  unsigned mid[kSsMaxMid];        // segments of 33..kSsLong entries (relative numbers)
  unsigned scratch[34];
  unsigned nmid;
};

template <typename I, typename V>
__device__ __forceinline__ bool ss_less(I ka, V va, I kb, V vb) {
  if (ka != kb) return ka < kb;
  if constexpr (has_val<V>)
    return va < vb;
  else
    return false;
}

// Loader concept:  int64_t seg_base(int64_t seg)  -- source offset of the segment's first entry
//                  I key(int64_t src_pos), V val(int64_t src_pos)
template <typename I, typename N, typename V, typename Loader>
__global__ void __launch_bounds__(kSsBlock)
    ss_tile_kernel(Loader ld, const N *__restrict__ ptr, const int64_t *__restrict__ tile_seg,
                   I *__restrict__ out_idx, V *__restrict__ out_val,
                   int64_t *__restrict__ long_list, unsigned *__restrict__ long_count) {
  extern __shared__ __align__(16) unsigned char ss_smem_raw[];
  SsSmem<I, V> &s = *reinterpret_cast<SsSmem<I, V> *>(ss_smem_raw);
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const int64_t t = blockIdx.x;
  const int64_t r0 = tile_seg[t];
  int64_t r1 = tile_seg[t + 1];
  if (r0 >= r1) return;
  // at most the LAST segment starting in this window can be long (kSsLong >= kSsTile)
  {
    const int64_t last_len = (int64_t)ptr[r1] - (int64_t)ptr[r1 - 1];
    if (last_len > kSsLong) {
      if (threadIdx.x == 0) long_list[atomicAdd(long_count, 1u)] = r1 - 1;
      r1--;
    }
  }
  if (r0 >= r1) return;
  const int64_t first = ptr[r0];
  const int count = (int)((int64_t)ptr[r1] - first);
  if (count == 0) return;

  if (threadIdx.x == 0) s.nmid = 0;
  for (int q = threadIdx.x; q < count; q += kSsBlock) s.lrow[q] = 0;
  __syncthreads();

  // ---- every non-empty segment marks its first position with (relative number + 1) ----
  for (int64_t r = r0 + threadIdx.x; r < r1; r += kSsBlock) {
    const int64_t sb = (int64_t)ptr[r] - first, se = (int64_t)ptr[r + 1] - first;
    if (se > sb) {
      s.lrow[sb] = (unsigned)(r - r0) + 1u;
      if (se - sb > kSsEnum) s.mid[atomicAdd(&s.nmid, 1u)] = (unsigned)(r - r0);
    }
  }
  __syncthreads();

  // ---- propagate segment numbers to every position: inclusive max-scan of the marks ----
  {
    constexpr int kPer = kSsCap / kSsBlock;  // consecutive positions per thread
    const int q0 = threadIdx.x * kPer;
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < kPer; k++) {
      int q = q0 + k;
      unsigned h = q < count ? s.lrow[q] : 0u;
      m = h > m ? h : m;
    }
    unsigned inc = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned tt = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)lane >= o) inc = tt > inc ? tt : inc;
    }
    if (lane == 31) s.scratch[wid] = inc;
    __syncthreads();
    unsigned carry = 0;
    for (unsigned w = 0; w < wid; w++) carry = s.scratch[w] > carry ? s.scratch[w] : carry;
    unsigned prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = 0;
    unsigned run = prev > carry ? prev : carry;
#pragma unroll
    for (int k = 0; k < kPer; k++) {
      int q = q0 + k;
      if (q < count) {
        unsigned h = s.lrow[q];
        run = h > run ? h : run;
        s.lrow[q] = run - 1u;
      }
    }
  }
  __syncthreads();

  // ---- gather the entries through the loader (coalesced in the output layout); loads are
  //      batched 8 deep per thread so that the dependent gathers overlap ----
  for (int qb = 0; qb < count; qb += kSsBlock * 8) {
    I k[8];
    [[maybe_unused]] typename std::conditional<has_val<V>, V, char>::type v[8];
    int sb[8], se[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int q = qb + u * kSsBlock + (int)threadIdx.x;
      if (q < count) {
        const int64_t r = r0 + s.lrow[q];
        sb[u] = (int)((int64_t)ptr[r] - first);
        se[u] = (int)((int64_t)ptr[r + 1] - first);
        const int64_t p = ld.seg_base(r) + (q - sb[u]);
        k[u] = ld.key(p);
        if constexpr (has_val<V>) v[u] = ld.val(p);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int q = qb + u * kSsBlock + (int)threadIdx.x;
      if (q < count) {
        s.key[q] = k[u];
        if constexpr (has_val<V>) s.val[q] = v[u];
        s.sbeg[q] = (unsigned short)sb[u];
        s.send[q] = (unsigned short)se[u];
      }
    }
  }
  __syncthreads();

  // ---- short segments: rank by enumeration, write straight to the final position ----
  for (int q = threadIdx.x; q < count; q += kSsBlock) {
    const int sb = s.sbeg[q], se = s.send[q];
    if (se - sb > kSsEnum) continue;
    const I k = s.key[q];
    int rank = 0;
    for (int j = sb; j < se; j++) {
      const I kj = s.key[j];
      rank += (kj < k || (kj == k && j < q)) ? 1 : 0;
    }
    out_idx[first + sb + rank] = k;
    if constexpr (has_val<V>) out_val[first + sb + rank] = s.val[q];
  }

  // ---- mid segments: one warp each, normalized bitonic network in shared memory ----
  const unsigned nmid = s.nmid;
  for (unsigned mi = wid; mi < nmid; mi += kSsBlock / 32) {
    const int64_t r = r0 + s.mid[mi];
    const int sb = (int)((int64_t)ptr[r] - first), len = (int)((int64_t)ptr[r + 1] - ptr[r]);
    I *key = s.key + sb;
    int P = 64;
    while (P < len) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        const bool flip = (j == (k >> 1));
        for (int x = lane; x < (P >> 1); x += 32) {
          // x-th comparator of this step: low index a has bit j clear
          const int a = ((x & ~(j - 1)) << 1) | (x & (j - 1));
          const int b = flip ? (a ^ (k - 1)) : (a | j);
          if (b < len) {
            const I ka = key[a], kb = key[b];
            if constexpr (has_val<V>) {
              V *val = reinterpret_cast<V *>(s.val) + sb;
              const V va = val[a], vb = val[b];
              if (ss_less<I, V>(kb, vb, ka, va)) {
                key[a] = kb;
                key[b] = ka;
                val[a] = vb;
                val[b] = va;
              }
            } else {
              if (kb < ka) {
                key[a] = kb;
                key[b] = ka;
              }
            }
          }
        }
        __syncwarp();
      }
    }
    for (int x = lane; x < len; x += 32) {
      out_idx[first + sb + x] = key[x];
      if constexpr (has_val<V>) out_val[first + sb + x] = reinterpret_cast<V *>(s.val)[sb + x];
    }
  }
}

// ------------------------------------------------------------------ long segments
template <typename N>
struct LongLenFn {
  const int64_t *list;
  const N *ptr;
  __device__ int64_t operator()(int64_t k) const {
    int64_t r = list[k];
    return (int64_t)ptr[r + 1] - (int64_t)ptr[r];
  }
};

template <typename I, typename N, typename V, typename Loader>
__global__ void ss_long_fill_kernel(Loader ld, const N *__restrict__ ptr,
                                    const int64_t *__restrict__ list,
                                    const int64_t *__restrict__ offs, int64_t nlong,
                                    int64_t total, int idx_bits, uint64_t *__restrict__ keys,
                                    V *__restrict__ vals) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = nlong;  // last k with offs[k] <= e
    while (hi - lo > 1) {
      int64_t mid = (lo + hi) >> 1;
      if (offs[mid] <= e)
        lo = mid;
      else
        hi = mid;
    }
    const int64_t r = list[lo];
    const int64_t p = ld.seg_base(r) + (e - offs[lo]);
    keys[e] = ((uint64_t)lo << idx_bits) | (uint64_t)ld.key(p);
    if constexpr (has_val<V>) vals[e] = ld.val(p);
  }
}

template <typename I, typename N, typename V>
__global__ void ss_long_store_kernel(const N *__restrict__ ptr, const int64_t *__restrict__ list,
                                     const int64_t *__restrict__ offs, int64_t total,
                                     int idx_bits, const uint64_t *__restrict__ keys,
                                     const V *__restrict__ vals, I *__restrict__ out_idx,
                                     V *__restrict__ out_val) {
  const uint64_t mask = (1ull << idx_bits) - 1ull;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[e];
    const int64_t lo = (int64_t)(k >> idx_bits);
    const int64_t dst = (int64_t)ptr[list[lo]] + (e - offs[lo]);
    out_idx[dst] = (I)(k & mask);
    if constexpr (has_val<V>) out_val[dst] = vals[e];
  }
}

// Sorts every segment of the output layout `ptr` (n_seg+1 offsets, nnz entries), pulling the
// entries through `ld`.  n_idx bounds the index values (for the long-segment key width).
// Synchronises the stream once (to learn how many long segments there are).
template <typename I, typename N, typename V, typename Loader>
void segmented_sort(Workspace &ws, Loader ld, const N *ptr, int64_t n_seg, int64_t n_idx,
                    int64_t nnz, I *out_idx, V *out_val) {
  if (nnz <= 0 || n_seg <= 0) return;
  cudaStream_t st = ws.stream();
  const int64_t ntiles = ceil_div(nnz, kSsTile);
  int64_t *tile_seg = ws.alloc<int64_t>(ntiles + 1);
  // every long segment is the last one of a distinct window
  int64_t *long_list = ws.alloc<int64_t>(ntiles + 1);
  unsigned *long_count = ws.alloc<unsigned>(1);
  SB_CUDA(cudaMemsetAsync(long_count, 0, sizeof(unsigned), st));
  SB_LAUNCH((ss_tile_bounds_kernel<N>), (unsigned)ceil_div(ntiles + 1, 256), 256, 0, st, ptr,
            n_seg, ntiles, tile_seg);
  auto kern = ss_tile_kernel<I, N, V, Loader>;
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(SsSmem<I, V>)));
  SB_LAUNCH(kern, (unsigned)ntiles, kSsBlock, sizeof(SsSmem<I, V>), st, ld, ptr, tile_seg,
            out_idx, out_val, long_list, long_count);
  unsigned nlong = 0;
  SB_CUDA(cudaMemcpyAsync(&nlong, long_count, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (nlong == 0) return;

  // ---- long segments: composite-key global radix sort ----
  int64_t *offs = ws.alloc<int64_t>((int64_t)nlong + 1);
  exclusive_scan<int64_t>(ws, LongLenFn<N>{long_list, ptr}, offs, (int64_t)nlong);
  int64_t total = 0;
  SB_CUDA(cudaMemcpyAsync(&total, offs + nlong, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const int idx_bits = bits_for((uint64_t)(n_idx > 0 ? n_idx - 1 : 0));
  const int rank_bits = bits_for((uint64_t)nlong - 1);
  SB_REQUIRE(idx_bits + rank_bits <= 64, SB200_ERR_BAD_ARG,
             "long-segment key does not fit 64 bits (%d + %d)", idx_bits, rank_bits);
  uint64_t *kin = ws.alloc<uint64_t>(total), *ka = ws.alloc<uint64_t>(total),
           *kb = ws.alloc<uint64_t>(total);
  V *vin = nullptr, *va = nullptr, *vb = nullptr;
  if constexpr (has_val<V>) {
    vin = ws.alloc<V>(total);
    va = ws.alloc<V>(total);
    vb = ws.alloc<V>(total);
  }
  const int grid = device_info(ws.device()).sm_count * 8;
  SB_LAUNCH((ss_long_fill_kernel<I, N, V, Loader>), grid, 256, 0, st, ld, ptr, long_list, offs,
            (int64_t)nlong, total, idx_bits, kin, vin);
  std::vector<RsBitRange> ranges = {{0, idx_bits}, {idx_bits, idx_bits + rank_bits}};
  radix_sort<uint64_t, V, NoVal>(ws, {kin, vin, nullptr}, {ka, va, nullptr}, {kb, vb, nullptr},
                                 total, ranges);
  SB_LAUNCH((ss_long_store_kernel<I, N, V>), grid, 256, 0, st, ptr, long_list, offs, total,
            idx_bits, (const uint64_t *)ka, (const V *)va, out_idx, out_val);
}

}  // namespace sb200
