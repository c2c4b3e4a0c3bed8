// segsort.cuh -- fused gather / renumber / per-segment sort, one pass over HBM.
//
// Implements the arithmetic of the CSR / CSC constructor's per-row sort
// (format/csr.cc:123-157, format/csc.cc:123-157) and, with a gathering loader, the whole of
// PermuteOrderTwoCSR + that constructor sort (permute/permute_order_two.cc:64-77): each CTA
// owns the segments that START inside one 1024-entry window of the OUTPUT layout, pulls their
// entries through the loader into shared memory (for Permute2D: old row located through the
// inverted row order, column ids renumbered through col_order), sorts every segment on chip
// and writes it to its final place.  Every nonzero is read once and written once.
//
//   segment length <= 32   : rank-by-enumeration, one thread per entry
//   33 .. kSsWarpMax       : bitonic network run by one warp in shared memory
//   .. kSsLong             : the same network run by the whole CTA, one segment after the other
//                            (one warp alone on a 1000-entry segment kept the other seven
//                            waiting at the end of the CTA: 48 % of the kernel's stall samples)
//   > kSsLong              : appended to a list; sorted afterwards by a segmented radix sort
//                            on the index (radix_sort_segmented, see "long segments")
//
// Ties (duplicate indices inside a segment): the kernels here order them arbitrarily, flag the
// segment (DupCtx, common.cuh) and report whether any segment was unsorted in source order;
// dup_fix_kernel then applies the reference's rule -- (index, value) order in every segment when
// any segment was unsorted (std::less<pair<IDType, ValueType>>, csr.cc:147, values compared in
// their real type), source order otherwise (csr.cc:99-118: nothing is sorted then).
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
constexpr int kSsWarpMax = 128;  // longest segment sorted by a single warp
constexpr int kSsMaxMid = kSsCap / (kSsEnum + 1) + 1;
constexpr int kSsScanPer = kSsCap / kSsBlock;  // consecutive positions per thread in the scan
static_assert(kSsScanPer == 8, "the mark scan moves 8 u16 marks per thread as one 16-byte word");

// Per window t: the first segment r with ptr[r] >= t*kSsTile, its start ptr[r], and the start
// of the segment before it (= the last segment that starts inside window t-1), so that a CTA
// learns everything about its window from two adjacent 32-byte records (one dependent load).
struct SsTileRec {
  int64_t seg;
  int64_t first;
  int64_t prev_start;
  int64_t pad;
};

template <typename N>
__global__ void ss_tile_bounds_kernel(const N *__restrict__ ptr, int64_t n_seg, int64_t ntiles,
                                      SsTileRec *__restrict__ rec) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  int64_t lo = n_seg;
  if (t < ntiles) {
    const int64_t target = t * kSsTile;
    int64_t hi = n_seg;  // lower_bound over ptr[0..n_seg)
    lo = 0;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if ((int64_t)ptr[mid] < target)
        lo = mid + 1;
      else
        hi = mid;
    }
  }
  SsTileRec r;
  r.seg = lo;
  r.first = (int64_t)ptr[lo];
  r.prev_start = lo > 0 ? (int64_t)ptr[lo - 1] : 0;
  r.pad = 0;
  rec[t] = r;
}

template <typename I, typename N, typename V>
struct SsSmem {
  I key[kSsCap];
  typename std::conditional<has_val<V>, V, char>::type val[has_val<V> ? kSsCap : 1];
  // mark[q]: during the row pass, (q + 1) at the first position of every non-empty segment and
  // 0 elsewhere; after the scan, (start of q's segment + 1) at every position
  // (mark and base are dead once the gather and the short segments are done: the CTA-wide sort
  // uses the two arrays, contiguous on purpose, as its exchange area)
  __align__(16) unsigned short mark[kSsCap];
  N base[kSsCap];              // at a segment's first position: source offset of its entries
  unsigned short len[kSsCap];  // at a segment's first position: its length
  unsigned mid[kSsMaxMid];     // first positions of the segments with 33..kSsLong entries
  unsigned scratch[kSsBlock / 32];
  unsigned nmid;
};

// segment that owns output position `pos` (the last r with ptr[r] <= pos; rare path only)
template <typename N>
__device__ int64_t ss_segment_of(const N *__restrict__ ptr, int64_t n_seg, int64_t pos) {
  int64_t lo = 0, hi = n_seg;  // first r with ptr[r] > pos
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)ptr[mid] <= pos)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo - 1;
}

// Bitonic sorting network over P = T * E (key, value) pairs held in registers: thread t of the
// group (T = 32: one warp, T = kSsBlock: the CTA) owns the E consecutive elements t * E + i.
// Comparators between elements of one thread are register moves, between lanes of a warp two
// shuffles per element, and only the strides that cross warps (CTA version: 6 of the 55 steps
// at P = 1024) go through shared memory with barriers.  The network this replaced ran every
// step in shared memory with a barrier (57 warp-instructions per entry of the tile kernel on
// R-MAT-24, three quarters of them in the compare-exchange loop).
// Keys are compared as unsigned so that the all-ones padding sorts last; equal keys never swap.
template <typename U>
__device__ __forceinline__ U ss_min(U a, U b) {
  return a < b ? a : b;
}
template <typename U>
__device__ __forceinline__ U ss_max(U a, U b) {
  return a < b ? b : a;
}
template <int E, int T, typename UI, typename VV, bool HV>
__device__ __forceinline__ void ss_bitonic_regs(UI (&key)[E], VV (&val)[E], int t, UI *xk,
                                                VV *xv) {
  constexpr int P = T * E;
#pragma unroll 1
  for (int k = 2; k <= P; k <<= 1) {
    const bool asc_t = ((t * E) & k) == 0;  // direction of this thread's elements when k >= 2E
    int j = k >> 1;
    if constexpr (T > 32) {
#pragma unroll 1
      for (; j >= 32 * E; j >>= 1) {  // partner in another warp
        const int tm = j / E;
        const int tp = t ^ tm;
        __syncthreads();  // (the readers of the previous exchange are done)
#pragma unroll
        for (int i = 0; i < E; i++) {
          xk[t * E + i] = key[i];
          if constexpr (HV) xv[t * E + i] = val[i];
        }
        __syncthreads();
        const UI m_min = (((t & tm) == 0) == asc_t) ? ~(UI)0 : (UI)0;  // all ones: I keep the min
#pragma unroll
        for (int i = 0; i < E; i++) {
          const UI pk = xk[tp * E + i];
          // (min / max + a bit mask, not comparisons under a condition: nvcc turned those into
          // divergent branches around every shuffle)
          const UI nk = (ss_min(key[i], pk) & m_min) | (ss_max(key[i], pk) & ~m_min);
          if constexpr (HV) {
            const VV pv = xv[tp * E + i];
            val[i] = nk != key[i] ? pv : val[i];
          }
          key[i] = nk;
        }
      }
    }
#pragma unroll 1
    for (; j >= E; j >>= 1) {  // partner in another lane of the warp
      const int lm = j / E;
      const UI m_min = (((t & lm) == 0) == asc_t) ? ~(UI)0 : (UI)0;
#pragma unroll
      for (int i = 0; i < E; i++) {
        const UI pk = __shfl_xor_sync(0xffffffffu, key[i], lm);
        [[maybe_unused]] VV pv = val[i];
        if constexpr (HV) pv = __shfl_xor_sync(0xffffffffu, val[i], lm);
        const UI nk = (ss_min(key[i], pk) & m_min) | (ss_max(key[i], pk) & ~m_min);
        if constexpr (HV) val[i] = nk != key[i] ? pv : val[i];
        key[i] = nk;
      }
    }
#pragma unroll
    for (int jj = E / 2; jj >= 1; jj >>= 1) {  // partner in my own registers
      if (jj < k) {
#pragma unroll
        for (int i = 0; i < E; i++) {
          const int p2 = i ^ jj;
          if (p2 > i) {
            const bool asc = ((t * E + i) & k) == 0;
            const UI ka = key[i], kb = key[p2];
            const UI m_asc = asc ? ~(UI)0 : (UI)0;
            const UI lo = ss_min(ka, kb), hi = ss_max(ka, kb);
            key[i] = (lo & m_asc) | (hi & ~m_asc);
            key[p2] = (hi & m_asc) | (lo & ~m_asc);
            if constexpr (HV) {
              const bool sw = key[i] != ka;
              const VV va = val[i], vb = val[p2];
              val[i] = sw ? vb : va;
              val[p2] = sw ? va : vb;
            }
          }
        }
      }
    }
  }
}

// Sorts the `len` entries at skey / sval (shared memory) with the network above: loads them
// into registers (padding up to T * E), sorts, writes them back.  The caller synchronises.
template <int E, int T, typename I, typename VV, bool HV>
__device__ __forceinline__ void ss_sort_in_regs(I *skey, VV *sval, int len, int t, void *xch) {
  using UI = typename std::make_unsigned<I>::type;
  UI key[E];
  VV val[E];
#pragma unroll
  for (int i = 0; i < E; i++) {
    const int idx = t * E + i;
    key[i] = idx < len ? (UI)skey[idx] : ~(UI)0;
    if constexpr (HV) {
      val[i] = idx < len ? sval[idx] : VV();
    } else {
      val[i] = VV();
    }
  }
  UI *xk = reinterpret_cast<UI *>(xch);
  VV *xv = reinterpret_cast<VV *>(reinterpret_cast<unsigned char *>(xch) + (size_t)T * E * sizeof(UI));
  ss_bitonic_regs<E, T, UI, VV, HV>(key, val, t, xk, xv);
#pragma unroll
  for (int i = 0; i < E; i++) {
    const int idx = t * E + i;
    if (idx < len) {
      skey[idx] = (I)key[i];
      if constexpr (HV) sval[idx] = val[i];
    }
  }
}

// Loader concept:  int64_t seg_base(int64_t seg)  source offset of the segment's first entry
//                  I raw_key(int64_t src_pos)      the stored index
//                  I map_key(I raw)                renumbering applied to it (identity if none)
//                  V val(int64_t src_pos)
//
// Dependent global loads per CTA: window record -> {ptr[r], seg_base(r)} -> raw_key/val ->
// map_key.  Everything else (which segment a position belongs to, where its source is) is
// resolved in shared memory: segments mark their first position, an inclusive max-scan spreads
// the mark, and the per-segment source offset / length sit in tables indexed by that position.
template <typename I, typename N, typename V, typename Loader>
__global__ void __launch_bounds__(kSsBlock)
    ss_tile_kernel(Loader ld, const N *__restrict__ ptr, int64_t n_seg,
                   const SsTileRec *__restrict__ rec, I *__restrict__ out_idx,
                   V *__restrict__ out_val, int64_t *__restrict__ long_list,
                   unsigned *__restrict__ long_count, DupCtx dc) {
  extern __shared__ __align__(16) unsigned char ss_smem_raw[];
  SsSmem<I, N, V> &s = *reinterpret_cast<SsSmem<I, N, V> *>(ss_smem_raw);
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const int64_t t = blockIdx.x;
  const SsTileRec a = rec[t], b = rec[t + 1];
  // zero the marks while the records are in flight
  {
    uint4 *m4 = reinterpret_cast<uint4 *>(s.mark);
    m4[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) s.nmid = 0;
  }
  const int64_t r0 = a.seg;
  int64_t r1 = b.seg;
  if (r0 >= r1) return;
  const int64_t first = a.first;
  int64_t last = b.first;
  // at most the LAST segment starting in this window can be long (kSsLong >= kSsTile)
  if (last - b.prev_start > kSsLong) {
    if (threadIdx.x == 0) long_list[atomicAdd(long_count, 1u)] = r1 - 1;
    r1--;
    last = b.prev_start;
  }
  const int count = (int)(last - first);
  if (count <= 0) return;
  __syncthreads();

  // ---- row pass: every non-empty segment records itself at its first position ----
  for (int64_t r = r0 + threadIdx.x; r < r1; r += kSsBlock) {
    const int sb = (int)((int64_t)ptr[r] - first), se = (int)((int64_t)ptr[r + 1] - first);
    if (se > sb) {
      s.mark[sb] = (unsigned short)(sb + 1);
      s.len[sb] = (unsigned short)(se - sb);
      s.base[sb] = (N)ld.seg_base(r);
      if (se - sb > kSsEnum) s.mid[atomicAdd(&s.nmid, 1u)] = (unsigned)sb;
    }
  }
  __syncthreads();

  // ---- inclusive max-scan of the marks: every position learns its segment's start ----
  {
    uint4 *m4 = reinterpret_cast<uint4 *>(s.mark);
    uint4 w = m4[threadIdx.x];
    unsigned h[8] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16,
                     w.z & 0xffffu, w.z >> 16, w.w & 0xffffu, w.w >> 16};
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) m = h[k] > m ? h[k] : m;
    unsigned inc = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned tt = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int)lane >= o) inc = tt > inc ? tt : inc;
    }
    if (lane == 31) s.scratch[wid] = inc;
    __syncthreads();
    unsigned run = 0;
    for (unsigned w2 = 0; w2 < wid; w2++) run = s.scratch[w2] > run ? s.scratch[w2] : run;
    const unsigned prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane > 0) run = prev > run ? prev : run;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      run = h[k] > run ? h[k] : run;
      h[k] = run;
    }
    w.x = h[0] | (h[1] << 16);
    w.y = h[2] | (h[3] << 16);
    w.z = h[4] | (h[5] << 16);
    w.w = h[6] | (h[7] << 16);
    m4[threadIdx.x] = w;
  }
  __syncthreads();

  // ---- gather through the loader, coalesced in the output layout; the loads of a batch are
  //      issued together so that the dependent map_key gathers overlap ----
  constexpr int kBatch = 4;
  for (int qb = 0; qb < count; qb += kSsBlock * kBatch) {
    I raw[kBatch];
    [[maybe_unused]] typename std::conditional<has_val<V>, V, char>::type v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int q = qb + u * kSsBlock + (int)threadIdx.x;
      if (q < count) {
        const int sb = (int)s.mark[q] - 1;
        const int64_t p = (int64_t)s.base[sb] + (q - sb);
        raw[u] = ld.raw_key(p);
        if constexpr (has_val<V>) v[u] = ld.val(p);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int q = qb + u * kSsBlock + (int)threadIdx.x;
      if (q < count) {
        s.key[q] = ld.map_key(raw[u]);
        if constexpr (has_val<V>) s.val[q] = v[u];
      }
    }
  }
  __syncthreads();

  // ---- short segments: rank by enumeration, write straight to the final position ----
  bool unsorted = false;  // some segment has an inversion in source order (csr.cc:99-116)
  for (int q = threadIdx.x; q < count; q += kSsBlock) {
    const int sb = (int)s.mark[q] - 1;
    const int len = s.len[sb];
    if (len > kSsEnum) continue;  // (their warp checks and sorts them below)
    const I k = s.key[q];
    if (q > sb && s.key[q - 1] > k) unsorted = true;
    int rank = 0, same = 0;  // entries before q count when <=, entries after q when < (stable)
#pragma unroll 4
    for (int j = sb; j < q; j++) {
      rank += s.key[j] <= k ? 1 : 0;
      same += s.key[j] == k ? 1 : 0;
    }
#pragma unroll 4
    for (int j = q + 1; j < sb + len; j++) rank += s.key[j] < k ? 1 : 0;
    if (same == 1 && dc.seg_flag) dc.flag(ss_segment_of<N>(ptr, n_seg, first + sb));
    // outputs are written once and never read here: streaming stores keep them from pushing
    // the renumbering table out of L2
    st_stream(out_idx + first + sb + rank, k);
    if constexpr (has_val<V>) st_stream(out_val + first + sb + rank, (V)s.val[q]);
  }

  // ---- mid segments: normalized bitonic network in shared memory; comparator x of a step
  //      works on (a2, b2), the low index a2 has bit j clear ----
  const unsigned nmid = s.nmid;
  auto compare_exchange = [&](I *key, int sb, int len, int k, int j, int x) {
    const bool flip = (j == (k >> 1));
    const int a2 = ((x & ~(j - 1)) << 1) | (x & (j - 1));
    const int b2 = flip ? (a2 ^ (k - 1)) : (a2 | j);
    if (b2 < len) {
      const I ka = key[a2], kb = key[b2];
      if (kb < ka) {
        key[a2] = kb;
        key[b2] = ka;
        if constexpr (has_val<V>) {
          V *val = reinterpret_cast<V *>(s.val) + sb;
          const V va = val[a2];
          val[a2] = val[b2];
          val[b2] = va;
        }
      }
    }
  };
  // up to kSsWarpMax entries: one warp per segment
  for (unsigned mi = wid; mi < nmid; mi += kSsBlock / 32) {
    const int sb = (int)s.mid[mi], len = s.len[sb];
    if (len > kSsWarpMax) continue;
    I *key = s.key + sb;
    for (int x = lane + 1; x < len; x += 32)
      if (key[x - 1] > key[x]) unsorted = true;
    __syncwarp();
    {
      using VV = typename std::conditional<has_val<V>, V, char>::type;
      VV *sval = reinterpret_cast<VV *>(s.val) + (has_val<V> ? sb : 0);
      if (len <= 64)
        ss_sort_in_regs<2, 32, I, VV, has_val<V>>(key, sval, len, (int)lane, nullptr);
      else
        ss_sort_in_regs<4, 32, I, VV, has_val<V>>(key, sval, len, (int)lane, nullptr);
    }
    __syncwarp();
    bool dup = false;
    for (int x = lane; x < len; x += 32) {
      if (x + 1 < len && key[x] == key[x + 1]) dup = true;
      st_stream(out_idx + first + sb + x, key[x]);
      if constexpr (has_val<V>) st_stream(out_val + first + sb + x, reinterpret_cast<V *>(s.val)[sb + x]);
    }
    if (dc.seg_flag && __any_sync(0xffffffffu, dup) && lane == 0)
      dc.flag(ss_segment_of<N>(ptr, n_seg, first + sb));
  }
  // longer ones: the whole CTA on one segment at a time (uniform control flow: the list and the
  // lengths are in shared memory)
  for (unsigned mi = 0; mi < nmid; mi++) {
    const int sb = (int)s.mid[mi], len = s.len[sb];
    if (len <= kSsWarpMax) continue;
    I *key = s.key + sb;
    for (int x = threadIdx.x + 1; x < len; x += kSsBlock)
      if (key[x - 1] > key[x]) unsorted = true;
    __syncthreads();
    using VV = typename std::conditional<has_val<V>, V, char>::type;
    constexpr size_t kXchBytes = sizeof(s.mark) + sizeof(s.base);
    constexpr bool kRegsFit =
        (size_t)kSsLong * (sizeof(I) + (has_val<V> ? sizeof(V) : 0)) <= kXchBytes;
    if constexpr (kRegsFit) {
      VV *sval = reinterpret_cast<VV *>(s.val) + (has_val<V> ? sb : 0);
      void *xch = s.mark;  // mark + base: dead since the barrier above
      if (len <= 256)
        ss_sort_in_regs<1, kSsBlock, I, VV, has_val<V>>(key, sval, len, (int)threadIdx.x, xch);
      else if (len <= 512)
        ss_sort_in_regs<2, kSsBlock, I, VV, has_val<V>>(key, sval, len, (int)threadIdx.x, xch);
      else
        ss_sort_in_regs<4, kSsBlock, I, VV, has_val<V>>(key, sval, len, (int)threadIdx.x, xch);
      __syncthreads();
    } else {  // (64-bit ids with 32-bit offsets: the exchange area is too small)
      int P = 256;
      while (P < len) P <<= 1;
      for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int x = threadIdx.x; x < (P >> 1); x += kSsBlock)
            compare_exchange(key, sb, len, k, j, x);
          __syncthreads();
        }
      }
    }
    bool dup = false;
    for (int x = threadIdx.x; x < len; x += kSsBlock) {
      if (x + 1 < len && key[x] == key[x + 1]) dup = true;
      st_stream(out_idx + first + sb + x, key[x]);
      if constexpr (has_val<V>) st_stream(out_val + first + sb + x, reinterpret_cast<V *>(s.val)[sb + x]);
    }
    if (dc.seg_flag && __syncthreads_or(dup ? 1 : 0) && threadIdx.x == 0)
      dc.flag(ss_segment_of<N>(ptr, n_seg, first + sb));
  }
  dc.report_unsorted(unsorted);
}

// ------------------------------------------------------------------ duplicate ids (rare)
// Runs after the fast kernels when some segment was flagged (see DupCtx).  One warp per flagged
// segment of the OUTPUT layout:
//   the sort happens (some segment was unsorted in source order, or the caller checked before):
//     inside every run of equal ids the values are put in ascending order of their real type
//   nothing was unsorted: the reference leaves every segment as it was -> the segment is copied
//     again from the source, in source order
template <typename I, typename N, typename V, typename Loader>
__global__ void __launch_bounds__(256)
    dup_fix_kernel(Loader ld, const N *__restrict__ ptr, int64_t n_seg, DupCtx dc, int sort_known,
                   I *__restrict__ out_idx, V *__restrict__ out_val, int vkind) {
  if (*reinterpret_cast<volatile unsigned *>(dc.any_dup) == 0u) return;
  const bool sorted =
      sort_known || (dc.unsorted && *reinterpret_cast<volatile unsigned *>(dc.unsorted) != 0u);
  const unsigned lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp * 32; r0 < n_seg; r0 += nwarps * 32) {
    // 32 flags per warp step, then the flagged segments one at a time
    const unsigned flagged =
        __ballot_sync(0xffffffffu, r0 + lane < n_seg && dc.seg_flag[r0 + lane] != 0);
    for (unsigned rest = flagged; rest; rest &= rest - 1) {
      const int64_t r = r0 + (__ffs(rest) - 1);
      const int64_t b = (int64_t)ptr[r], e = (int64_t)ptr[r + 1];
      if (sorted) {
        for (int64_t p = b + lane; p + 1 < e; p += 32) {
          const I k = out_idx[p];
          if (out_idx[p + 1] != k || (p > b && out_idx[p - 1] == k)) continue;
          int64_t end = p + 2;  // run [p, end) of equal ids: insertion sort of its values
          while (end < e && out_idx[end] == k) end++;
          for (int64_t x = p + 1; x < end; x++) {
            const V vx = out_val[x];
            const V kx = val_order_key<V>(vx, vkind);
            int64_t y = x;
            while (y > p && val_order_key<V>(out_val[y - 1], vkind) > kx) {
              out_val[y] = out_val[y - 1];
              y--;
            }
            out_val[y] = vx;
          }
        }
      } else {
        const int64_t src = ld.seg_base(r);
        for (int64_t t = lane; t < e - b; t += 32) {
          out_idx[b + t] = ld.map_key(ld.raw_key(src + t));
          out_val[b + t] = ld.val(src + t);
        }
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------ long segments
// Segments longer than kSsLong are sorted by a SEGMENTED radix sort on the index alone
// (radix_sort_segmented): the entries of the long segments are concatenated, every segment is
// cut into tile-sized chunks, and each LSD pass ranks a record inside its own segment.  Cost:
// ceil(bits(n_idx) / 8) passes over (index, value) records -- the composite (segment rank,
// index) key of the first version needed 5-7 passes over wider records and was 55 % of
// Permute2D on R-MAT-25.
template <typename N>
struct LongLenFn {
  const int64_t *list;
  const N *ptr;
  __device__ int64_t operator()(int64_t k) const {
    int64_t r = list[k];
    return (int64_t)ptr[r + 1] - (int64_t)ptr[r];
  }
};
template <typename N>
struct LongChunksFn {  // number of tile-sized chunks of long segment k
  const int64_t *list;
  const N *ptr;
  int tile;
  __device__ int64_t operator()(int64_t k) const {
    int64_t r = list[k];
    return ((int64_t)ptr[r + 1] - (int64_t)ptr[r] + tile - 1) / tile;
  }
};

// One thread per chunk: which long segment it belongs to, and where its records / spine slots are.
static __global__ void ss_long_chunks_kernel(const int64_t *__restrict__ offs,
                                      const int64_t *__restrict__ chunk_first, int64_t nlong,
                                      int64_t nchunks, int tile, RsSeg *__restrict__ seg,
                                      int *__restrict__ chunk_seg) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  int64_t lo = 0, hi = nlong;  // last k with chunk_first[k] <= c
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (chunk_first[mid] <= c)
      lo = mid;
    else
      hi = mid;
  }
  const int64_t local = c - chunk_first[lo];
  const int64_t len = offs[lo + 1] - offs[lo];
  const int64_t rest = len - local * tile;
  RsSeg g;
  g.begin = offs[lo] + local * tile;
  g.count = (int)(rest < tile ? rest : tile);
  g.stride = (int)(chunk_first[lo + 1] - chunk_first[lo]);
  g.spine_base = chunk_first[lo] * kRsMaxBins + local;
  seg[c] = g;
  chunk_seg[c] = (int)lo;
}

// One CTA per chunk: pull the chunk's entries through the loader (coalesced inside the source
// row, renumbering gathers batched) into the concatenated arrays.
template <typename I, typename N, typename V, typename Loader>
__global__ void __launch_bounds__(256)
    ss_long_fill_kernel(Loader ld, const int64_t *__restrict__ list,
                        const int64_t *__restrict__ offs, const RsSeg *__restrict__ seg,
                        const int *__restrict__ chunk_seg,
                        typename std::make_unsigned<I>::type *__restrict__ keys,
                        V *__restrict__ vals, DupCtx dc) {
  using UI = typename std::make_unsigned<I>::type;
  const RsSeg g = seg[blockIdx.x];
  const int k = chunk_seg[blockIdx.x];
  const int64_t src0 = ld.seg_base(list[k]) + (g.begin - offs[k]);
  constexpr int kBatch = 4;
  bool unsorted = false;
  for (int q0 = 0; q0 < g.count; q0 += 256 * kBatch) {
    I raw[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int q = q0 + u * 256 + (int)threadIdx.x;
      if (q < g.count) {
        raw[u] = ld.raw_key(src0 + q);
        if constexpr (has_val<V>) vals[g.begin + q] = ld.val(src0 + q);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int q = q0 + u * 256 + (int)threadIdx.x;
      if (q < g.count) keys[g.begin + q] = (UI)ld.map_key(raw[u]);
    }
  }
  if (dc.unsorted) {  // inversions in source order (the keys just written, re-read from L2)
    __syncthreads();
    for (int q = threadIdx.x; q < g.count; q += 256) {
      const UI cur = __ldcg(keys + g.begin + q);
      UI prev = cur;
      if (q > 0)
        prev = __ldcg(keys + g.begin + q - 1);
      else if (g.begin > offs[k])  // the entry before this chunk belongs to another CTA
        prev = (UI)ld.map_key(ld.raw_key(src0 - 1));
      if (prev > cur) unsorted = true;
    }
  }
  dc.report_unsorted(unsorted);
}

template <typename I, typename N, typename V>
__global__ void __launch_bounds__(256)
    ss_long_store_kernel(const N *__restrict__ ptr, const int64_t *__restrict__ list,
                         const int64_t *__restrict__ offs, const RsSeg *__restrict__ seg,
                         const int *__restrict__ chunk_seg,
                         const typename std::make_unsigned<I>::type *__restrict__ keys,
                         const V *__restrict__ vals, I *__restrict__ out_idx,
                         V *__restrict__ out_val, DupCtx dc) {
  const RsSeg g = seg[blockIdx.x];
  const int k = chunk_seg[blockIdx.x];
  const int64_t dst0 = (int64_t)ptr[list[k]] + (g.begin - offs[k]);
  bool dup = false;
  for (int q = threadIdx.x; q < g.count; q += 256) {
    const auto key = ld_stream(keys + g.begin + q);
    if (dc.seg_flag && g.begin + q > offs[k] && keys[g.begin + q - 1] == key) dup = true;
    st_stream(out_idx + dst0 + q, (I)key);
    if constexpr (has_val<V>) st_stream(out_val + dst0 + q, ld_stream(vals + g.begin + q));
  }
  if (dc.seg_flag && __syncthreads_or(dup ? 1 : 0) && threadIdx.x == 0) dc.flag(list[k]);
}

// ------------------------------------------------------------------ big segments
// Segments longer than kSsLong that still fit in shared memory are sorted by ONE CTA entirely
// on chip: gather + renumber into shared memory, ceil(bits / 9) stable LSD passes between two
// shared buffers (every warp owns a contiguous part of the segment: per-warp digit counts, one
// scan over (digit, warp), ranks inside a batch of 32 from nine ballots -- no atomics: ATOMS
// costs 2 cycles per lane, MATCH.ANY one step per distinct digit), one coalesced write.  The
// segmented global radix sort below moves these entries through HBM once per pass in runs of a
// few entries per digit (a 2 000-entry segment has 8 entries per digit): 21 of the 54 ms of
// Permute2D on R-MAT-26.  Two shapes: 256 threads / 72 KB (three CTAs per SM: one gathers
// while the others sort) for segments up to 3 072 entries (4-byte ids and values), 512 threads
// / 200 KB up to 9 216.  A segment that does not fit is appended to `next_list`.
constexpr int kSbDigitBits = 9;
constexpr int kSbBins = 1 << kSbDigitBits;
template <int THREADS>
constexpr int ss_big_hist_bytes() {
  return (THREADS / 32) * kSbBins * (int)sizeof(unsigned);
}
template <typename I, typename V>
constexpr int ss_big_entry_bytes() {  // two (key, value) buffers + the 2-byte rank of the pass
  return 2 * ((int)sizeof(I) + (has_val<V> ? (int)sizeof(V) : 0)) + 2;
}
template <typename I, typename V, int THREADS, int BUDGET_KB>
constexpr int ss_big_cap() {
  return ((BUDGET_KB * 1024 - ss_big_hist_bytes<THREADS>()) / ss_big_entry_bytes<I, V>()) / 512 *
         512;
}
template <typename I, typename V, int THREADS, int BUDGET_KB>
constexpr int ss_big_smem() {
  return ss_big_cap<I, V, THREADS, BUDGET_KB>() * ss_big_entry_bytes<I, V>() +
         ss_big_hist_bytes<THREADS>();
}

// lanes with the same digit, for two independent batches at once (the two chains of dependent
// VOTE results interleave)
__device__ __forceinline__ void sb_match_digit2(unsigned da, unsigned db, int bits,
                                                unsigned &peers_a, unsigned &peers_b) {
  unsigned pa = 0xffffffffu, pb = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < kSbDigitBits; b++) {
    if (b < bits) {  // (warp-uniform)
      const bool bit_a = (da & (1u << b)) != 0u, bit_b = (db & (1u << b)) != 0u;
      const unsigned bal_a = __ballot_sync(0xffffffffu, bit_a);
      const unsigned bal_b = __ballot_sync(0xffffffffu, bit_b);
      pa &= bit_a ? bal_a : ~bal_a;
      pb &= bit_b ? bal_b : ~bal_b;
    }
  }
  peers_a = pa;
  peers_b = pb;
}

// grid: one CTA per entry of `list` (n_list read on the device when list_count is given: the
// second shape is launched before the host knows how many segments the first one passed on).
template <typename I, typename N, typename V, typename Loader, int THREADS, int BUDGET_KB>
__global__ void __launch_bounds__(THREADS)
    ss_big_kernel(Loader ld, const N *__restrict__ ptr, const int64_t *__restrict__ list,
                  const unsigned *__restrict__ list_count, int idx_bits, I *__restrict__ out_idx,
                  V *__restrict__ out_val, int64_t *__restrict__ next_list,
                  unsigned *__restrict__ next_count, DupCtx dc) {
  using UI = typename std::make_unsigned<I>::type;
  constexpr int CAP = ss_big_cap<I, V, THREADS, BUDGET_KB>();
  constexpr int WARPS = THREADS / 32;
  extern __shared__ __align__(16) unsigned char sb_raw[];
  __shared__ unsigned digit_base[kSbBins];
  __shared__ unsigned warp_tot[kSbBins / 32];
  if (list_count && blockIdx.x >= *list_count) return;
  UI *k0 = reinterpret_cast<UI *>(sb_raw), *k1 = k0 + CAP;
  V *v0 = nullptr, *v1 = nullptr;
  unsigned *hist;
  if constexpr (has_val<V>) {
    v0 = reinterpret_cast<V *>(k1 + CAP);
    v1 = v0 + CAP;
    hist = reinterpret_cast<unsigned *>(v1 + CAP);
  } else {
    hist = reinterpret_cast<unsigned *>(k1 + CAP);
  }
  // loc[q] = rank of entry q among the entries of its digit inside its warp's part
  unsigned short *loc = reinterpret_cast<unsigned short *>(hist + WARPS * kSbBins);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t seg = list[blockIdx.x];
  const int64_t dst0 = (int64_t)ptr[seg];
  const int64_t len64 = (int64_t)ptr[seg + 1] - dst0;
  if (len64 > CAP) {
    if (tid == 0) next_list[atomicAdd(next_count, 1u)] = seg;
    return;
  }
  const int len = (int)len64;
  const int64_t src0 = ld.seg_base(seg);
  // ---- gather (coalesced inside the source row, renumbering gathers batched)
  constexpr int kBatch = 4;
  for (int q0 = 0; q0 < len; q0 += THREADS * kBatch) {
    I raw[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int q = q0 + u * THREADS + tid;
      if (q < len) {
        raw[u] = ld.raw_key(src0 + q);
        if constexpr (has_val<V>) v0[q] = ld.val(src0 + q);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int q = q0 + u * THREADS + tid;
      if (q < len) k0[q] = (UI)ld.map_key(raw[u]);
    }
  }
  __syncthreads();
  bool unsorted = false;
  if (dc.unsorted)
    for (int q = tid + 1; q < len; q += THREADS)
      if (k0[q - 1] > k0[q]) unsorted = true;
  dc.report_unsorted(unsorted);
  // ---- LSD passes
  const int P = (idx_bits + kSbDigitBits - 1) / kSbDigitBits;
  const int per_warp = (((len + WARPS - 1) / WARPS) + 31) & ~31;
  const int wb = w * per_warp < len ? w * per_warp : len;
  const int we = wb + per_warp < len ? wb + per_warp : len;
  unsigned *my_hist = hist + w * kSbBins;
  int shift = 0;
  for (int pass = 0; pass < P; pass++) {
    const int bits = (idx_bits - shift + (P - pass) - 1) / (P - pass);
    const int nbins = 1 << bits;
    const unsigned mask = (unsigned)nbins - 1u;
    for (int k = tid; k < WARPS * kSbBins; k += THREADS) hist[k] = 0;
    __syncthreads();
    // digit counts of my part + every entry's rank inside its digit (two batches of 32 per
    // step; the lowest lane of a group of equal digits moves the counter)
    for (int base = wb; base < we; base += 64) {
      const int qa = base + lane, qb = qa + 32;
      const bool in_a = qa < we, in_b = qb < we;
      const unsigned da = in_a ? ((unsigned)(k0[qa] >> shift) & mask) : mask;
      const unsigned db = in_b ? ((unsigned)(k0[qb] >> shift) & mask) : mask;
      unsigned pa, pb;
      sb_match_digit2(da, db, bits, pa, pb);
      pa &= __ballot_sync(0xffffffffu, in_a);
      pb &= __ballot_sync(0xffffffffu, in_b);
      const unsigned lt = (1u << lane) - 1u;
      const int lead_a = __ffs(pa) - 1, lead_b = __ffs(pb) - 1;
      unsigned cur = 0;
      if (in_a && lane == lead_a) {
        cur = my_hist[da];
        my_hist[da] = cur + __popc(pa);
      }
      cur = __shfl_sync(0xffffffffu, cur, lead_a & 31);
      if (in_a) loc[qa] = (unsigned short)(cur + __popc(pa & lt));
      __syncwarp();
      cur = 0;
      if (in_b && lane == lead_b) {
        cur = my_hist[db];
        my_hist[db] = cur + __popc(pb);
      }
      cur = __shfl_sync(0xffffffffu, cur, lead_b & 31);
      if (in_b) loc[qb] = (unsigned short)(cur + __popc(pb & lt));
      __syncwarp();
    }
    __syncthreads();
    // hist[w][d] -> entries of digit d in the warps before w; digit_base[d] -> smaller digits
    for (int d0 = 0; d0 < nbins; d0 += THREADS) {  // (block-uniform trip count)
      const int d = d0 + tid;
      unsigned tot = 0;
      if (d < nbins) {
#pragma unroll
        for (int ww = 0; ww < WARPS; ww++) {
          const unsigned c = hist[ww * kSbBins + d];
          hist[ww * kSbBins + d] = tot;
          tot += c;
        }
      }
      const unsigned incl = warp_inclusive_scan(tot);
      if (lane == 31) warp_tot[(d0 >> 5) + w] = incl;
      if (d < nbins) digit_base[d] = incl - tot;  // exclusive inside its group of 32 digits
    }
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the (at most 16) group totals
      const int groups = (nbins + 31) >> 5;
      const unsigned t = tid < groups ? warp_tot[tid] : 0u;
      const unsigned incl = warp_inclusive_scan(t);
      if (tid < groups) warp_tot[tid] = incl - t;
    }
    __syncthreads();
    // (no ranking left to do: position = smaller digits + same digit in the warps before
    // mine + rank inside my part)
    for (int q = wb + lane; q < we; q += 32) {
      const UI key = k0[q];
      const unsigned d = (unsigned)(key >> shift) & mask;
      const unsigned pos = digit_base[d] + warp_tot[d >> 5] + my_hist[d] + loc[q];
      k1[pos] = key;
      if constexpr (has_val<V>) v1[pos] = v0[q];
    }
    __syncthreads();
    UI *tk = k0;
    k0 = k1;
    k1 = tk;
    if constexpr (has_val<V>) {
      V *tv = v0;
      v0 = v1;
      v1 = tv;
    }
    shift += bits;
  }
  // ---- store
  bool dup = false;
  for (int q = tid; q < len; q += THREADS) {
    const UI key = k0[q];
    if (dc.seg_flag && q > 0 && k0[q - 1] == key) dup = true;
    st_stream(out_idx + dst0 + q, (I)key);
    if constexpr (has_val<V>) st_stream(out_val + dst0 + q, v0[q]);
  }
  if (dc.seg_flag && __syncthreads_or(dup ? 1 : 0) && tid == 0) dc.flag(seg);
}

inline bool ss_big_enabled() {  // SB200_SS_BIG=0: every long segment takes the global path
  static const bool on = [] {
    const char *e = getenv("SB200_SS_BIG");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Scratch of the duplicate-id rule for n_seg segments (flags zeroed on the stream).
// `detect_unsorted` = false when the caller already knows that the sort happens.
inline DupCtx make_dup_ctx(Workspace &ws, int64_t n_seg, bool detect_unsorted) {
  cudaStream_t st = ws.stream();
  DupCtx dc;
  dc.seg_flag = ws.alloc<unsigned char>(n_seg > 0 ? n_seg : 1);
  unsigned *two = ws.alloc<unsigned>(2);
  dc.any_dup = two;
  dc.unsorted = detect_unsorted ? two + 1 : nullptr;
  SB_CUDA(cudaMemsetAsync(dc.seg_flag, 0, n_seg > 0 ? n_seg : 1, st));
  SB_CUDA(cudaMemsetAsync(two, 0, 2 * sizeof(unsigned), st));
  return dc;
}

template <typename I, typename N, typename V, typename Loader>
void launch_dup_fix(Workspace &ws, Loader ld, const N *ptr, int64_t n_seg, const DupCtx &dc,
                    bool sort_known, I *out_idx, V *out_val, int vkind) {
  if constexpr (has_val<V>) {
    if (!dc.seg_flag || n_seg <= 0) return;
    SB_LAUNCH((dup_fix_kernel<I, N, V, Loader>), device_info(ws.device()).sm_count * 2, 256, 0,
              ws.stream(), ld, ptr, n_seg, dc, sort_known ? 1 : 0, out_idx, out_val, vkind);
  }
}

// Sorts every segment of the output layout `ptr` (n_seg+1 offsets, nnz entries), pulling the
// entries through `ld`.  n_idx bounds the index values (for the long-segment key width).
// Synchronises the stream once (to learn how many long segments there are).
// sort_known: the caller established that some segment is unsorted (constructor check), so
// every segment is (index, value)-sorted; otherwise that is detected on the way (Permute2D).
template <typename I, typename N, typename V, typename Loader>
void segmented_sort(Workspace &ws, Loader ld, const N *ptr, int64_t n_seg, int64_t n_idx,
                    int64_t nnz, I *out_idx, V *out_val, int vkind, bool sort_known) {
  if (nnz <= 0 || n_seg <= 0) return;
  cudaStream_t st = ws.stream();
  DupCtx dc = {nullptr, nullptr, nullptr};
  if constexpr (has_val<V>) dc = make_dup_ctx(ws, n_seg, !sort_known);
  const int64_t ntiles = ceil_div(nnz, kSsTile);
  SsTileRec *tile_rec = ws.alloc<SsTileRec>(ntiles + 1);
  // every long segment is the last one of a distinct window
  int64_t *long_list = ws.alloc<int64_t>(ntiles + 1);
  unsigned *long_count = ws.alloc<unsigned>(1);
  SB_CUDA(cudaMemsetAsync(long_count, 0, sizeof(unsigned), st));
  SB_LAUNCH((ss_tile_bounds_kernel<N>), (unsigned)ceil_div(ntiles + 1, 256), 256, 0, st, ptr,
            n_seg, ntiles, tile_rec);
  auto kern = ss_tile_kernel<I, N, V, Loader>;
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(SsSmem<I, N, V>)));
  SB_LAUNCH(kern, (unsigned)ntiles, kSsBlock, sizeof(SsSmem<I, N, V>), st, ld, ptr, n_seg,
            (const SsTileRec *)tile_rec, out_idx, out_val, long_list, long_count, dc);
  unsigned nlong = 0;
  SB_CUDA(cudaMemcpyAsync(&nlong, long_count, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (nlong == 0) {
    launch_dup_fix<I, N, V, Loader>(ws, ld, ptr, n_seg, dc, sort_known, out_idx, out_val, vkind);
    return;
  }

  using UI = typename std::make_unsigned<I>::type;
  const int idx_bits = bits_for((uint64_t)(n_idx > 0 ? n_idx - 1 : 0));
  // ---- big segments: one CTA each, on chip; what does not fit is listed for the global path
  if (ss_big_enabled()) {
    int64_t *mid_list = ws.alloc<int64_t>((int64_t)nlong);
    int64_t *huge_list = ws.alloc<int64_t>((int64_t)nlong);
    unsigned *counts = ws.alloc<unsigned>(2);
    SB_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned), st));
    auto k_small = ss_big_kernel<I, N, V, Loader, 256, 72>;
    auto k_large = ss_big_kernel<I, N, V, Loader, 512, 200>;
    constexpr int smem_small = ss_big_smem<I, V, 256, 72>();
    constexpr int smem_large = ss_big_smem<I, V, 512, 200>();
    SB_CUDA(cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem_small));
    SB_CUDA(cudaFuncSetAttribute(k_large, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem_large));
    SB_LAUNCH(k_small, nlong, 256, smem_small, st, ld, ptr, (const int64_t *)long_list,
              (const unsigned *)nullptr, idx_bits, out_idx, out_val, mid_list, counts, dc);
    SB_LAUNCH(k_large, nlong, 512, smem_large, st, ld, ptr, (const int64_t *)mid_list,
              (const unsigned *)counts, idx_bits, out_idx, out_val, huge_list, counts + 1, dc);
    SB_CUDA(cudaMemcpyAsync(&nlong, counts + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (nlong == 0) {
      launch_dup_fix<I, N, V, Loader>(ws, ld, ptr, n_seg, dc, sort_known, out_idx, out_val, vkind);
      return;
    }
    long_list = huge_list;
  }

  // ---- long segments: segmented radix sort on the index ----
  constexpr int kTileL = rs_seg_tile<UI, V, NoVal>();
  int64_t *offs = ws.alloc<int64_t>((int64_t)nlong + 1);
  int64_t *chunk_first = ws.alloc<int64_t>((int64_t)nlong + 1);
  exclusive_scan<int64_t>(ws, LongLenFn<N>{long_list, ptr}, offs, (int64_t)nlong);
  exclusive_scan<int64_t>(ws, LongChunksFn<N>{long_list, ptr, kTileL}, chunk_first,
                          (int64_t)nlong);
  int64_t total = 0, nchunks = 0;
  SB_CUDA(cudaMemcpyAsync(&total, offs + nlong, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&nchunks, chunk_first + nlong, sizeof(int64_t), cudaMemcpyDeviceToHost,
                          st));
  SB_CUDA(cudaStreamSynchronize(st));
  SB_REQUIRE(nchunks < (1ll << 31), SB200_ERR_BAD_ARG, "too many long-segment chunks");
  const int P = (idx_bits + kRsMaxBits - 1) / kRsMaxBits;
  RsSeg *seg = ws.alloc<RsSeg>(nchunks);
  int *chunk_seg = ws.alloc<int>(nchunks);
  SB_LAUNCH(ss_long_chunks_kernel, (unsigned)ceil_div(nchunks, 256), 256, 0, st,
            (const int64_t *)offs, (const int64_t *)chunk_first, (int64_t)nlong, nchunks, kTileL,
            seg, chunk_seg);
  UI *kin = ws.alloc<UI>(total), *ka = ws.alloc<UI>(total);
  UI *kb = P > 1 ? ws.alloc<UI>(total) : nullptr;
  V *vin = nullptr, *va = nullptr, *vb = nullptr;
  if constexpr (has_val<V>) {
    vin = ws.alloc<V>(total);
    va = ws.alloc<V>(total);
    vb = P > 1 ? ws.alloc<V>(total) : nullptr;
  }
  SB_LAUNCH((ss_long_fill_kernel<I, N, V, Loader>), (unsigned)nchunks, 256, 0, st, ld,
            (const int64_t *)long_list, (const int64_t *)offs, (const RsSeg *)seg,
            (const int *)chunk_seg, kin, vin, dc);
  radix_sort_segmented<UI, V, NoVal>(ws, {kin, vin, nullptr}, {ka, va, nullptr},
                                     {kb, vb, nullptr}, total, seg, nchunks, idx_bits);
  SB_LAUNCH((ss_long_store_kernel<I, N, V>), (unsigned)nchunks, 256, 0, st, ptr,
            (const int64_t *)long_list, (const int64_t *)offs, (const RsSeg *)seg,
            (const int *)chunk_seg, (const UI *)ka, (const V *)va, out_idx, out_val, dc);
  launch_dup_fix<I, N, V, Loader>(ws, ld, ptr, n_seg, dc, sort_known, out_idx, out_val, vkind);
}

}  // namespace sb200
