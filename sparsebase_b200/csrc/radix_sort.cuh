// radix_sort.cuh -- stable LSD radix sort of (key, payload1, payload2) records in HBM.
//
// This is the engine behind every "stable counting scatter" of the hot path:
//   * COO constructor sort by (row,col)                  (format/coo.cc:110-157)
//   * COO/CSR -> CSC stable scatter by column            (converter_order_two.cc:49-66)
//   * DegreeReorder's counting sort by degree            (reorder/degree_reorder.cc:33-46)
//   * the per-level (parent, degree, id) order of RCM    (reorder/rcm_reorder.cc:130-143)
//   * rows too long for the on-chip segmented sort       (format/csr.cc:123-157)
//
// Design (reduce-then-scan, deterministic, no spinning):
//   upsweep   : grid of `nchunks` CTAs, each histograms the current digit of its contiguous
//               chunk of tiles in shared memory (warp-private counters) -> spine[bin][chunk]
//   spine     : exclusive scan of the bin-major spine (decoupled look-back scan, scan.cuh)
//   downsweep : same chunks; per tile (8192 records by default): the tile arrives by bulk
//               async copy (cp.async.bulk + mbarrier, two stage buffers), warp-synchronous
//               stable ranking (ballot matching + warp-private counters), cross-warp digit scan,
//               records laid out in digit order in the stage buffer so that global writes are
//               coalesced runs; running per-bin offsets stay in smem
//               (rs_downsweep_pipe_kernel; rs_downsweep_kernel is the register-prefetch
//               version kept for unaligned inputs).
// Per pass HBM traffic: keys read twice, payloads read once, everything written once.
// radix_sort_segmented sorts every segment of a concatenation independently with the same
// kernels (tile-sized chunks per segment, segment-major spine).
#pragma once
#include <utility>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace sb200 {

// upsweep block (the downsweep block / tile are template parameters, see rs_config)
constexpr int kRsBlock = 256;
constexpr int kRsWarps = kRsBlock / 32;
constexpr int kRsMaxBits = 8;
constexpr int kRsMaxBins = 1 << kRsMaxBits;

template <typename K>
__device__ __forceinline__ unsigned rs_digit(K key, int shift, unsigned mask) {
  return (unsigned)(key >> shift) & mask;
}

// Segmented mode (radix_sort_segmented): the array is a concatenation of independent segments,
// every chunk is ONE tile-sized piece of one segment, and the spine is laid out
// [segment][bin][chunk of the segment], so that a flat exclusive scan of it yields
// (segment start) + (records of the segment with a smaller digit) + (same digit, earlier chunk).
struct RsSeg {
  int64_t begin;       // first record of the chunk
  int64_t spine_base;  // spine slot of (bin 0, this chunk)
  int count;           // records in the chunk (<= tile)
  int stride;          // chunks in this chunk's segment = spine distance between bins
};

struct RsChunking {
  int64_t n;
  int64_t tiles;  // in units of `tile` records (the downsweep tile of the chosen configuration)
  int nchunks;
  int tile;
  const RsSeg *seg = nullptr;  // non-null: segmented mode
  __host__ __device__ int64_t tile_begin(int c) const { return tiles * c / nchunks; }
  __host__ __device__ int64_t tile_end(int c) const { return tiles * (c + 1) / nchunks; }
};

// ------------------------------------------------------------------ upsweep
// Histogram of one digit over a chunk.  Keys are read with 16-byte vector loads, four per
// thread in flight (64 B/thread, ~64 KB/SM at 1024 resident threads) -- enough bytes in flight
// to cover the HBM bandwidth-delay product.
template <typename K>
__global__ void __launch_bounds__(kRsBlock)
    rs_upsweep_kernel(const K *__restrict__ keys, RsChunking ch, int shift, int bits,
                      int64_t *__restrict__ spine) {
  __shared__ unsigned hist[kRsWarps][kRsMaxBins];
  constexpr int kVec = 16 / sizeof(K);  // keys per 16-byte load
  const unsigned wid = threadIdx.x >> 5;
  const unsigned mask = (1u << bits) - 1u;
  for (int i = threadIdx.x; i < kRsWarps * kRsMaxBins; i += kRsBlock) (&hist[0][0])[i] = 0;
  __syncthreads();
  const int c = blockIdx.x;
  int64_t begin = ch.tile_begin(c) * ch.tile;
  int64_t end = ch.tile_end(c) * ch.tile;
  if (end > ch.n) end = ch.n;
  bool aligned = (reinterpret_cast<uintptr_t>(keys) & 15) == 0;
  RsSeg sg = {};
  if (ch.seg) {
    sg = ch.seg[c];
    begin = sg.begin;
    end = begin + sg.count;
    aligned = false;  // chunk starts are arbitrary
  }
  int64_t base = begin;
  if (aligned) {
    const int64_t step = (int64_t)kRsBlock * kVec * 4;
    for (; base + step <= end; base += step) {
      uint4 q[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
        q[k] = __ldcs(reinterpret_cast<const uint4 *>(keys + base) + k * kRsBlock + threadIdx.x);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if constexpr (sizeof(K) == 4) {
          atomicAdd(&hist[wid][rs_digit<K>(q[k].x, shift, mask)], 1u);
          atomicAdd(&hist[wid][rs_digit<K>(q[k].y, shift, mask)], 1u);
          atomicAdd(&hist[wid][rs_digit<K>(q[k].z, shift, mask)], 1u);
          atomicAdd(&hist[wid][rs_digit<K>(q[k].w, shift, mask)], 1u);
        } else {
          const K k0 = ((K)q[k].y << 32) | q[k].x, k1 = ((K)q[k].w << 32) | q[k].z;
          atomicAdd(&hist[wid][rs_digit<K>(k0, shift, mask)], 1u);
          atomicAdd(&hist[wid][rs_digit<K>(k1, shift, mask)], 1u);
        }
      }
    }
  }
  for (; base < end; base += (int64_t)kRsBlock * 4) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int64_t i = base + k * kRsBlock + threadIdx.x;
      if (i < end) atomicAdd(&hist[wid][rs_digit(ld_stream(keys + i), shift, mask)], 1u);
    }
  }
  __syncthreads();
  const int nbins = ch.seg ? kRsMaxBins : 1 << bits;  // segmented spine: fixed 256 bins
  for (int d = threadIdx.x; d < nbins; d += kRsBlock) {
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; w++) s += hist[w][d];
    if (ch.seg)
      spine[sg.spine_base + (int64_t)d * sg.stride] = s;
    else
      spine[(int64_t)d * ch.nchunks + c] = s;
  }
}

// ------------------------------------------------------------------ downsweep
// Per tile of BLOCK * IPT records:
//   1. stable rank of every record among the records of its digit inside its warp
//      (match_any: rounds in order, lanes in order) with warp-private digit counters
//   2. per digit: exclusive scan over the warps, then over the digits (tile histogram);
//      delta[d] = (running global offset of digit d for this chunk) - (first slot of d in the tile)
//   3. key AND payloads are staged in shared memory at their tile-local sorted slot
//   4. slot j goes to global position delta[digit(key_j)] + j: consecutive slots of one digit
//      are consecutive in global memory, so every warp store is a few contiguous runs, and the
//      destination is computed once for the three arrays
// Nothing of a tile is re-read from global memory; the keys of the next tile are requested
// before the current one is ranked.  Off = int32 when n < 2^31 halves the shared-memory traffic
// of step 4.
template <int BLOCK, int IPT, typename K, typename V1, typename V2, typename Off>
struct RsSmem {
  static constexpr int kTile = BLOCK * IPT;
  unsigned cnt[BLOCK / 32][kRsMaxBins];
  int64_t bin_off[kRsMaxBins];
  Off delta[kRsMaxBins];
  unsigned scan_scratch[34];
  K sk[kTile];
  typename std::conditional<has_val<V1>, V1, char>::type s1[has_val<V1> ? kTile : 1];
  typename std::conditional<has_val<V2>, V2, char>::type s2[has_val<V2> ? kTile : 1];
};

template <int BLOCK, int IPT, int MINB, typename K, typename V1, typename V2, typename Off>
__global__ void __launch_bounds__(BLOCK, MINB)
    rs_downsweep_kernel(const K *__restrict__ kin, K *__restrict__ kout,
                        const V1 *__restrict__ v1in, V1 *__restrict__ v1out,
                        const V2 *__restrict__ v2in, V2 *__restrict__ v2out, RsChunking ch,
                        int shift, int bits, const int64_t *__restrict__ spine) {
  using Smem = RsSmem<BLOCK, IPT, K, V1, V2, Off>;
  constexpr int kTile = BLOCK * IPT;
  extern __shared__ __align__(16) unsigned char rs_smem_raw[];
  Smem &s = *reinterpret_cast<Smem *>(rs_smem_raw);
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const unsigned mask = (1u << bits) - 1u;
  const int nbins = 1 << bits;
  const int c = blockIdx.x;
  const int64_t n = ch.n;

  if ((int)threadIdx.x < nbins)
    s.bin_off[threadIdx.x] = spine[(int64_t)threadIdx.x * ch.nchunks + c];

  K key[IPT], nkey[IPT];
  {
    const int64_t wb = ch.tile_begin(c) * kTile + (int64_t)wid * (IPT * 32);
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      const int64_t i = wb + r * 32 + lane;
      nkey[r] = i < n ? ld_stream(kin + i) : K(0);
    }
  }
  for (int64_t tile = ch.tile_begin(c); tile < ch.tile_end(c); tile++) {
    const int64_t tile_base_idx = tile * kTile;
    const int64_t rem = n - tile_base_idx;
    const int tile_count = rem < kTile ? (int)rem : kTile;
    const int64_t warp_base = tile_base_idx + (int64_t)wid * (IPT * 32);

    for (int i = threadIdx.x; i < (BLOCK / 32) * kRsMaxBins; i += BLOCK) (&s.cnt[0][0])[i] = 0;
#pragma unroll
    for (int r = 0; r < IPT; r++) key[r] = nkey[r];
    if (tile + 1 < ch.tile_end(c)) {
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const int64_t i = warp_base + kTile + r * 32 + lane;
        nkey[r] = i < n ? ld_stream(kin + i) : K(0);
      }
    }
    // payload loads of this tile are issued now and consumed after the ranking
    [[maybe_unused]] typename std::conditional<has_val<V1>, V1, char>::type p1[IPT];
    [[maybe_unused]] typename std::conditional<has_val<V2>, V2, char>::type p2[IPT];
    if constexpr (has_val<V1>) {
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const int64_t i = warp_base + r * 32 + lane;
        if (i < n) p1[r] = ld_stream(v1in + i);
      }
    }
    if constexpr (has_val<V2>) {
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const int64_t i = warp_base + r * 32 + lane;
        if (i < n) p2[r] = ld_stream(v2in + i);
      }
    }
    __syncthreads();  // counters zeroed; everybody is done with the previous tile's staging

    // ---- 1. stable ranking inside the warp ----
    unsigned lp[IPT];
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      const bool valid = warp_base + r * 32 + lane < n;
      const unsigned d = valid ? rs_digit(key[r], shift, mask) : 0xffffffffu;
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(peers) - 1;
      unsigned base = 0;
      if ((int)lane == leader && valid) {
        base = s.cnt[wid][d];
        s.cnt[wid][d] = base + __popc(peers);
      }
      base = __shfl_sync(0xffffffffu, base, leader);
      lp[r] = base + __popc(peers & lanemask_lt());
      __syncwarp();
    }
    __syncthreads();

    // ---- 2. per digit: exclusive scan over warps, then exclusive scan over digits ----
    unsigned hist = 0;
    if ((int)threadIdx.x < nbins) {
      unsigned run = 0;
#pragma unroll
      for (int w = 0; w < BLOCK / 32; w++) {
        const unsigned t = s.cnt[w][threadIdx.x];
        s.cnt[w][threadIdx.x] = run;
        run += t;
      }
      hist = run;
    }
    const unsigned excl = block_exclusive_scan(hist, s.scan_scratch);
    if ((int)threadIdx.x < nbins) {
#pragma unroll
      for (int w = 0; w < BLOCK / 32; w++) s.cnt[w][threadIdx.x] += excl;  // tile-local slot base
      const int64_t off = s.bin_off[threadIdx.x];
      s.delta[threadIdx.x] = (Off)(off - (int64_t)excl);
      s.bin_off[threadIdx.x] = off + hist;
    }
    __syncthreads();

    // ---- 3. stage the records at their tile-local sorted slot ----
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      if (warp_base + r * 32 + lane < n) {
        const unsigned d = rs_digit(key[r], shift, mask);
        const unsigned at = lp[r] + s.cnt[wid][d];
        s.sk[at] = key[r];
        if constexpr (has_val<V1>) s.s1[at] = p1[r];
        if constexpr (has_val<V2>) s.s2[at] = p2[r];
      }
    }
    __syncthreads();

    // ---- 4. coalesced runs to global memory ----
#pragma unroll
    for (int k = 0; k < IPT; k++) {
      const int j = k * BLOCK + (int)threadIdx.x;
      if (j < tile_count) {
        const K kk = s.sk[j];
        const int64_t dst = (int64_t)s.delta[rs_digit(kk, shift, mask)] + j;
        kout[dst] = kk;
        if constexpr (has_val<V1>) v1out[dst] = s.s1[j];
        if constexpr (has_val<V2>) v2out[dst] = s.s2[j];
      }
    }
    // no barrier here: the next tile writes cnt (last read in step 3), and touches delta and
    // the staging arrays only after its own first two barriers
  }
}

// ------------------------------------------------------------------ downsweep, bulk-copy pipeline
// Same ranking and the same results as rs_downsweep_kernel, restructured around the load latency
// that bounded it (ncu: 47 % of the stall samples were long-scoreboard waits on the tile's own
// key / payload loads, issue rate 0.15 per scheduler):
//   * the whole tile (keys + payloads, 12 B x 4096 records = 48 KB) is fetched by ONE thread with
//     three 1-D bulk async copies (cp.async.bulk global -> shared, completion on an mbarrier);
//     two stage buffers, so the copy of tile t+1 is in flight for the whole of tile t and no
//     register is spent on prefetching;
//   * a stage buffer is read into registers as soon as it lands and is then reused as the staging
//     area in which the tile is laid out in digit order before the coalesced global writes;
//   * all match_any instructions of a tile are issued back to back before the serial counter
//     updates; warp-private counters are 16 bits (a tile has < 65536 records).
// Tiles that cannot be bulk-copied (the partial last tile) are read with plain loads.
__device__ __forceinline__ uint32_t rs_smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void rs_mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rs_smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rs_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint64_t *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(rs_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void rs_bulk_g2s(void *dst, const void *src, unsigned bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(rs_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(rs_smem_u32(bar))
      : "memory");
}

// Lanes holding the same 8-bit digit, from eight ballots.  MATCH.ANY has a data-dependent cost
// (one step per distinct value in the warp; digits of a tile are almost all distinct) and was
// 38 % of the stall samples of this kernel; VOTE is a fixed-latency instruction.
__device__ __forceinline__ unsigned rs_match_digit(unsigned d) {
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < kRsMaxBits; b++) {
    const bool bit = (d & (1u << b)) != 0u;  // one LOP3 with a predicate result
    const unsigned bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

template <int BLOCK, int IPT, typename K, typename V1, typename V2, typename Off>
struct RsPipeSmem {
  static constexpr int kTile = BLOCK * IPT;
  using P1 = typename std::conditional<has_val<V1>, V1, char>::type;
  using P2 = typename std::conditional<has_val<V2>, V2, char>::type;
  __align__(16) K sk[2][kTile];
  __align__(16) P1 s1[2][has_val<V1> ? kTile : 16];
  __align__(16) P2 s2[2][has_val<V2> ? kTile : 16];
  unsigned short cnt[BLOCK / 32][kRsMaxBins];
  unsigned short excl[kRsMaxBins];
  int64_t bin_off[kRsMaxBins];
  Off delta[kRsMaxBins];
  unsigned scan_scratch[34];
  __align__(8) uint64_t bar[2];
};

template <int BLOCK, int IPT, int MINB, bool SEG, typename K, typename V1, typename V2,
          typename Off>
__global__ void __launch_bounds__(BLOCK, MINB)
    rs_downsweep_pipe_kernel(const K *__restrict__ kin, K *__restrict__ kout,
                             const V1 *__restrict__ v1in, V1 *__restrict__ v1out,
                             const V2 *__restrict__ v2in, V2 *__restrict__ v2out, RsChunking ch,
                             int shift, int bits, const int64_t *__restrict__ spine) {
  using Smem = RsPipeSmem<BLOCK, IPT, K, V1, V2, Off>;
  constexpr int kTile = BLOCK * IPT;
  static_assert(kTile < 65536, "16-bit slot counters");
  constexpr unsigned kTileBytes =
      (unsigned)(kTile * (sizeof(K) + (has_val<V1> ? sizeof(V1) : 0) +
                          (has_val<V2> ? sizeof(V2) : 0)));
  extern __shared__ __align__(128) unsigned char rs_pipe_smem_raw[];
  Smem &s = *reinterpret_cast<Smem *>(rs_pipe_smem_raw);
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const unsigned mask = (1u << bits) - 1u;
  const int nbins = 1 << bits;
  const int c = blockIdx.x;
  const int64_t n = ch.n;
  RsSeg sg = {};
  if (SEG) sg = ch.seg[c];
  // segmented mode: the chunk is one (possibly short, arbitrarily aligned) tile
  const int64_t t_begin = SEG ? 0 : ch.tile_begin(c), t_end = SEG ? 1 : ch.tile_end(c);
  if (t_begin >= t_end) return;

  auto issue = [&](int64_t tile, int st) {  // one thread
    rs_mbar_expect_tx(&s.bar[st], kTileBytes);
    const int64_t base = tile * kTile;
    rs_bulk_g2s(s.sk[st], kin + base, (unsigned)(kTile * sizeof(K)), &s.bar[st]);
    if constexpr (has_val<V1>)
      rs_bulk_g2s(s.s1[st], v1in + base, (unsigned)(kTile * sizeof(V1)), &s.bar[st]);
    if constexpr (has_val<V2>)
      rs_bulk_g2s(s.s2[st], v2in + base, (unsigned)(kTile * sizeof(V2)), &s.bar[st]);
  };
  auto full = [&](int64_t tile) { return !SEG && (tile + 1) * kTile <= n; };

  if (threadIdx.x == 0) {
    rs_mbar_init(&s.bar[0], 1);
    rs_mbar_init(&s.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (full(t_begin)) issue(t_begin, 0);
    if (t_begin + 1 < t_end && full(t_begin + 1)) issue(t_begin + 1, 1);
  }
  if ((int)threadIdx.x < nbins)
    s.bin_off[threadIdx.x] = SEG ? spine[sg.spine_base + (int64_t)threadIdx.x * sg.stride]
                                    : spine[(int64_t)threadIdx.x * ch.nchunks + c];
  __syncthreads();  // barriers initialised before anybody waits on them

  unsigned phase0 = 0, phase1 = 0;
  for (int64_t tile = t_begin; tile < t_end; tile++) {
    const int st = (int)((tile - t_begin) & 1);
    const int64_t tile_base_idx = SEG ? sg.begin : tile * kTile;
    const int tile_count = SEG ? sg.count : (full(tile) ? kTile : (int)(n - tile_base_idx));
    const bool bulk_tile = full(tile);
    const int local_base = (int)wid * (IPT * 32) + (int)lane;
    K *const sk = s.sk[st];
    [[maybe_unused]] typename Smem::P1 *const s1 = s.s1[st];
    [[maybe_unused]] typename Smem::P2 *const s2 = s.s2[st];

    if (bulk_tile) {
      if (st == 0) {
        rs_mbar_wait(&s.bar[0], phase0);
        phase0 ^= 1u;
      } else {
        rs_mbar_wait(&s.bar[1], phase1);
        phase1 ^= 1u;
      }
    } else {
      // the partial last tile of the array: plain loads into the stage buffer, padded with
      // all-ones keys.  Padding has the largest digit and comes last in the tile, so it ranks
      // behind every real record (slots >= tile_count, never written) and only inflates the
      // count of the last bin of the last tile, which nobody reads any more.
      const int pad_end = SEG ? (tile_count + 31) & ~31 : kTile;  // SEG: later rounds are skipped
      for (int q = threadIdx.x; q < pad_end; q += BLOCK) {
        if (q < tile_count) {
          sk[q] = ld_stream(kin + tile_base_idx + q);
          if constexpr (has_val<V1>) s1[q] = ld_stream(v1in + tile_base_idx + q);
          if constexpr (has_val<V2>) s2[q] = ld_stream(v2in + tile_base_idx + q);
        } else {
          sk[q] = (K)~K(0);
        }
      }
      __syncthreads();
    }

    K key[IPT];
    [[maybe_unused]] typename Smem::P1 p1[IPT];
    [[maybe_unused]] typename Smem::P2 p2[IPT];
    // rounds of this warp that hold records (all of them except in a short segmented chunk,
    // where whole warps have nothing to rank)
    const int warp_first = (int)wid * (IPT * 32);
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      const int q = local_base + r * 32;
      if (!SEG || warp_first + r * 32 < tile_count) {
        key[r] = sk[q];
        if constexpr (has_val<V1>) p1[r] = s1[q];
        if constexpr (has_val<V2>) p2[r] = s2[q];
      }
    }
    {
      unsigned *z = reinterpret_cast<unsigned *>(&s.cnt[0][0]);
#pragma unroll
      for (int i = 0; i < (BLOCK / 32) * kRsMaxBins / 2 / BLOCK; i++) z[i * BLOCK + threadIdx.x] = 0;
    }
    // ---- 1a. all the matches of the tile, back to back ----
    unsigned peers[IPT];
#pragma unroll
    for (int r = 0; r < IPT; r++)
      if (!SEG || warp_first + r * 32 < tile_count)
        peers[r] = rs_match_digit(rs_digit(key[r], shift, mask));
    __syncthreads();  // stage buffer is in registers everywhere; counters are zero

    // ---- 1b. stable ranking inside the warp (rounds in order, lanes in order) ----
    unsigned short *const my_cnt = s.cnt[wid];
    unsigned lp[IPT];  // (rank inside the warp << 8) | digit
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      if (SEG && warp_first + r * 32 >= tile_count) break;
      const unsigned d = rs_digit(key[r], shift, mask);
      const int leader = __ffs(peers[r]) - 1;
      unsigned base = 0;
      if ((int)lane == leader) {
        base = my_cnt[d];
        my_cnt[d] = (unsigned short)(base + __popc(peers[r]));
      }
      base = __shfl_sync(0xffffffffu, base, leader);
      lp[r] = ((base + __popc(peers[r] & lanemask_lt())) << 8) | d;
      __syncwarp();
    }
    __syncthreads();

    // ---- 2. per digit: exclusive scan over warps, then exclusive scan over digits ----
    unsigned hist = 0;
    if ((int)threadIdx.x < nbins) {
#pragma unroll
      for (int w = 0; w < BLOCK / 32; w++) {
        const unsigned t = s.cnt[w][threadIdx.x];
        s.cnt[w][threadIdx.x] = (unsigned short)hist;
        hist += t;
      }
    }
    const unsigned excl = block_exclusive_scan(hist, s.scan_scratch);
    if ((int)threadIdx.x < nbins) {
      s.excl[threadIdx.x] = (unsigned short)excl;  // first tile-local slot of the digit
      const int64_t off = s.bin_off[threadIdx.x];
      s.delta[threadIdx.x] = (Off)(off - (int64_t)excl);
      s.bin_off[threadIdx.x] = off + hist;
    }
    __syncthreads();

    // ---- 3. lay the tile out in digit order in the (now free) stage buffer ----
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      if (SEG && warp_first + r * 32 >= tile_count) break;
      const unsigned d = lp[r] & 0xffu;
      const unsigned at = (lp[r] >> 8) + my_cnt[d] + s.excl[d];
      sk[at] = key[r];
      if constexpr (has_val<V1>) s1[at] = p1[r];
      if constexpr (has_val<V2>) s2[at] = p2[r];
    }
    __syncthreads();

    // ---- 4. coalesced runs to global memory (element index in Off: one IMAD.WIDE per array
    //         when the array has fewer than 2^31 records) ----
#pragma unroll
    for (int k = 0; k < IPT; k++) {
      const int j = k * BLOCK + (int)threadIdx.x;
      if (j < tile_count) {
        const K kk = sk[j];
        const Off dst = s.delta[rs_digit(kk, shift, mask)] + (Off)j;
        kout[dst] = kk;
        if constexpr (has_val<V1>) v1out[dst] = s1[j];
        if constexpr (has_val<V2>) v2out[dst] = s2[j];
      }
    }
    __syncthreads();  // the stage buffer has been read; it may be overwritten by the next copy
    if (threadIdx.x == 0 && tile + 2 < t_end && full(tile + 2)) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(tile + 2, st);
    }
  }
}

// ------------------------------------------------------------------ host driver
struct RsBitRange {
  int begin, end;  // sort on key bits [begin, end)
};

inline int bits_for(uint64_t max_value) {  // number of bits needed to represent max_value
  int b = 0;
  while (max_value) {
    b++;
    max_value >>= 1;
  }
  return b;
}

inline int rs_num_passes(const std::vector<RsBitRange> &ranges) {
  int p = 0;
  for (const RsBitRange &rg : ranges)
    if (rg.end > rg.begin) p += (rg.end - rg.begin + kRsMaxBits - 1) / kRsMaxBits;
  return p;
}

template <typename T>
struct LoadFn {
  const T *p;
  __device__ T operator()(int64_t i) const { return p[i]; }
};

template <typename K, typename V1, typename V2>
struct RsBufs {
  K *k;
  V1 *v1;
  V2 *v2;
};

// One downsweep launch for a fixed (BLOCK, IPT, MINB) configuration.
template <int BLOCK, int IPT, int MINB, typename K, typename V1, typename V2>
void rs_launch_downsweep(cudaStream_t st, RsBufs<K, V1, V2> src, RsBufs<K, V1, V2> dst,
                         const RsChunking &ch, int shift, int bits, const int64_t *spine) {
  if (ch.n < (1ll << 31)) {
    auto kern = rs_downsweep_kernel<BLOCK, IPT, MINB, K, V1, V2, int32_t>;
    constexpr int smem = (int)sizeof(RsSmem<BLOCK, IPT, K, V1, V2, int32_t>);
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SB_LAUNCH(kern, ch.nchunks, BLOCK, smem, st, (const K *)src.k, dst.k, (const V1 *)src.v1,
              dst.v1, (const V2 *)src.v2, dst.v2, ch, shift, bits, spine);
  } else {
    auto kern = rs_downsweep_kernel<BLOCK, IPT, MINB, K, V1, V2, int64_t>;
    constexpr int smem = (int)sizeof(RsSmem<BLOCK, IPT, K, V1, V2, int64_t>);
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SB_LAUNCH(kern, ch.nchunks, BLOCK, smem, st, (const K *)src.k, dst.k, (const V1 *)src.v1,
              dst.v1, (const V2 *)src.v2, dst.v2, ch, shift, bits, spine);
  }
}

// One launch of the bulk-copy pipelined downsweep (inputs must be 16-byte aligned).  Records
// narrow enough for two resident CTAs per SM are compiled for 64 registers; wider ones (one
// CTA per SM by shared memory anyway) get the full register file.
template <int BLOCK, int IPT, bool SEG, typename K, typename V1, typename V2, typename Off>
void rs_launch_downsweep_pipe_off(cudaStream_t st, RsBufs<K, V1, V2> src, RsBufs<K, V1, V2> dst,
                                  const RsChunking &ch, int shift, int bits,
                                  const int64_t *spine) {
  constexpr int smem = (int)sizeof(RsPipeSmem<BLOCK, IPT, K, V1, V2, Off>);
  constexpr int MINB = (smem <= 112 * 1024 && BLOCK <= 512 && IPT <= 8) ? 2 : 1;
  SB_REQUIRE(SEG == (ch.seg != nullptr), SB200_ERR_BAD_ARG, "segmented-mode mismatch");
  auto kern = rs_downsweep_pipe_kernel<BLOCK, IPT, MINB, SEG, K, V1, V2, Off>;
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  SB_LAUNCH(kern, ch.nchunks, BLOCK, smem, st, (const K *)src.k, dst.k, (const V1 *)src.v1,
            dst.v1, (const V2 *)src.v2, dst.v2, ch, shift, bits, spine);
}
template <int BLOCK, int IPT, typename K, typename V1, typename V2, bool SEG = false>
void rs_launch_downsweep_pipe(cudaStream_t st, RsBufs<K, V1, V2> src, RsBufs<K, V1, V2> dst,
                              const RsChunking &ch, int shift, int bits, const int64_t *spine) {
  if (ch.n < (1ll << 31))
    rs_launch_downsweep_pipe_off<BLOCK, IPT, SEG, K, V1, V2, int32_t>(st, src, dst, ch, shift, bits,
                                                                 spine);
  else
    rs_launch_downsweep_pipe_off<BLOCK, IPT, SEG, K, V1, V2, int64_t>(st, src, dst, ch, shift, bits,
                                                                 spine);
}

// Default shape of the pipelined downsweep: 512 threads, ONE CTA per SM with the full register
// file, and the largest tile whose two stage buffers fit the 227 KB of shared memory -- 8192
// records (16 per thread) for records of up to 12 bytes.  Per tile and bin a CTA writes a run of
// tile/256 records; doubling the tile from 4096 doubled the runs to 128 bytes and took 30 % off
// the pass (CSR->CSC on C2 3.6 -> 2.6 ms, on C3 11.4 -> 7.9 ms) although half as many warps are
// resident.
template <typename K, typename V1, typename V2>
constexpr int rs_auto_ipt() {
  constexpr int rec = (int)sizeof(K) + (has_val<V1> ? (int)sizeof(V1) : 0) +
                      (has_val<V2> ? (int)sizeof(V2) : 0);
  constexpr int budget = 212 * 1024;  // stage buffers; counters, offsets and barriers on top
  return 2 * 512 * 16 * rec <= budget ? 16 : (2 * 512 * 12 * rec <= budget ? 12 : 8);
}

template <typename K, typename V1, typename V2>
inline bool rs_bulk_aligned(const RsBufs<K, V1, V2> &b) {
  uintptr_t a = reinterpret_cast<uintptr_t>(b.k);
  if constexpr (has_val<V1>) a |= reinterpret_cast<uintptr_t>(b.v1);
  if constexpr (has_val<V2>) a |= reinterpret_cast<uintptr_t>(b.v2);
  return (a & 15) == 0;
}

// Downsweep configuration.  Default (0): bulk-copy pipelined kernel in the shape chosen by
// rs_auto_ipt.  SB200_RS_CONFIG=5 is its 4096-record / two-CTA shape and 10 the
// register-prefetch kernel (also the route for unaligned inputs); both kept for A/B runs.
inline int rs_config() {
  static const int cfg = [] {
    const char *e = getenv("SB200_RS_CONFIG");
    return e ? atoi(e) : 0;
  }();
  return cfg;
}
inline int rs_chunks_per_sm() {
  static const int v = [] {
    const char *e = getenv("SB200_RS_CHUNKS_PER_SM");
    const int x = e ? atoi(e) : 2;
    return x > 0 ? x : 2;
  }();
  return v;
}
template <typename K, typename V1, typename V2>
inline int rs_config_tile(int cfg, bool bulk) {
  if (cfg == 0) return bulk ? 512 * rs_auto_ipt<K, V1, V2>() : 512 * 8;
  return 512 * 8;  // cfg 5 (pipelined, two CTAs per SM) and cfg 10 (register-prefetch kernel)
}

// Stable sort of every segment of a concatenation of segments by the key bits [0, key_bits):
// `seg` holds `nchunks` tile-sized chunk descriptors (device memory, tile = rs_seg_tile<...>()),
// built by the caller.  Same buffer convention as radix_sort.  Records never leave their segment.
template <typename K, typename V1, typename V2>
constexpr int rs_seg_tile() {
  // smaller than the default tile: most segments are not much longer than the caller's
  // threshold, and a chunk ranks whole 32-record rounds (two CTAs per SM here)
  return 512 * 8;
}
template <typename K, typename V1, typename V2>
void radix_sort_segmented(Workspace &ws, RsBufs<K, V1, V2> in, RsBufs<K, V1, V2> out,
                          RsBufs<K, V1, V2> tmp, int64_t n, const RsSeg *seg, int64_t nchunks,
                          int key_bits) {
  if (n <= 0 || nchunks <= 0) return;
  cudaStream_t st = ws.stream();
  const int P = (key_bits + kRsMaxBits - 1) / kRsMaxBits;
  if (P == 0) {
    SB_CUDA(cudaMemcpyAsync(out.k, in.k, n * sizeof(K), cudaMemcpyDeviceToDevice, st));
    if constexpr (has_val<V1>)
      SB_CUDA(cudaMemcpyAsync(out.v1, in.v1, n * sizeof(V1), cudaMemcpyDeviceToDevice, st));
    if constexpr (has_val<V2>)
      SB_CUDA(cudaMemcpyAsync(out.v2, in.v2, n * sizeof(V2), cudaMemcpyDeviceToDevice, st));
    return;
  }
  RsChunking ch;
  ch.n = n;
  ch.tile = rs_seg_tile<K, V1, V2>();
  ch.tiles = nchunks;
  ch.nchunks = (int)nchunks;
  ch.seg = seg;
  const int64_t spine_len = (int64_t)kRsMaxBins * nchunks;
  int64_t *spine_in = ws.alloc<int64_t>(spine_len + 1);
  int64_t *spine = ws.alloc<int64_t>(spine_len + 1);
  RsBufs<K, V1, V2> src = in;
  int shift = 0;
  for (int pass = 0; pass < P; pass++) {
    const int bits = (key_bits - shift + (P - pass) - 1) / (P - pass);
    RsBufs<K, V1, V2> dst = ((P - 1 - pass) % 2 == 0) ? out : tmp;
    SB_LAUNCH((rs_upsweep_kernel<K>), ch.nchunks, kRsBlock, 0, st, (const K *)src.k, ch, shift,
              bits, spine_in);
    exclusive_scan<int64_t>(ws, LoadFn<int64_t>{spine_in}, spine, spine_len);
    rs_launch_downsweep_pipe<512, 8, K, V1, V2, true>(st, src, dst, ch, shift, bits, spine);
    src = dst;
    shift += bits;
  }
}

// Stable sort of n records by the given key bit ranges (least significant range first).
// `in` is only read; the sorted records always land in `out`; `tmp` is scratch of the same
// size and is only touched when more than one pass is needed (rs_num_passes(ranges) > 1).
template <typename K, typename V1, typename V2>
void radix_sort(Workspace &ws, RsBufs<K, V1, V2> in, RsBufs<K, V1, V2> out,
                RsBufs<K, V1, V2> tmp, int64_t n, const std::vector<RsBitRange> &ranges) {
  if (n <= 0) return;
  cudaStream_t st = ws.stream();
  const int P = n > 1 ? rs_num_passes(ranges) : 0;
  if (P == 0) {
    SB_CUDA(cudaMemcpyAsync(out.k, in.k, n * sizeof(K), cudaMemcpyDeviceToDevice, st));
    if constexpr (has_val<V1>)
      SB_CUDA(cudaMemcpyAsync(out.v1, in.v1, n * sizeof(V1), cudaMemcpyDeviceToDevice, st));
    if constexpr (has_val<V2>)
      SB_CUDA(cudaMemcpyAsync(out.v2, in.v2, n * sizeof(V2), cudaMemcpyDeviceToDevice, st));
    return;
  }
  const DeviceInfo &di = device_info(ws.device());
  const int cfg = rs_config();
  // the bulk-copy kernel needs 16-byte aligned inputs in every pass: `in` for the first one,
  // then out / tmp alternately
  const bool bulk = rs_bulk_aligned(in) && rs_bulk_aligned(out) && (P < 2 || rs_bulk_aligned(tmp));
  RsChunking ch;
  ch.n = n;
  ch.tile = rs_config_tile<K, V1, V2>(cfg, bulk);
  ch.tiles = ceil_div(n, ch.tile);
  int64_t max_chunks = (int64_t)di.sm_count * rs_chunks_per_sm();
  ch.nchunks = (int)(ch.tiles < max_chunks ? ch.tiles : max_chunks);
  int64_t *spine_in = ws.alloc<int64_t>((int64_t)kRsMaxBins * ch.nchunks + 1);
  int64_t *spine = ws.alloc<int64_t>((int64_t)kRsMaxBins * ch.nchunks + 1);
  int pass = 0;
  RsBufs<K, V1, V2> src = in;
  for (const RsBitRange &rg : ranges) {
    const int total = rg.end - rg.begin;
    if (total <= 0) continue;
    const int passes = (total + kRsMaxBits - 1) / kRsMaxBits;
    int shift = rg.begin;
    for (int p = 0; p < passes; p++, pass++) {
      const int bits = (total - (shift - rg.begin) + (passes - p) - 1) / (passes - p);
      RsBufs<K, V1, V2> dst = ((P - 1 - pass) % 2 == 0) ? out : tmp;
      SB_LAUNCH((rs_upsweep_kernel<K>), ch.nchunks, kRsBlock, 0, st, (const K *)src.k, ch, shift,
                bits, spine_in);
      exclusive_scan<int64_t>(ws, LoadFn<int64_t>{spine_in}, spine,
                              (int64_t)(1 << bits) * ch.nchunks);
      if (cfg == 5 && bulk)
        rs_launch_downsweep_pipe<512, 8, K, V1, V2>(st, src, dst, ch, shift, bits, spine);
      else if (cfg == 0 && bulk)
        rs_launch_downsweep_pipe<512, rs_auto_ipt<K, V1, V2>(), K, V1, V2>(st, src, dst, ch, shift,
                                                                         bits, spine);
      else
        rs_launch_downsweep<512, 8, 2, K, V1, V2>(st, src, dst, ch, shift, bits, spine);
      src = dst;
      shift += bits;
    }
  }
}

}  // namespace sb200
