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
//   downsweep : same chunks; per 4096-key tile: warp-synchronous match_any ranking (stable),
//               cross-warp digit scan, records staged through shared memory in digit order so
//               that global writes are coalesced runs; running per-bin offsets stay in smem.
// Per pass HBM traffic: keys read twice, payloads read once, everything written once.
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

struct RsChunking {
  int64_t n;
  int64_t tiles;  // in units of `tile` records (the downsweep tile of the chosen configuration)
  int nchunks;
  int tile;
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
  const int64_t begin = ch.tile_begin(c) * ch.tile;
  int64_t end = ch.tile_end(c) * ch.tile;
  if (end > ch.n) end = ch.n;
  const bool aligned = (reinterpret_cast<uintptr_t>(keys) & 15) == 0;
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
  const int nbins = 1 << bits;
  for (int d = threadIdx.x; d < nbins; d += kRsBlock) {
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; w++) s += hist[w][d];
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

// Downsweep configuration: 512 threads x 8 records (4096-record tile, two CTAs per SM, 32
// resident warps).  SB200_RS_CONFIG selects the alternatives kept for tuning runs.
inline int rs_config() {
  static const int cfg = [] {
    const char *e = getenv("SB200_RS_CONFIG");
    return e ? atoi(e) : 0;
  }();
  return cfg;
}
inline int rs_config_tile(int cfg) {
  switch (cfg) {
    case 1: return 256 * 12;
    case 2: return 512 * 12;
    case 3: return 384 * 10;
    case 4: return 256 * 16;
    default: return 512 * 8;
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
  RsChunking ch;
  ch.n = n;
  ch.tile = rs_config_tile(cfg);
  ch.tiles = ceil_div(n, ch.tile);
  int64_t max_chunks = (int64_t)di.sm_count * 4;
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
      if (cfg == 1)
        rs_launch_downsweep<256, 12, 2, K, V1, V2>(st, src, dst, ch, shift, bits, spine);
      else if (cfg == 2)
        rs_launch_downsweep<512, 12, 1, K, V1, V2>(st, src, dst, ch, shift, bits, spine);
      else if (cfg == 3)
        rs_launch_downsweep<384, 10, 2, K, V1, V2>(st, src, dst, ch, shift, bits, spine);
      else if (cfg == 4)
        rs_launch_downsweep<256, 16, 2, K, V1, V2>(st, src, dst, ch, shift, bits, spine);
      else
        rs_launch_downsweep<512, 8, 2, K, V1, V2>(st, src, dst, ch, shift, bits, spine);
      src = dst;
      shift += bits;
    }
  }
}

}  // namespace sb200
