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

constexpr int kRsBlock = 256;
constexpr int kRsWarps = kRsBlock / 32;
constexpr int kRsIpt = 12;
constexpr int kRsTile = kRsBlock * kRsIpt;  // 3072 records per tile
constexpr int kRsMaxBits = 8;
constexpr int kRsMaxBins = 1 << kRsMaxBits;

template <typename K>
__device__ __forceinline__ unsigned rs_digit(K key, int shift, unsigned mask) {
  return (unsigned)(key >> shift) & mask;
}

struct RsChunking {
  int64_t n;
  int64_t tiles;
  int nchunks;
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
  const int64_t begin = ch.tile_begin(c) * kRsTile;
  int64_t end = ch.tile_end(c) * kRsTile;
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
struct RsSmem {
  unsigned cnt[kRsWarps][kRsMaxBins];
  int64_t bin_off[kRsMaxBins];
  int64_t delta[kRsMaxBins];
  unsigned tile_base[kRsMaxBins];
  unsigned scan_scratch[34];
  unsigned char sdig[kRsTile];
  uint64_t stage[kRsTile];
};

template <typename T>
__device__ __forceinline__ void rs_load_payload(const T *__restrict__ in, int64_t warp_base,
                                                int64_t n, T (&v)[kRsIpt]) {
  const unsigned lane = lane_id();
#pragma unroll
  for (int r = 0; r < kRsIpt; r++) {
    int64_t i = warp_base + r * 32 + lane;
    if (i < n) v[r] = ld_stream(in + i);
  }
}

template <typename T>
__device__ __forceinline__ void rs_store_payload(const T (&v)[kRsIpt], T *__restrict__ out,
                                                 RsSmem &s, int64_t warp_base, int64_t n,
                                                 const unsigned (&lp)[kRsIpt], int tile_count) {
  T *stage = reinterpret_cast<T *>(s.stage);
  const unsigned lane = lane_id();
#pragma unroll
  for (int r = 0; r < kRsIpt; r++) {
    int64_t i = warp_base + r * 32 + lane;
    if (i < n) stage[lp[r]] = v[r];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kRsIpt; k++) {
    int j = k * kRsBlock + threadIdx.x;
    if (j < tile_count) out[s.delta[s.sdig[j]] + j] = stage[j];
  }
  __syncthreads();
}

template <typename K, typename V1, typename V2>
__global__ void __launch_bounds__(kRsBlock, 2)
    rs_downsweep_kernel(const K *__restrict__ kin, K *__restrict__ kout,
                        const V1 *__restrict__ v1in, V1 *__restrict__ v1out,
                        const V2 *__restrict__ v2in, V2 *__restrict__ v2out, RsChunking ch,
                        int shift, int bits, const int64_t *__restrict__ spine) {
  extern __shared__ __align__(16) unsigned char rs_smem_raw[];
  RsSmem &s = *reinterpret_cast<RsSmem *>(rs_smem_raw);
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const unsigned mask = (1u << bits) - 1u;
  const int nbins = 1 << bits;
  const int c = blockIdx.x;
  const int64_t n = ch.n;

  if ((int)threadIdx.x < nbins)
    s.bin_off[threadIdx.x] = spine[(int64_t)threadIdx.x * ch.nchunks + c];

  // software pipeline: the keys of tile t+1 are requested before tile t is processed, so the
  // ranking of a tile never waits for its own loads (ncu: 37% of the stall samples sat on the
  // first use of the freshly loaded keys before this)
  K key[kRsIpt], nkey[kRsIpt];
  {
    const int64_t wb = ch.tile_begin(c) * kRsTile + (int64_t)wid * (kRsIpt * 32);
#pragma unroll
    for (int r = 0; r < kRsIpt; r++) {
      int64_t i = wb + r * 32 + lane;
      nkey[r] = i < n ? ld_stream(kin + i) : K(0);
    }
  }
  for (int64_t tile = ch.tile_begin(c); tile < ch.tile_end(c); tile++) {
    const int64_t tile_base_idx = tile * kRsTile;
    const int64_t rem = n - tile_base_idx;
    const int tile_count = rem < kRsTile ? (int)rem : kRsTile;
    const int64_t warp_base = tile_base_idx + (int64_t)wid * (kRsIpt * 32);

    for (int i = threadIdx.x; i < kRsWarps * kRsMaxBins; i += kRsBlock) (&s.cnt[0][0])[i] = 0;

#pragma unroll
    for (int r = 0; r < kRsIpt; r++) key[r] = nkey[r];
    if (tile + 1 < ch.tile_end(c)) {
#pragma unroll
      for (int r = 0; r < kRsIpt; r++) {
        int64_t i = warp_base + kRsTile + r * 32 + lane;
        nkey[r] = i < n ? ld_stream(kin + i) : K(0);
      }
    }
    // payload loads of this tile are issued now and consumed after the ranking
    [[maybe_unused]] typename std::conditional<has_val<V1>, V1, char>::type p1[kRsIpt];
    [[maybe_unused]] typename std::conditional<has_val<V2>, V2, char>::type p2[kRsIpt];
    if constexpr (has_val<V1>) rs_load_payload<V1>(v1in, warp_base, n, p1);
    if constexpr (has_val<V2>) rs_load_payload<V2>(v2in, warp_base, n, p2);
    __syncthreads();  // counters zeroed

    // ---- stable ranking inside the warp: rounds in order, lanes in order ----
    unsigned lp[kRsIpt];
#pragma unroll
    for (int r = 0; r < kRsIpt; r++) {
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

    // ---- per digit: exclusive scan over warps, then exclusive scan over digits ----
    unsigned hist = 0;
    if ((int)threadIdx.x < nbins) {
      unsigned run = 0;
#pragma unroll
      for (int w = 0; w < kRsWarps; w++) {
        unsigned t = s.cnt[w][threadIdx.x];
        s.cnt[w][threadIdx.x] = run;
        run += t;
      }
      hist = run;
    }
    unsigned excl = block_exclusive_scan(hist, s.scan_scratch);
    if ((int)threadIdx.x < nbins) {
      s.tile_base[threadIdx.x] = excl;
      int64_t off = s.bin_off[threadIdx.x];
      s.delta[threadIdx.x] = off - (int64_t)excl;
      s.bin_off[threadIdx.x] = off + hist;
    }
    __syncthreads();

    // ---- keys: stage in digit order, then coalesced runs to global ----
    K *stage_k = reinterpret_cast<K *>(s.stage);
#pragma unroll
    for (int r = 0; r < kRsIpt; r++) {
      if (warp_base + r * 32 + lane < n) {
        const unsigned d = rs_digit(key[r], shift, mask);
        lp[r] += s.tile_base[d] + s.cnt[wid][d];
        stage_k[lp[r]] = key[r];
        s.sdig[lp[r]] = (unsigned char)d;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRsIpt; k++) {
      int j = k * kRsBlock + threadIdx.x;
      if (j < tile_count) kout[s.delta[s.sdig[j]] + j] = stage_k[j];
    }
    __syncthreads();
    if constexpr (has_val<V1>) rs_store_payload<V1>(p1, v1out, s, warp_base, n, lp, tile_count);
    if constexpr (has_val<V2>) rs_store_payload<V2>(p2, v2out, s, warp_base, n, lp, tile_count);
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
  RsChunking ch;
  ch.n = n;
  ch.tiles = ceil_div(n, kRsTile);
  int64_t max_chunks = (int64_t)di.sm_count * 4;
  ch.nchunks = (int)(ch.tiles < max_chunks ? ch.tiles : max_chunks);
  int64_t *spine_in = ws.alloc<int64_t>((int64_t)kRsMaxBins * ch.nchunks + 1);
  int64_t *spine = ws.alloc<int64_t>((int64_t)kRsMaxBins * ch.nchunks + 1);
  auto kern = rs_downsweep_kernel<K, V1, V2>;
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(RsSmem)));
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
      SB_LAUNCH(kern, ch.nchunks, kRsBlock, sizeof(RsSmem), st, (const K *)src.k, dst.k,
                (const V1 *)src.v1, dst.v1, (const V2 *)src.v2, dst.v2, ch, shift, bits,
                (const int64_t *)spine);
      src = dst;
      shift += bits;
    }
  }
}

}  // namespace sb200
