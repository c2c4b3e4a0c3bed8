// scan.cuh -- single-pass exclusive prefix sum with decoupled look-back.
//
// Produces out[0..n] (n+1 entries, out[n] = total) from a per-element input functor, i.e. the
// row_ptr / col_ptr arrays of the reference (the hist -> inclusive scan -> shift sequence of
// converter/converter_order_two.cc:185-192 and the nxadj prefix of
// permute/permute_order_two.cc:66 collapse into one exclusive scan).
//
// HBM traffic: one read of the input + one write of the output (n*(sizeof in + sizeof T)); the
// tile status words (8 B per 4096 elements) are the only extra traffic.
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr int kScanBlock = 256;
constexpr int kScanIpt = 16;
constexpr int kScanTile = kScanBlock * kScanIpt;

constexpr uint64_t kScanFlagAgg = 1ull << 62;
constexpr uint64_t kScanFlagPrefix = 2ull << 62;
constexpr uint64_t kScanValueMask = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
  return *reinterpret_cast<const volatile uint64_t *>(p);
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v) {
  *reinterpret_cast<volatile uint64_t *>(p) = v;
}

// status[ntiles] and counter[1] must be zero on entry.
template <typename T, typename InFn>
__global__ void __launch_bounds__(kScanBlock)
    scan_lookback_kernel(InFn in, T *__restrict__ out, int64_t n, uint64_t *status,
                         unsigned *counter) {
  __shared__ T s_warp[34];
  __shared__ unsigned s_tile;
  __shared__ T s_prefix;

  if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u);  // tiles start in index order
  __syncthreads();
  const int64_t tile = s_tile;
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const int64_t warp_base = tile * kScanTile + (int64_t)wid * (kScanIpt * 32);

  T v[kScanIpt], inc[kScanIpt];
#pragma unroll
  for (int r = 0; r < kScanIpt; r++) {
    int64_t i = warp_base + r * 32 + lane;
    v[r] = i < n ? in(i) : T(0);
  }
  T carry = 0;
#pragma unroll
  for (int r = 0; r < kScanIpt; r++) {
    T s = warp_inclusive_scan(v[r]);
    inc[r] = carry + s;
    carry += __shfl_sync(0xffffffffu, s, 31);
  }
  // carry == this warp's total (same in all lanes)
  T warp_excl;
  {
    if (lane == 0) s_warp[wid] = carry;
    __syncthreads();
    if (wid == 0) {
      T s = lane < (kScanBlock / 32) ? s_warp[lane] : T(0);
      T si = warp_inclusive_scan(s);
      T agg = __shfl_sync(0xffffffffu, si, 31);
      // ---- decoupled look-back (warp 0) ----
      T excl = 0;
      if (tile == 0) {
        if (lane == 0) st_volatile_u64(&status[0], kScanFlagPrefix | (uint64_t)agg);
      } else {
        if (lane == 0) st_volatile_u64(&status[tile], kScanFlagAgg | (uint64_t)agg);
        int64_t look = tile - 1;
        while (true) {
          int64_t idx = look - lane;
          uint64_t st = idx >= 0 ? ld_volatile_u64(&status[idx]) : kScanFlagPrefix;
          while (__any_sync(0xffffffffu, (st >> 62) == 0)) {
            if ((st >> 62) == 0) st = ld_volatile_u64(&status[idx]);
          }
          unsigned pm = __ballot_sync(0xffffffffu, (st >> 62) == 2);
          T val = (T)(st & kScanValueMask);
          if (pm) {
            int first = __ffs(pm) - 1;
            excl += warp_reduce_sum((int)lane <= first ? val : T(0));
            break;
          }
          excl += warp_reduce_sum(val);
          look -= 32;
        }
        if (lane == 0)
          st_volatile_u64(&status[tile], kScanFlagPrefix | (uint64_t)(excl + agg));
      }
      if (lane == 0) s_prefix = excl;
      s_warp[lane] = si - s;  // exclusive warp offsets
    }
    __syncthreads();
    warp_excl = s_warp[wid] + s_prefix;
  }
#pragma unroll
  for (int r = 0; r < kScanIpt; r++) {
    int64_t i = warp_base + r * 32 + lane;
    if (i < n) {
      out[i] = warp_excl + inc[r] - v[r];
      if (i == n - 1) out[n] = warp_excl + inc[r];
    }
  }
}

// Exclusive scan of in(0..n-1) into out[0..n] (n+1 entries).
template <typename T, typename InFn>
void exclusive_scan(Workspace &ws, InFn in, T *out, int64_t n) {
  cudaStream_t st = ws.stream();
  if (n <= 0) {
    SB_CUDA(cudaMemsetAsync(out, 0, sizeof(T), st));
    return;
  }
  int64_t ntiles = ceil_div(n, kScanTile);
  uint64_t *status = ws.alloc<uint64_t>(ntiles + 1);
  SB_CUDA(cudaMemsetAsync(status, 0, (ntiles + 1) * sizeof(uint64_t), st));
  unsigned *counter = reinterpret_cast<unsigned *>(status + ntiles);
  SB_LAUNCH((scan_lookback_kernel<T, InFn>), (unsigned)ntiles, kScanBlock, 0, st, in, out, n,
            status, counter);
}

}  // namespace sb200
