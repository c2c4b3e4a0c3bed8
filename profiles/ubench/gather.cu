// Micro-benchmark: random 4-byte gathers from a table of T MB while a stream of S bytes per gather
// passes through L2 (the Permute2D access mix).  Variants of the gather load instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu && ./gather
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
template <int MODE>
__device__ __forceinline__ unsigned gload(const unsigned *p, uint64_t pol) {
  unsigned v;
  if (MODE == 0) v = *p;
  else if (MODE == 1) v = __ldg(p);
  else if (MODE == 2) asm volatile("ld.global.nc.L2::128B.b32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (MODE == 3) asm volatile("ld.global.nc.L2::256B.b32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (MODE == 4) asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  else if (MODE == 5) asm volatile("ld.global.nc.L2::cache_hint.L2::256B.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  else if (MODE == 6) asm volatile("ld.global.L2::cache_hint.L2::256B.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
// each thread: 4 x (stream-read idx+val, gather table[idx], stream-write)
// ROWS: the two input streams are read as randomly placed 16-element rows (a row gather)
// instead of sequentially.
__device__ bool g_rows = false;
template <int MODE>
__global__ void k(const unsigned *__restrict__ idx, const unsigned *__restrict__ val,
                  const unsigned *__restrict__ table, unsigned *__restrict__ o1,
                  unsigned *__restrict__ o2, int64_t n) {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
  unsigned a[4], b[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    int64_t i = i0 + u * blockDim.x;
    if (i < n) {
      int64_t p = i;
      if (g_rows) p = (int64_t)(((uint64_t)hash32((unsigned)(i >> 4) * 0x9e3779b9u) * (uint64_t)(n >> 4)) >> 32) * 16 + (i & 15);
      a[u] = __ldcs(idx + p); b[u] = __ldcs(val + p);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; u++) { int64_t i = i0 + u * blockDim.x; if (i < n) a[u] = gload<MODE>(table + a[u], pol); }
#pragma unroll
  for (int u = 0; u < 4; u++) { int64_t i = i0 + u * blockDim.x; if (i < n) { __stcs(o1 + i, a[u]); __stcs(o2 + i, b[u]); } }
}
__global__ void fill_idx(unsigned *idx, int64_t n, unsigned tsize, unsigned seed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (unsigned)(((uint64_t)hash32((unsigned)i * 2654435761u + seed) * tsize) >> 32);
}
template <int MODE>
float run(const unsigned *idx, const unsigned *val, const unsigned *table, unsigned *o1, unsigned *o2, int64_t n) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int grid = (int)((n + 1023) / 1024);
  k<MODE><<<grid, 256>>>(idx, val, table, o1, o2, n);
  cudaEventRecord(a);
  for (int r = 0; r < 3; r++) k<MODE><<<grid, 256>>>(idx, val, table, o1, o2, n);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 3;
}
int main() {
  const int64_t n = 1ll << 28;  // gathers per launch
  unsigned *idx, *val, *o1, *o2, *table;
  cudaMalloc(&idx, n * 4); cudaMalloc(&val, n * 4); cudaMalloc(&o1, n * 4); cudaMalloc(&o2, n * 4);
  cudaMalloc(&table, 1ll << 30);
  cudaMemset(table, 0, 1ll << 30); cudaMemset(val, 0, n * 4);
  size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("persisting L2 limit %zu, max %d, l2 %d\n", lim, p.persistingL2CacheMaxSize, p.l2CacheSize);
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) { bool t = true; cudaMemcpyToSymbol(g_rows, &t, sizeof(bool)); printf("-- input streams read as random 64-byte rows --\n"); }
    for (int mb : {16, 32, 64, 128, 256}) {
      unsigned tsize = (unsigned)((int64_t)mb << 18);
      fill_idx<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, tsize, 7u);
      cudaDeviceSynchronize();
      float t[7];
      t[0] = run<0>(idx, val, table, o1, o2, n); t[1] = run<1>(idx, val, table, o1, o2, n);
      t[2] = run<2>(idx, val, table, o1, o2, n); t[3] = run<3>(idx, val, table, o1, o2, n);
      t[4] = run<4>(idx, val, table, o1, o2, n); t[5] = run<5>(idx, val, table, o1, o2, n);
      t[6] = run<6>(idx, val, table, o1, o2, n);
      printf("table %3d MB: ld %.3f  ldg %.3f  nc.128B %.3f  nc.256B %.3f  nc.keep %.3f  nc.keep.256B %.3f  keep.256B %.3f ms  (%.1f Gg/s best)\n", mb,
             t[0], t[1], t[2], t[3], t[4], t[5], t[6], n / 1e6 / fminf(fminf(fminf(t[0], t[1]), fminf(t[2], t[3])), fminf(fminf(t[4], t[5]), t[6])));
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
