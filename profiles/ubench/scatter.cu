// Micro-benchmark: the write pattern of a radix-partition pass at different fan-outs.
// Each CTA (one per SM, or two) walks its chunk of the input in tiles of T records; per tile it
// reads 12 bytes per record as a stream and writes the tile's records to F bins, T/F consecutive
// records per bin at the chunk's running cursor of that bin (the chunk-major layout of the
// reduce-then-scan radix sort), three 4-byte arrays.  Question: how far can F grow before the
// 126 MB L2 stops merging the short runs into full sectors (DRAM write amplification)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter scatter.cu && ./scatter
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int T>
__global__ void __launch_bounds__(512)
    k(const unsigned *__restrict__ a, const unsigned *__restrict__ b,
      const unsigned *__restrict__ c, unsigned *__restrict__ oa, unsigned *__restrict__ ob,
      unsigned *__restrict__ oc, int64_t n, int F, int shuffle) {
  // chunk = contiguous range of tiles; bin f of chunk k owns [f * n/F + k * per, ...) of the output
  const int64_t tiles = n / T, nchunks = gridDim.x;
  const int64_t t0 = tiles * blockIdx.x / nchunks, t1 = tiles * (blockIdx.x + 1) / nchunks;
  const int64_t bin_len = n / F;                  // records per bin overall
  const int64_t per = bin_len / nchunks;          // this chunk's share of a bin
  const int run = T / F > 0 ? T / F : 1;          // records per bin and tile
  for (int64_t t = t0; t < t1; t++) {
    const int64_t done = (t - t0) * run;          // cursor advance so far
    for (int j = threadIdx.x; j < T; j += 512) {
      const int64_t i = t * T + j;
      const unsigned x = __ldcs(a + i), y = __ldcs(b + i), z = __ldcs(c + i);
      // record j of the tile goes to bin f, position r of the run
      int f = j / run, r = j % run;
      if (shuffle) f = (int)(((unsigned)f * 2654435761u) % (unsigned)F);  // bins in random order
      if (f >= F) continue;
      const int64_t dst = (int64_t)f * bin_len + blockIdx.x * per + done + r;
      if (done + r < per) {
        oa[dst] = x;
        ob[dst] = y;
        oc[dst] = z;
      }
    }
  }
}

int main() {
  const int64_t n = 256ll << 20;  // 268 M records
  unsigned *a, *b, *c, *oa, *ob, *oc;
  cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&c, n * 4);
  cudaMalloc(&oa, n * 4); cudaMalloc(&ob, n * 4); cudaMalloc(&oc, n * 4);
  cudaMemset(a, 1, n * 4); cudaMemset(b, 2, n * 4); cudaMemset(c, 3, n * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148;
  printf("records %lld, 12 B in + 12 B out each; GB/s counts 24 B per record\n", (long long)n);
  for (int ctas : {148, 296}) {
    for (int shuffle : {0, 1}) {
      for (int F : {64, 256, 512, 1024, 2048, 4096, 8192}) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; rep++) {
          cudaEventRecord(e0);
          k<8192><<<ctas, 512>>>(a, b, c, oa, ob, oc, n, F, shuffle);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        printf("ctas %d shuffle %d F %5d run %4d rec: %.3f ms  %.0f GB/s  (%s)\n", ctas, shuffle, F,
               8192 / F > 0 ? 8192 / F : 1, best, n * 24.0 / best / 1e6,
               cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  (void)sms;
  return 0;
}
