// One process, R ranks: sb200_mg_comm_create_local + sb200_mg_run_ranks (one host thread per
// rank) drive the peer-memory operators through the C ABI alone -- no torch, no second process.
// The ranks are dealt out round-robin over the visible GPUs; on a single GPU all ranks share
// it (every rank on its own non-blocking stream, windows in the same HBM), which still runs the
// real exchange kernels and barriers.  Every result is compared with the single-GPU operator.
//
//   mg_local_test [ranks=2] [log2_vertices=14]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sb200.h"

#define CK(expr)                                                                       \
  do {                                                                                 \
    const int rc__ = (expr);                                                           \
    if (rc__ != 0) {                                                                   \
      fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #expr, rc__,       \
              sb200_last_error());                                                     \
      return rc__ ? rc__ : -1;                                                         \
    }                                                                                  \
  } while (0)
#define CU(expr)                                                                       \
  do {                                                                                 \
    const cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                          \
      fprintf(stderr, "%s:%d: %s -> %s\n", __FILE__, __LINE__, #expr,                  \
              cudaGetErrorString(e__));                                                \
      return -2;                                                                       \
    }                                                                                  \
  } while (0)
#define EXPECT(cond, what)                                                             \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      fprintf(stderr, "rank %d: MISMATCH %s (%s:%d)\n", rank, what, __FILE__, __LINE__); \
      return -3;                                                                       \
    }                                                                                  \
  } while (0)

template <typename T>
static T *up(const T *h, size_t count) {
  T *d = nullptr;
  cudaMalloc(&d, (count ? count : 1) * sizeof(T));
  if (count) cudaMemcpy(d, h, count * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}
template <typename T>
static std::vector<T> down(const T *d, size_t count) {
  std::vector<T> h(count);
  if (count) cudaMemcpy(h.data(), d, count * sizeof(T), cudaMemcpyDeviceToHost);
  return h;
}

struct Shared {
  int ranks;
  std::vector<int> devices;
  int64_t n, nnz;
  std::vector<int32_t> row, col, x;  // (row, col)-sorted unique COO, x = a vector to permute
  std::vector<float> val;
  std::vector<int64_t> bounds;  // nnz-balanced row blocks
  std::vector<int64_t> first;   // first entry of every block
  // single-GPU results
  std::vector<int32_t> e_rp, e_col, e_inv, p_rp, p_col, c_cp, c_row, e_x;
  std::vector<float> e_val, p_val, c_val;
  sb200_mg_comm_t *comm[16];
  // per-rank device buffers, allocated before and freed after the collective section: a
  // cudaMalloc / cudaFree in one rank's thread synchronises the whole device, and with several
  // ranks on ONE GPU that waits for the other rank's barrier kernel, which waits for this rank
  struct Buf {
    int32_t *d_row, *d_col, *rp, *oc, *inv, *q_ptr, *q_idx, *d_x, *d_o;
    float *d_val, *ov, *q_val;
  } buf[16];
};

static int rank_body(int rank, void *user) {
  Shared &S = *static_cast<Shared *>(user);
  const int dev = S.devices[rank], R = S.ranks;
  CU(cudaSetDevice(dev));
  cudaStream_t st;
  CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  sb200_mg_comm_t *c = S.comm[rank];
  const int64_t lo = S.bounds[rank], nl = S.bounds[rank + 1] - lo;
  const int64_t a = S.first[rank], nz = S.first[rank + 1] - a;
  Shared::Buf &B = S.buf[rank];
  int32_t *d_row = B.d_row, *d_col = B.d_col, *rp = B.rp, *oc = B.oc;
  float *d_val = B.d_val, *ov = B.ov;
  CU(cudaMemcpyAsync(d_row, S.row.data() + a, nz * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_col, S.col.data() + a, nz * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_val, S.val.data() + a, nz * 4, cudaMemcpyHostToDevice, st));
  int64_t out2[3] = {0, 0, 0};
  // ---- COO -> CSR of the block
  CK(sb200_mg_coo_to_csr(c, lo, nl, S.n, nz, d_row, d_col, d_val, rp, oc, ov, out2, SB200_I32,
                         SB200_I32, SB200_F32, st));
  CU(cudaStreamSynchronize(st));
  EXPECT(out2[0] == S.nnz && out2[1] == a, "coo_to_csr totals");
  {
    auto h = down(rp, nl + 1);
    for (int64_t i = 0; i <= nl; i++) EXPECT(h[i] == S.e_rp[lo + i] - S.e_rp[lo], "coo_to_csr row_ptr");
    auto hc = down(oc, nz);
    auto hv = down(ov, nz);
    EXPECT(std::equal(hc.begin(), hc.end(), S.e_col.begin() + a), "coo_to_csr col");
    EXPECT(memcmp(hv.data(), S.e_val.data() + a, nz * sizeof(float)) == 0, "coo_to_csr vals");
  }
  // ---- DegreeReorder: the full permutation on every rank
  int32_t *inv = B.inv;
  CK(sb200_mg_degree_reorder(c, S.n, S.bounds.data(), rp, 1, inv, SB200_I32, SB200_I32, st));
  CU(cudaStreamSynchronize(st));
  EXPECT(down(inv, S.n) == S.e_inv, "degree_reorder");
  // ---- Permute2D: this rank's block of the new rows
  std::vector<int64_t> nb(R + 1);
  CK(sb200_mg_permute2d_run(c, S.n, S.n, S.nnz, S.bounds.data(), rp, oc, ov, inv, inv, nb.data(),
                            out2, SB200_I32, SB200_I32, SB200_F32, st));
  {
    const int64_t rows = out2[0], pn = out2[1], before = out2[2];
    int32_t *q_rp = B.q_ptr, *q_col = B.q_idx;
    float *q_val = B.q_val;
    CK(sb200_mg_permute2d_fetch(c, S.n, rows, pn, q_rp, q_col, q_val, SB200_I32, SB200_I32,
                                SB200_F32, st));
    CU(cudaStreamSynchronize(st));
    EXPECT(rows == nb[rank + 1] - nb[rank] && before == S.p_rp[nb[rank]], "permute2d block");
    auto h = down(q_rp, rows + 1);
    for (int64_t i = 0; i <= rows; i++) EXPECT(h[i] == S.p_rp[nb[rank] + i] - before, "permute2d row_ptr");
    auto hc = down(q_col, pn);
    auto hv = down(q_val, pn);
    EXPECT(std::equal(hc.begin(), hc.end(), S.p_col.begin() + before), "permute2d col");
    EXPECT(memcmp(hv.data(), S.p_val.data() + before, pn * sizeof(float)) == 0, "permute2d vals");
  }
  // ---- CSR -> CSC: this rank's block of the columns
  std::vector<int64_t> cb(R + 1);
  CK(sb200_mg_csr_to_csc_run(c, S.n, S.n, S.nnz, S.bounds.data(), rp, oc, ov, cb.data(), out2,
                             SB200_I32, SB200_I32, SB200_F32, st));
  {
    const int64_t cols = out2[0], cn = out2[1], before = out2[2];
    int32_t *q_cp = B.q_ptr, *q_row = B.q_idx;
    float *q_val = B.q_val;
    CK(sb200_mg_csr_to_csc_fetch(c, S.n, cb[rank], cols, q_cp, q_row, q_val, SB200_I32, SB200_I32,
                                 SB200_F32, st));
    CU(cudaStreamSynchronize(st));
    EXPECT(before == S.c_cp[cb[rank]], "csr_to_csc block");
    auto h = down(q_cp, cols + 1);
    for (int64_t i = 0; i <= cols; i++) EXPECT(h[i] == S.c_cp[cb[rank] + i] - before, "csr_to_csc col_ptr");
    auto hr = down(q_row, cn);
    auto hv = down(q_val, cn);
    EXPECT(std::equal(hr.begin(), hr.end(), S.c_row.begin() + before), "csr_to_csc row");
    EXPECT(memcmp(hv.data(), S.c_val.data() + before, cn * sizeof(float)) == 0, "csr_to_csc vals");
  }
  // ---- Permute1D on the row blocks
  {
    int32_t *d_x = B.d_x, *d_o = B.d_o;
    CU(cudaMemcpyAsync(d_x, S.x.data() + lo, nl * 4, cudaMemcpyHostToDevice, st));
    CK(sb200_mg_permute1d(c, S.bounds.data(), d_x, inv + lo, d_o, SB200_I32, SB200_I32, st));
    CU(cudaStreamSynchronize(st));
    auto h = down(d_o, nl);
    EXPECT(std::equal(h.begin(), h.end(), S.e_x.begin() + lo), "permute1d");
  }
  CK(sb200_mg_barrier(c, st));
  CU(cudaStreamSynchronize(st));
  cudaStreamDestroy(st);
  return 0;
}

int main(int argc, char **argv) {
  Shared S;
  S.ranks = argc > 1 ? atoi(argv[1]) : 2;
  const int lg = argc > 2 ? atoi(argv[2]) : 14;
  int ngpu = 0;
  if (cudaGetDeviceCount(&ngpu) != cudaSuccess || ngpu < 1) {
    fprintf(stderr, "no CUDA device\n");
    return 2;
  }
  if (S.ranks < 1 || S.ranks > 16) return 2;
  for (int r = 0; r < S.ranks; r++) S.devices.push_back(r % ngpu);
  // ---- a power-law-ish symmetric graph: (row, col)-sorted, unique, no self loops
  S.n = (int64_t)1 << lg;
  std::vector<uint64_t> keys;
  uint64_t s = 88172645463325252ull;
  auto rnd = [&] {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    return s;
  };
  for (int64_t e = 0; e < S.n * 8; e++) {
    uint64_t u = rnd() % S.n, v = rnd() % S.n;
    if (e % 3 == 0) u = u % 37;  // a few heavy rows next to each other
    if (u == v) continue;
    keys.push_back(u * S.n + v);
    keys.push_back(v * S.n + u);
  }
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  S.nnz = (int64_t)keys.size();
  for (uint64_t k : keys) {
    S.row.push_back((int32_t)(k / S.n));
    S.col.push_back((int32_t)(k % S.n));
    S.val.push_back((float)((k * 2654435761ull) % 1000) - 500.0f);
  }
  for (int64_t i = 0; i < S.n; i++) S.x.push_back((int32_t)((i * 7919) % 100003));
  // ---- single-GPU results on device 0
  CU(cudaSetDevice(0));
  {
    int32_t *d_row = up(S.row.data(), S.nnz), *d_col = up(S.col.data(), S.nnz);
    float *d_val = up(S.val.data(), S.nnz);
    int32_t *rp, *oc, *inv, *p_rp, *p_col, *cp, *c_row, *d_x = up(S.x.data(), S.n), *o_x;
    float *ov, *p_val, *c_val;
    cudaMalloc(&rp, (S.n + 1) * 4), cudaMalloc(&oc, S.nnz * 4), cudaMalloc(&ov, S.nnz * 4);
    cudaMalloc(&inv, S.n * 4), cudaMalloc(&p_rp, (S.n + 1) * 4), cudaMalloc(&p_col, S.nnz * 4);
    cudaMalloc(&p_val, S.nnz * 4), cudaMalloc(&cp, (S.n + 1) * 4), cudaMalloc(&c_row, S.nnz * 4);
    cudaMalloc(&c_val, S.nnz * 4), cudaMalloc(&o_x, S.n * 4);
    CK(sb200_coo_to_csr(0, S.n, S.n, S.nnz, d_row, d_col, d_val, rp, oc, ov, SB200_I32, SB200_I32,
                        SB200_F32, nullptr));
    CK(sb200_degree_reorder(0, S.n, rp, 1, inv, SB200_I32, SB200_I32, nullptr));
    CK(sb200_permute2d(0, S.n, S.n, S.nnz, rp, oc, ov, inv, inv, p_rp, p_col, p_val, SB200_I32,
                       SB200_I32, SB200_F32, nullptr));
    CK(sb200_csr_to_csc(0, S.n, S.n, S.nnz, rp, oc, ov, cp, c_row, c_val, SB200_I32, SB200_I32,
                        SB200_F32, nullptr));
    CK(sb200_permute1d(0, S.n, d_x, inv, o_x, SB200_I32, SB200_I32, nullptr));
    CU(cudaDeviceSynchronize());
    S.e_rp = down(rp, S.n + 1), S.e_col = down(oc, S.nnz), S.e_val = down(ov, S.nnz);
    S.e_inv = down(inv, S.n), S.p_rp = down(p_rp, S.n + 1), S.p_col = down(p_col, S.nnz);
    S.p_val = down(p_val, S.nnz), S.c_cp = down(cp, S.n + 1), S.c_row = down(c_row, S.nnz);
    S.c_val = down(c_val, S.nnz), S.e_x = down(o_x, S.n);
    S.bounds.resize(S.ranks + 1);
    CK(sb200_partition_rows(0, S.n, S.nnz, rp, SB200_I32, S.ranks, S.bounds.data(), nullptr));
    for (int r = 0; r <= S.ranks; r++) S.first.push_back(S.e_rp[S.bounds[r]]);
    cudaFree(d_row), cudaFree(d_col), cudaFree(d_val), cudaFree(rp), cudaFree(oc), cudaFree(ov);
    cudaFree(inv), cudaFree(p_rp), cudaFree(p_col), cudaFree(p_val), cudaFree(cp), cudaFree(c_row);
    cudaFree(c_val), cudaFree(d_x), cudaFree(o_x);
  }
  // ---- the ranks
  const size_t window = (size_t)S.nnz * 8 * 3 + (size_t)S.n * 64 + ((size_t)64 << 20);
  CK(sb200_mg_comm_create_local(S.ranks, S.devices.data(), window, S.comm));
  int rk = -1, wd = -1;
  size_t wb = 0;
  CK(sb200_mg_comm_info(S.comm[S.ranks - 1], &rk, &wd, &wb));
  if (rk != S.ranks - 1 || wd != S.ranks || wb != window) {
    fprintf(stderr, "comm_info mismatch\n");
    return 3;
  }
  for (int r = 0; r < S.ranks; r++) {
    CU(cudaSetDevice(S.devices[r]));
    Shared::Buf &B = S.buf[r];
    const size_t e = (size_t)S.nnz * 4, v = (size_t)(S.n + 1) * 4;
    void **ents[] = {(void **)&B.d_row, (void **)&B.d_col, (void **)&B.d_val, (void **)&B.oc,
                     (void **)&B.ov, (void **)&B.q_idx, (void **)&B.q_val};
    void **vecs[] = {(void **)&B.rp, (void **)&B.inv, (void **)&B.q_ptr, (void **)&B.d_x,
                     (void **)&B.d_o};
    for (void **q : ents) CU(cudaMalloc(q, e));
    for (void **q : vecs) CU(cudaMalloc(q, v));
  }
  const int rc = sb200_mg_run_ranks(S.ranks, rank_body, &S);
  for (int r = 0; r < S.ranks; r++) {
    cudaSetDevice(S.devices[r]);
    cudaDeviceSynchronize();
    Shared::Buf &B = S.buf[r];
    cudaFree(B.d_row), cudaFree(B.d_col), cudaFree(B.d_val), cudaFree(B.rp), cudaFree(B.oc);
    cudaFree(B.ov), cudaFree(B.inv), cudaFree(B.q_ptr), cudaFree(B.q_idx), cudaFree(B.q_val);
    cudaFree(B.d_x), cudaFree(B.d_o);
  }
  for (int r = 0; r < S.ranks; r++) sb200_mg_comm_destroy(S.comm[r]);
  if (rc != 0) {
    fprintf(stderr, "FAILED rc=%d\n", rc);
    return 1;
  }
  printf("MG LOCAL OK ranks=%d gpus=%d n=%lld nnz=%lld\n", S.ranks, ngpu, (long long)S.n,
         (long long)S.nnz);
  return 0;
}
