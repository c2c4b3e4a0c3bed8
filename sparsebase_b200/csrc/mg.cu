// mg.cu -- multi-GPU (row-block sharded) operators over peer memory (see mg.cuh).
//
// The reference has no distributed code (its whole "multi-GPU" is the peer
// copy of converter/converter_order_two_cuda.cu:41-76); these are the sharded forms of the same
// operators: every rank owns a contiguous row block [row_lo, row_lo +
// n_local) with block-local row_ptr and global column ids, results are bit-identical to the
// single-GPU operators.  SPMD: every rank calls the same entry point with its own shard
// (one process per GPU with IPC-connected communicators, or one host thread per GPU inside one
// process, sb200_mg_comm_create_local).
//
//   sb200_mg_coo_to_csr      local COO->CSR of the block + all-gather of the block sizes
//   sb200_mg_degree_reorder  local rank by degree + per-degree offsets from the gathered degree
//                            histograms; the full permutation lands on every rank
//   sb200_mg_permute2d       degrees all-gathered -> new row_ptr (replicated scan) -> every rank
//                            renumbers / sorts its rows and STORES each into the window of the
//                            rank that owns the new row, at its final offset
//   sb200_mg_csr_to_csc      local transpose of the block, column counts reduced through the
//                            windows, every source stores its column slices into the owner's
//                            window, the owner interleaves them (source order = row order)
//   sb200_mg_permute1d       out[order[i]] = vals[i] stored straight into the owner's window
#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include "mg.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

// ------------------------------------------------------------------ barrier / small gathers
// A peer that failed (or never made the matching call) must not hang this GPU: after
// kMgBarrierTimeoutNs the barrier gives up and raises pad[0] of this rank's control area, which
// the operators turn into SB200_ERR_INTERNAL at their next synchronisation point.
constexpr unsigned long long kMgBarrierTimeoutNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ unsigned long long mg_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void mg_barrier_kernel(MgPeers p, unsigned long long epoch) {
  const int r = threadIdx.x;
  if (r < p.world) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(&p.ctl(r)->flag[p.rank]) = epoch;
    __threadfence_system();
    const volatile unsigned long long *mine =
        reinterpret_cast<const volatile unsigned long long *>(&p.ctl(p.rank)->flag[r]);
    const unsigned long long t0 = mg_now_ns();
    while (*mine < epoch) {
      if (mg_now_ns() - t0 > kMgBarrierTimeoutNs) {
        p.ctl(p.rank)->pad[0] = 1ull;
        break;
      }
    }
    __threadfence_system();
  }
}

// All peer stores issued on `st` before this call are visible to every rank after it (in stream
// order), and every rank has reached this point.
void mg_barrier(sb200_mg_comm *c, cudaStream_t st) {
  c->epoch++;
  if (c->peers.world == 1) return;
  SB_LAUNCH(mg_barrier_kernel, 1, 32, 0, st, c->peers, c->epoch);
}

// table[my rank][slot0 + k] = vals[k] on every rank (k < count)
__global__ void mg_put_table_kernel(MgPeers p, int slot0, const unsigned long long *__restrict__ vals,
                                    int count) {
  const int r = blockIdx.x, k = threadIdx.x;
  if (r < p.world && k < count) p.ctl(r)->table[p.rank][slot0 + k] = vals[k];
}
void mg_put_table(sb200_mg_comm *c, cudaStream_t st, int slot0, const unsigned long long *d_vals,
                  int count) {
  SB_REQUIRE(slot0 >= 0 && slot0 + count <= kMgSlots && count <= 64, SB200_ERR_INTERNAL,
             "table slots out of range");
  SB_LAUNCH(mg_put_table_kernel, c->peers.world, 64, 0, st, c->peers, slot0, d_vals, count);
}

// my slice [elems] of bytes-wide elements -> offset `at` of region `region` in EVERY rank's window
__global__ void __launch_bounds__(256)
    mg_bcast_slice_kernel(MgPeers p, size_t region, const uint4 *__restrict__ src, size_t at_bytes,
                          size_t bytes) {
  // 16-byte words; tails handled by the caller's padding (regions are 256-byte aligned and the
  // slices are padded to 16 bytes by mg_bcast_slice)
  const size_t words = bytes / 16;
  for (int r = blockIdx.y; r < p.world; r += gridDim.y) {
    uint4 *dst = reinterpret_cast<uint4 *>(p.data(r) + region + at_bytes);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words;
         i += (size_t)gridDim.x * blockDim.x)
      dst[i] = src[i];
  }
}
__global__ void mg_bcast_tail_kernel(MgPeers p, size_t region, const char *__restrict__ src,
                                     size_t at_bytes, size_t from, size_t bytes) {
  const size_t i = from + threadIdx.x;
  if (i < bytes)
    for (int r = 0; r < p.world; r++) (p.data(r) + region + at_bytes)[i] = src[i];
}
// Slices of 4- or 8-byte elements start at arbitrary element offsets, so source and destination
// rarely share their 16-byte phase: the destination side is aligned (a short scalar head, then
// 16-byte stores -- NVLink moves full 128-byte lines per warp), the source side is read as
// 4-byte words.
__global__ void __launch_bounds__(256)
    mg_bcast_words_kernel(MgPeers p, size_t region, const unsigned *__restrict__ src,
                          size_t at_bytes, size_t words) {
  const size_t head = ((16 - (at_bytes & 15)) & 15) / 4;  // words before the first aligned store
  const size_t h = head < words ? head : words;
  const size_t quads = (words - h) / 4;
  const size_t tail0 = h + quads * 4;
  for (int r = blockIdx.y; r < p.world; r += gridDim.y) {
    unsigned *dst = reinterpret_cast<unsigned *>(p.data(r) + region + at_bytes);
    uint4 *dq = reinterpret_cast<uint4 *>(dst + h);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads;
         i += (size_t)gridDim.x * blockDim.x) {
      const unsigned *s4 = src + h + i * 4;
      dq[i] = make_uint4(s4[0], s4[1], s4[2], s4[3]);
    }
    if (blockIdx.x == 0) {
      if (threadIdx.x < h) dst[threadIdx.x] = src[threadIdx.x];
      if (tail0 + threadIdx.x < words) dst[tail0 + threadIdx.x] = src[tail0 + threadIdx.x];
    }
  }
}
void mg_bcast_slice(sb200_mg_comm *c, cudaStream_t st, size_t region, const void *src,
                    size_t at_bytes, size_t bytes) {
  if (bytes == 0) return;
  const int sms = device_info(c->device).sm_count;
  dim3 grid((unsigned)(sms * 4 / c->peers.world > 0 ? sms * 4 / c->peers.world : 1),
            (unsigned)c->peers.world);
  const bool aligned = (at_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
  if (aligned && bytes >= 16) {
    SB_LAUNCH(mg_bcast_slice_kernel, grid, 256, 0, st, c->peers, region, (const uint4 *)src,
              at_bytes, bytes);
    const size_t done = bytes / 16 * 16;
    if (done < bytes)
      SB_LAUNCH(mg_bcast_tail_kernel, 1, 16, 0, st, c->peers, region, (const char *)src, at_bytes,
                done, bytes);
  } else if (at_bytes % 4 == 0 && bytes % 4 == 0 && reinterpret_cast<uintptr_t>(src) % 4 == 0) {
    SB_LAUNCH(mg_bcast_words_kernel, grid, 256, 0, st, c->peers, region, (const unsigned *)src,
              at_bytes, bytes / 4);
  } else {
    for (int r = 0; r < c->peers.world; r++)
      SB_CUDA(cudaMemcpyAsync(c->peers.data(r) + region + at_bytes, src, bytes,
                              cudaMemcpyDeviceToDevice, st));
  }
}

// after a stream synchronisation: did a barrier of this rank give up?
static void check_barriers(const sb200_mg_comm *c) {
  if (c->peers.world == 1) return;
  unsigned long long failed = 0;
  SB_CUDA(cudaMemcpy(&failed, &c->peers.ctl(c->peers.rank)->pad[0], sizeof(failed),
                     cudaMemcpyDeviceToHost));
  SB_REQUIRE(failed == 0, SB200_ERR_INTERNAL,
             "multi-GPU barrier timed out: a peer rank failed or did not make the matching call");
}

// SB200_MG_TRACE=1: per-stage device time of the operators, printed by rank 0 (CUDA events on
// the operator's stream; the print happens after the operator's own final synchronisation).
class MgTrace {
 public:
  MgTrace(const sb200_mg_comm *c, cudaStream_t st, const char *op)
      : on_(false), rank_(c->peers.rank), st_(st), op_(op) {
    const char *e = getenv("SB200_MG_TRACE");  // "1": rank 0, "all": every rank
    on_ = e != nullptr && (c->peers.rank == 0 || strcmp(e, "all") == 0);
    t0_ = std::chrono::steady_clock::now();
    mark("start");
  }
  void mark(const char *name) {
    if (!on_) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st_);
    ev_.push_back(e);
    names_.push_back(name);
  }
  ~MgTrace() {
    if (!on_) return;
    cudaStreamSynchronize(st_);
    std::string line = std::string("[mg-trace] r") + std::to_string(rank_) + " " + op_ + ":";
    for (size_t i = 1; i < ev_.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev_[i - 1], ev_[i]);
      char buf[96];
      snprintf(buf, sizeof(buf), "  %s=%.3f", names_[i], ms);
      line += buf;
    }
    char wall[64];
    snprintf(wall, sizeof(wall), "  host_wall=%.3f",
             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_)
                 .count());
    line += wall;
    fprintf(stderr, "%s\n", line.c_str());
    for (cudaEvent_t e : ev_) cudaEventDestroy(e);
  }

 private:
  bool on_;
  int rank_;
  std::chrono::steady_clock::time_point t0_;
  cudaStream_t st_;
  const char *op_;
  std::vector<cudaEvent_t> ev_;
  std::vector<const char *> names_;
};

static void check_comm(const sb200_mg_comm *c) {
  SB_REQUIRE(c != nullptr && c->peers.world >= 1 && c->peers.world <= kMgMaxRanks &&
                 c->peers.rank >= 0 && c->peers.rank < c->peers.world,
             SB200_ERR_BAD_ARG, "bad communicator");
  for (int r = 0; r < c->peers.world; r++)
    SB_REQUIRE(c->peers.win[r] != nullptr, SB200_ERR_BAD_ARG,
               "communicator not connected (peer %d)", r);
}

// second stream (+ an event) of a communicator: exchanges that overlap local work
constexpr int kMgMaxChunks = 16;
static int mg_push_chunks() {
  const char *e = getenv("SB200_MG_CHUNKS");
  int k = e ? atoi(e) : 1;
  return k < 1 ? 1 : (k > kMgMaxChunks ? kMgMaxChunks : k);
}
static bool mg_push_dest_order() {  // SB200_MG_PUSH_ORDER=source: rows pushed as they lie
  const char *e = getenv("SB200_MG_PUSH_ORDER");
  return !(e && e[0] == 's');
}
static int64_t mg_push_chunk_min() {  // entries below which a chunk is not worth its launches
  const char *e = getenv("SB200_MG_CHUNK_MIN");
  const long long v = e ? atoll(e) : (1ll << 20);
  return v < 1 ? 1 : v;
}
static cudaStream_t mg_aux_stream(sb200_mg_comm *c) {
  if (!c->aux_stream) {
    cudaStream_t s;
    SB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    c->aux_stream = s;
  }
  return (cudaStream_t)c->aux_stream;
}
static cudaEvent_t mg_aux_event(sb200_mg_comm *c) {
  if (!c->aux_event) {
    cudaEvent_t e;
    SB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->aux_event = e;
  }
  return (cudaEvent_t)c->aux_event;
}

// ------------------------------------------------------------------ Permute1D
// out[order[i]] = vals[i]: the element goes straight into the window of the rank that owns
// position order[i] (blocks of `bounds`), then every rank copies its block out.
template <typename I, typename V>
__global__ void mg_permute1d_push_kernel(MgPeers p, size_t region, const int64_t *__restrict__ bounds,
                                         const V *__restrict__ vals, const I *__restrict__ order,
                                         int64_t n_local) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  const int64_t j = (int64_t)order[i];
  int d = 0;
  while (d + 1 < p.world && j >= bounds[d + 1]) d++;
  reinterpret_cast<V *>(p.data(d) + region)[j - bounds[d]] = ld_stream(vals + i);
}

// ------------------------------------------------------------------ DegreeReorder
// offset[d] = (vertices of smaller degree anywhere) + (same degree in LATER blocks: larger ids
// rank first, degree_reorder.cc:42-46) - (smaller degree inside this block)
struct MgHistTotalFn {
  const unsigned long long *hist;  // [world][nbins] in my window
  int world;
  int64_t nbins;
  __device__ int64_t operator()(int64_t d) const {
    unsigned long long s = 0;
    for (int r = 0; r < world; r++) s += hist[(int64_t)r * nbins + d];
    return (int64_t)s;
  }
};
struct MgHistMineFn {
  const unsigned long long *hist;
  int rank;
  int64_t nbins;
  __device__ int64_t operator()(int64_t d) const { return (int64_t)hist[(int64_t)rank * nbins + d]; }
};
__global__ void mg_degree_offset_kernel(const unsigned long long *__restrict__ hist, int rank,
                                        int world, int64_t nbins,
                                        const int64_t *__restrict__ g_start,
                                        const int64_t *__restrict__ l_start,
                                        int64_t *__restrict__ offset) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nbins) return;
  unsigned long long later = 0;
  for (int r = rank + 1; r < world; r++) later += hist[(int64_t)r * nbins + d];
  offset[d] = g_start[d] + (int64_t)later - l_start[d];
}

constexpr int64_t kMgChunk = 4096;  // entries per work item of the chunked copies

// ------------------------------------------------------------------ Permute2D
// replicated: new_deg[row_order[i]] = deg[i]
template <typename I, typename N>
__global__ void mg_new_degree_kernel(const N *__restrict__ deg, const I *__restrict__ row_order,
                                     int64_t n, N *__restrict__ new_deg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) new_deg[row_order ? (int64_t)row_order[i] : i] = deg[i];
}
template <typename I, typename UI>
__global__ void mg_order_keys_kernel(const I *__restrict__ new_id, int64_t n_local,
                                     UI *__restrict__ keys, UI *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_local) {
    keys[i] = (UI)new_id[i];
    idx[i] = (UI)i;
  }
}
template <typename I, typename N>
struct MgWalkLenFn {  // length of the k-th row of the walk
  const N *ptr;
  const I *sidx;
  __device__ N operator()(int64_t k) const {
    const int64_t i = (int64_t)sidx[k];
    return ptr[i + 1] - ptr[i];
  }
};
template <typename N>
__global__ void mg_rebase_ptr_kernel(const N *__restrict__ ptr, int64_t count, N base,
                                     N *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = ptr[i] - base;
}
// bounds[k] = first new row whose new_ptr >= k * nnz / world (as sb200_partition_rows)
template <typename N>
__global__ void mg_balance_kernel(const N *__restrict__ ptr, int64_t n, int64_t nnz, int parts,
                                  int64_t *__restrict__ bounds, int64_t *__restrict__ at) {
  const int k = threadIdx.x;
  if (k > parts) return;
  int64_t b;
  if (k == 0)
    b = 0;
  else if (k == parts)
    b = n;
  else {
    const int64_t target = (int64_t)(((__int128)nnz * k) / parts);
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)ptr[mid] < target)
        lo = mid + 1;
      else
        hi = mid;
    }
    b = lo;
  }
  bounds[k] = b;
  at[k] = (int64_t)ptr[b];
}
// The rows of this rank (already renumbered and sorted in `tcol` / `tval`, laid out by the
// block-local row_ptr) go into the windows of the ranks that own their new row ids.
// (Tried: local rows sorted by new id first, so that the remote stores become long runs -- the
// push of the low-degree block of R-MAT-26 went from 12.0 to 5.8 ms at 4 GPUs, but the local
// gather of the now scattered 2-entry rows cost 8 ms more than that.)
// Balanced over ENTRIES, not rows (the heavy rows of a power-law matrix sit next to each other:
// with one warp per 32 rows ncu showed 11 % of the warps active): every warp takes kMgPushChunk
// consecutive entries, finds the row holding the first one by binary search, and walks the rows
// from there -- lane l resolves row l's destination (owner rank + final offset, clipped to the
// chunk), then the warp moves the rows' concatenated entries 32 at a time: coalesced reads, and
// the lanes of one row store a contiguous run.  A hub row simply spans many chunks.
constexpr int64_t kMgPushChunk = 4096;
template <typename I, typename N, typename V>
__global__ void __launch_bounds__(256)
    mg_push_rows_kernel(MgPeers p, size_t col_region, size_t val_region,
                        const N *__restrict__ row_ptr, const I *__restrict__ tcol,
                        const V *__restrict__ tval, const I *__restrict__ row_order, int64_t row_lo,
                        int64_t n_local, int64_t nnz_local, const N *__restrict__ new_ptr,
                        const int64_t *__restrict__ nb, const int64_t *__restrict__ nb_at,
                        const I *__restrict__ sidx, const I *__restrict__ dest_id,
                        const N *__restrict__ dptr) {
  // walk order: the local rows as they lie (sidx == null), or sorted by new id: the k-th row of
  // the walk is local row sidx[k], its new id dest_id[k], and dptr is the row_ptr of that order
  const N *__restrict__ walk = sidx ? dptr : row_ptr;
  const unsigned lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nchunks = (nnz_local + kMgPushChunk - 1) / kMgPushChunk;
  for (int64_t w = warp; w < nchunks; w += nwarps) {
    const int64_t e0 = w * kMgPushChunk;
    const int64_t e1 = e0 + kMgPushChunk < nnz_local ? e0 + kMgPushChunk : nnz_local;
    int64_t lo = 0, hi = n_local;  // last row with row_ptr[row] <= e0
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)walk[mid] <= e0)
        lo = mid;
      else
        hi = mid;
    }
    for (int64_t g = lo; g < n_local; g += 32) {
      const int64_t i = g + lane;
      int64_t b = 0;  // first entry of the row inside the chunk, as a position in tcol / tval
      unsigned d = 0;
      I *oc = nullptr;
      V *ov = nullptr;
      int64_t row_begin = nnz_local;  // (rows past the end start beyond every chunk)
      if (i < n_local) {
        row_begin = (int64_t)walk[i];
        const int64_t row_end = (int64_t)walk[i + 1];
        b = row_begin > e0 ? row_begin : e0;
        const int64_t e = row_end < e1 ? row_end : e1;
        if (e > b) {
          // destination: the owner of the new row id, at the row's final offset in its block
          const int64_t j = sidx ? (int64_t)dest_id[i]
                                 : (row_order ? (int64_t)row_order[row_lo + i] : row_lo + i);
          int r = 0;
          while (r + 1 < p.world && j >= nb[r + 1]) r++;
          const int64_t dst = (int64_t)new_ptr[j] - nb_at[r] + (b - row_begin);
          oc = reinterpret_cast<I *>(p.data(r) + col_region) + dst;
          if constexpr (has_val<V>) ov = reinterpret_cast<V *>(p.data(r) + val_region) + dst;
          d = (unsigned)(e - b);
          if (sidx) b += (int64_t)row_ptr[sidx[i]] - row_begin;  // where the row lies locally
        }
      }
      // long runs first: the whole warp copies them with four independent loads per lane in
      // flight (a hub row spans many chunks; its pieces are 4096-entry runs)
      unsigned longs = __ballot_sync(0xffffffffu, d >= 128u);
      while (longs) {
        const int src = __ffs(longs) - 1;
        longs &= longs - 1;
        const unsigned len = __shfl_sync(0xffffffffu, d, src);
        const int64_t from = __shfl_sync(0xffffffffu, b, src);
        I *dc = reinterpret_cast<I *>(__shfl_sync(0xffffffffu, (unsigned long long)oc, src));
        [[maybe_unused]] V *dv =
            reinterpret_cast<V *>(__shfl_sync(0xffffffffu, (unsigned long long)ov, src));
        for (unsigned k0 = 0; k0 < len; k0 += 128) {
          I cbuf[4];
          [[maybe_unused]] V vbuf[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const unsigned k = k0 + u * 32 + lane;
            if (k < len) {
              cbuf[u] = ld_stream(tcol + from + k);
              if constexpr (has_val<V>) {
                if (tval) vbuf[u] = ld_stream(tval + from + k);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const unsigned k = k0 + u * 32 + lane;
            if (k < len) {
              dc[k] = cbuf[u];
              if constexpr (has_val<V>) {
                if (tval) dv[k] = vbuf[u];
              }
            }
          }
        }
        if ((int)lane == src) d = 0;
      }
      const unsigned incl = warp_inclusive_scan(d);
      const unsigned excl = incl - d;
      const unsigned tot = __shfl_sync(0xffffffffu, incl, 31);
      for (unsigned base = 0; base < tot; base += 32) {
        const unsigned s = base + lane;
        unsigned own = 0;  // number of lanes whose inclusive end <= s == owner lane
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const unsigned val = __shfl_sync(0xffffffffu, incl, (own + step - 1) & 31);
          if (val <= s) own += step;
        }
        const unsigned j = own & 31;
        const int64_t b_j = __shfl_sync(0xffffffffu, b, j);
        const unsigned excl_j = __shfl_sync(0xffffffffu, excl, j);
        const unsigned long long oc_j = __shfl_sync(0xffffffffu, (unsigned long long)oc, j);
        const unsigned long long ov_j = __shfl_sync(0xffffffffu, (unsigned long long)ov, j);
        if (s < tot) {
          const unsigned k = s - excl_j;
          reinterpret_cast<I *>(oc_j)[k] = ld_stream(tcol + b_j + k);
          if constexpr (has_val<V>) {
            if (tval) reinterpret_cast<V *>(ov_j)[k] = ld_stream(tval + b_j + k);
          }
        }
      }
      // the next 32 rows start at or beyond the end of the chunk: done
      if (__shfl_sync(0xffffffffu, i + 1 < n_local ? (int64_t)walk[i + 1] : nnz_local, 31) >= e1)
        break;
    }
  }
}

// ------------------------------------------------------------------ CSR -> CSC
// cp[s] = source s's block-local col_ptr over ALL m columns (the CP table in my window)
template <typename N>
struct MgColTotalFn {
  const N *cp;  // [world][m + 1]
  int world;
  int64_t m;
  __device__ N operator()(int64_t c) const {
    N s = 0;
    for (int r = 0; r < world; r++) s += cp[(int64_t)r * (m + 1) + c + 1] - cp[(int64_t)r * (m + 1) + c];
    return s;
  }
};
// Source `rank` stores its slice of columns [cb[d], cb[d+1]) -- rows and values are contiguous in
// its block-local CSC -- into destination d's staging area, behind the slices of the lower ranks.
template <typename I, typename N, typename V>
__global__ void __launch_bounds__(256)
    mg_push_cols_kernel(MgPeers p, size_t row_region, size_t val_region, const N *__restrict__ cp,
                        int64_t m, const int64_t *__restrict__ cb, const I *__restrict__ rows,
                        const V *__restrict__ vals) {
  const int d = blockIdx.y;
  const int64_t c0 = cb[d], c1 = cb[d + 1];
  const N *mine = cp + (int64_t)p.rank * (m + 1);
  const int64_t from = (int64_t)mine[c0], cnt = (int64_t)mine[c1] - from;
  int64_t off = 0;
  for (int r = 0; r < p.rank; r++)
    off += (int64_t)(cp[(int64_t)r * (m + 1) + c1] - cp[(int64_t)r * (m + 1) + c0]);
  I *orow = reinterpret_cast<I *>(p.data(d) + row_region) + off;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cnt;
       k += (int64_t)gridDim.x * blockDim.x)
    orow[k] = ld_stream(rows + from + k);
  if constexpr (has_val<V>) {
    if (vals) {
      V *oval = reinterpret_cast<V *>(p.data(d) + val_region) + off;
      for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cnt;
           k += (int64_t)gridDim.x * blockDim.x)
        oval[k] = ld_stream(vals + from + k);
    }
  }
}
// The owner of columns [c0, c1) interleaves the staged slices: inside a column the sources come
// in rank order, which is ascending row order because the row blocks are ordered.  One lane per
// column (adjacent lanes read adjacent places of every source's slice and write adjacent
// segments); columns of kMgLongCol entries or more go to a list and mg_interleave_long_kernel
// copies them in chunks dealt out over the whole grid -- a power-law matrix has columns of
// millions of entries, and its heavy columns sit next to each other.
constexpr int64_t kMgLongCol = 64;
template <typename I, typename N, typename V>
__global__ void __launch_bounds__(256)
    mg_interleave_kernel(int world, const N *__restrict__ cp, int64_t m, const N *__restrict__ gptr,
                         int64_t c0, int64_t c1, const I *__restrict__ stg_row,
                         const V *__restrict__ stg_val, N *__restrict__ out_col_ptr,
                         I *__restrict__ out_row, V *__restrict__ out_val,
                         int64_t *__restrict__ long_list, unsigned *__restrict__ long_count) {
  const int64_t g0 = (int64_t)gptr[c0];
  for (int64_t c = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < c1;
       c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t total = (int64_t)gptr[c + 1] - (int64_t)gptr[c];
    out_col_ptr[c - c0] = (N)((int64_t)gptr[c] - g0);
    if (c == c1 - 1) out_col_ptr[c1 - c0] = (N)((int64_t)gptr[c1] - g0);
    if (total >= kMgLongCol) {
      long_list[atomicAdd(long_count, 1u)] = c;
      continue;
    }
    int64_t dst = (int64_t)gptr[c] - g0, stage = 0;
    for (int r = 0; r < world; r++) {
      const N *q = cp + (int64_t)r * (m + 1);
      const int64_t src = stage + ((int64_t)q[c] - (int64_t)q[c0]);
      const int64_t len = (int64_t)q[c + 1] - (int64_t)q[c];
      for (int64_t k = 0; k < len; k++) {
        out_row[dst + k] = stg_row[src + k];
        if constexpr (has_val<V>) {
          if (out_val) out_val[dst + k] = stg_val[src + k];
        }
      }
      dst += len;
      stage += (int64_t)q[c1] - (int64_t)q[c0];
    }
  }
}
template <typename N>
struct MgColChunksFn {  // chunks of long column k of the list
  const int64_t *list;
  const unsigned *count;
  const N *gptr;
  __device__ int64_t operator()(int64_t k) const {
    if (k >= (int64_t)*count) return 0;
    const int64_t c = list[k];
    return ((int64_t)gptr[c + 1] - (int64_t)gptr[c] + kMgChunk - 1) / kMgChunk;
  }
};
// One CTA per chunk of kMgChunk output entries of a long column: the chunk's sources are
// resolved from the column's per-source counts (shared memory), every thread copies entries.
template <typename I, typename N, typename V>
__global__ void __launch_bounds__(256)
    mg_interleave_long_kernel(int world, const N *__restrict__ cp, int64_t m,
                              const N *__restrict__ gptr, int64_t c0, int64_t c1,
                              const I *__restrict__ stg_row, const V *__restrict__ stg_val,
                              I *__restrict__ out_row, V *__restrict__ out_val,
                              const int64_t *__restrict__ long_list,
                              const unsigned *__restrict__ long_count,
                              const int64_t *__restrict__ first_chunk) {
  __shared__ int64_t s_src[kMgMaxRanks], s_end[kMgMaxRanks + 1];
  const int64_t nlong = (int64_t)*long_count;
  if (nlong == 0) return;
  const int64_t nchunks = first_chunk[nlong];
  const int64_t g0 = (int64_t)gptr[c0];
  for (int64_t w = blockIdx.x; w < nchunks; w += gridDim.x) {
    int64_t lo = 0, hi = nlong;  // last list entry with first_chunk <= w
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (first_chunk[mid] <= w)
        lo = mid;
      else
        hi = mid;
    }
    const int64_t c = long_list[lo];
    __syncthreads();
    if (threadIdx.x == 0) {  // per source: where its piece starts in staging / ends in the column
      int64_t stage = 0, run = 0;
      for (int r = 0; r < world; r++) {
        const N *q = cp + (int64_t)r * (m + 1);
        s_src[r] = stage + ((int64_t)q[c] - (int64_t)q[c0]) - run;  // + position in column
        run += (int64_t)q[c + 1] - (int64_t)q[c];
        s_end[r] = run;
        stage += (int64_t)q[c1] - (int64_t)q[c0];
      }
    }
    __syncthreads();
    const int64_t total = (int64_t)gptr[c + 1] - (int64_t)gptr[c];
    const int64_t k0 = (w - first_chunk[lo]) * kMgChunk;
    const int64_t k1 = k0 + kMgChunk < total ? k0 + kMgChunk : total;
    const int64_t dst = (int64_t)gptr[c] - g0;
    for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
      int r = 0;
      while (r + 1 < world && k >= s_end[r]) r++;
      out_row[dst + k] = ld_stream(stg_row + s_src[r] + k);
      if constexpr (has_val<V>) {
        if (out_val) out_val[dst + k] = ld_stream(stg_val + s_src[r] + k);
      }
    }
  }
}

template <typename N>
__global__ void mg_local_ptr_kernel(const N *__restrict__ new_ptr, int64_t lo, int64_t cnt,
                                    N *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= cnt) out[i] = new_ptr[lo + i] - new_ptr[lo];
}

}  // namespace sb200

using namespace sb200;

// single-GPU entry points this file builds on
extern "C" {
int sb200_coo_to_csr_block(int, int64_t, int64_t, int64_t, int64_t, const void *, const void *,
                           const void *, void *, void *, void *, int, int, int, void *);
int sb200_csr_to_csc_block(int, int64_t, int64_t, int64_t, int64_t, const void *, const void *,
                           const void *, void *, void *, void *, int, int, int, void *);
int sb200_degree_reorder(int, int64_t, const void *, int, void *, int, int, void *);
int sb200_degree_histogram(int, int64_t, const void *, int, int64_t, void *, void *);
int sb200_degree_rank_combine(int, int64_t, const void *, const void *, const void *, int64_t,
                              void *, int, int, void *);
int sb200_degrees(int, int64_t, const void *, void *, int, int, void *);
int sb200_permute2d(int, int64_t, int64_t, int64_t, const void *, const void *, const void *,
                    const void *, const void *, void *, void *, void *, int, int, int, void *);
int sb200_max_degree(int, int64_t, const void *, int, int64_t *, void *);
int sb200_rank_keys(int, int64_t, const void *, int64_t, void *, int, void *);
}

#define SB_RC(expr)                                  \
  do {                                               \
    const int rc__ = (expr);                         \
    if (rc__ != SB200_OK) throw sb200::Error{rc__};  \
  } while (0)

extern "C" {

// ------------------------------------------------------------------ communicator
int sb200_mg_comm_create(int device, int rank, int world, size_t window_bytes,
                         sb200_mg_comm_t **out, void *h_out_handle64) {
  return guarded(device, [&] {
    SB_REQUIRE(out && world >= 1 && world <= kMgMaxRanks && rank >= 0 && rank < world,
               SB200_ERR_BAD_ARG, "bad rank / world (at most %d ranks)", kMgMaxRanks);
    SB_REQUIRE(window_bytes >= kMgControlBytes + 4096, SB200_ERR_BAD_ARG, "window too small");
    auto *c = new sb200_mg_comm();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->window_bytes = window_bytes;
    c->peers.rank = rank;
    c->peers.world = world;
    void *w = nullptr;
    cudaError_t e = cudaMalloc(&w, window_bytes);
    if (e != cudaSuccess) {
      delete c;
      set_error("cudaMalloc of the %zu-byte window failed: %s", window_bytes, cudaGetErrorString(e));
      throw Error{SB200_ERR_ALLOC};
    }
    c->owns_window = true;
    c->peers.win[rank] = (char *)w;
    SB_CUDA(cudaMemset(w, 0, kMgControlBytes));
    SB_CUDA(cudaDeviceSynchronize());
    if (h_out_handle64) {
      static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
      cudaIpcMemHandle_t h;
      if (world > 1)
        SB_CUDA(cudaIpcGetMemHandle(&h, w));
      else
        memset(&h, 0, sizeof(h));
      memcpy(h_out_handle64, &h, 64);
    }
    *out = c;
  });
}

int sb200_mg_comm_connect(sb200_mg_comm_t *c, const void *h_all_handles) {
  if (!c) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    SB_REQUIRE(h_all_handles || c->peers.world == 1, SB200_ERR_BAD_ARG, "null handles");
    for (int r = 0; r < c->peers.world; r++) {
      if (r == c->peers.rank) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char *)h_all_handles + (size_t)r * 64, 64);
      void *p = nullptr;
      SB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      c->peers.win[r] = (char *)p;
    }
    c->ipc = true;
  });
}

int sb200_mg_comm_create_local(int ndev, const int *devices, size_t window_bytes,
                               sb200_mg_comm_t **out_comms) {
  if (!devices || !out_comms || ndev < 1 || ndev > kMgMaxRanks) return SB200_ERR_BAD_ARG;
  for (int r = 0; r < ndev; r++) out_comms[r] = nullptr;
  for (int r = 0; r < ndev; r++) {
    const int rc = sb200_mg_comm_create(devices[r], r, ndev, window_bytes, &out_comms[r], nullptr);
    if (rc != SB200_OK) return rc;
  }
  for (int r = 0; r < ndev; r++) {
    const int rc = guarded(devices[r], [&] {
      for (int q = 0; q < ndev; q++) {
        if (q == r) continue;
        if (devices[q] != devices[r]) {
          int can = 0;
          SB_CUDA(cudaDeviceCanAccessPeer(&can, devices[r], devices[q]));
          SB_REQUIRE(can, SB200_ERR_BAD_DEVICE, "device %d cannot access device %d", devices[r],
                     devices[q]);
          cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled)
            cudaGetLastError();
          else
            SB_CUDA(e);
        }
        out_comms[r]->peers.win[q] = out_comms[q]->peers.win[q];
      }
    });
    if (rc != SB200_OK) return rc;
  }
  return SB200_OK;
}

int sb200_mg_comm_destroy(sb200_mg_comm_t *c) {
  if (!c) return SB200_OK;
  return guarded(c->device, [&] {
    cudaDeviceSynchronize();
    for (int r = 0; r < c->peers.world; r++) {
      if (r == c->peers.rank || !c->peers.win[r]) continue;
      if (c->ipc) cudaIpcCloseMemHandle(c->peers.win[r]);
    }
    if (c->owns_window && c->peers.win[c->peers.rank]) cudaFree(c->peers.win[c->peers.rank]);
    if (c->aux_event) cudaEventDestroy((cudaEvent_t)c->aux_event);
    if (c->aux_stream) cudaStreamDestroy((cudaStream_t)c->aux_stream);
    cudaGetLastError();
    delete c;
  });
}

int sb200_mg_comm_info(const sb200_mg_comm_t *c, int *h_rank, int *h_world, size_t *h_window_bytes) {
  if (!c) return SB200_ERR_BAD_ARG;
  if (h_rank) *h_rank = c->peers.rank;
  if (h_world) *h_world = c->peers.world;
  if (h_window_bytes) *h_window_bytes = c->window_bytes;
  return SB200_OK;
}

// One process, ndev GPUs: fn(rank, user) on one host thread per rank (the operators are
// collective: every rank has to be inside the same call at the same time).  Returns the first
// non-zero code any rank returned.
int sb200_mg_run_ranks(int ndev, int (*fn)(int rank, void *user), void *user) {
  if (ndev < 1 || ndev > kMgMaxRanks || !fn) return SB200_ERR_BAD_ARG;
  std::vector<int> rc(ndev, SB200_OK);
  std::vector<std::thread> th;
  th.reserve(ndev);
  for (int r = 0; r < ndev; r++) th.emplace_back([&, r] { rc[r] = fn(r, user); });
  for (auto &t : th) t.join();
  for (int r = 0; r < ndev; r++)
    if (rc[r] != SB200_OK) return rc[r];
  return SB200_OK;
}

int sb200_mg_barrier(sb200_mg_comm_t *c, void *stream) {
  if (!c) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    mg_barrier(c, (cudaStream_t)stream);
  });
}

// h_out[world] = every rank's value (a collective; synchronises the stream)
int sb200_mg_allgather_i64(sb200_mg_comm_t *c, int64_t value, int64_t *h_out, void *stream) {
  if (!c || !h_out) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws(c->device, st);
    unsigned long long *d = ws.alloc<unsigned long long>(1);
    const unsigned long long v = (unsigned long long)value;
    SB_CUDA(cudaMemcpyAsync(d, &v, sizeof(v), cudaMemcpyHostToDevice, st));
    mg_barrier(c, st);  // nobody still reads slot 0 of an earlier gather
    mg_put_table(c, st, 0, d, 1);
    mg_barrier(c, st);
    MgControl *ctl = c->peers.ctl(c->peers.rank);
    for (int r = 0; r < c->peers.world; r++)
      SB_CUDA(cudaMemcpyAsync(&h_out[r], &ctl->table[r][0], sizeof(int64_t),
                              cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    check_barriers(c);
  });
}

// ------------------------------------------------------------------ COO -> CSR
// The block's COO (all nonzeros of rows [row_lo, row_lo + n_local), (row, col)-sorted as the COO
// constructor leaves them) -> block-local CSR; h_out2 = {global nnz, nnz of the blocks before
// this one}.  Synchronises the stream (to return the totals).
int sb200_mg_coo_to_csr(sb200_mg_comm_t *c, int64_t row_lo, int64_t n_local, int64_t m,
                        int64_t nnz_local, const void *row, const void *col, const void *vals,
                        void *out_row_ptr, void *out_col, void *out_vals, int64_t *h_out2,
                        int id_type, int nnz_type, int val_type, void *stream) {
  if (!c || !h_out2) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    SB_RC(sb200_coo_to_csr_block(c->device, row_lo, n_local, m, nnz_local, row, col, vals,
                                 out_row_ptr, out_col, out_vals, id_type, nnz_type, val_type,
                                 stream));
    std::vector<int64_t> all(c->peers.world);
    SB_RC(sb200_mg_allgather_i64(c, nnz_local, all.data(), stream));
    int64_t total = 0, base = 0;
    for (int r = 0; r < c->peers.world; r++) {
      total += all[r];
      if (r < c->peers.rank) base += all[r];
    }
    h_out2[0] = total;
    h_out2[1] = base;
  });
}

// ------------------------------------------------------------------ Permute1D
// vals / order: this rank's block [bounds[rank], bounds[rank+1]) of the arrays; out: the same
// block of the result out[order[i]] = vals[i].  h_bounds: world+1 block boundaries (host).
int sb200_mg_permute1d(sb200_mg_comm_t *c, const int64_t *h_bounds, const void *vals,
                       const void *order, void *out, int id_type, int val_type, void *stream) {
  if (!c || !h_bounds) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    cudaStream_t st = (cudaStream_t)stream;
    const int world = c->peers.world, rank = c->peers.rank;
    const int64_t n_local = h_bounds[rank + 1] - h_bounds[rank];
    const int vb = dtype_size(val_type);
    SB_REQUIRE(vb == 4 || vb == 8, SB200_ERR_BAD_DTYPE, "val_type %d unsupported", val_type);
    int64_t largest = 0;
    for (int r = 0; r < world; r++)
      largest = std::max(largest, h_bounds[r + 1] - h_bounds[r]);
    MgLayout lay(c);
    const size_t region = lay.take((size_t)largest * vb);
    Workspace ws(c->device, st);
    int64_t *d_bounds = ws.alloc<int64_t>(world + 1);
    SB_CUDA(cudaMemcpyAsync(d_bounds, h_bounds, (world + 1) * sizeof(int64_t),
                            cudaMemcpyHostToDevice, st));
    mg_barrier(c, st);  // the region is free on every rank
    if (n_local > 0) {
      dispatch_id(id_type, [&](auto I_) {
        using I = decltype(I_);
        const unsigned grid = (unsigned)ceil_div(n_local, 256);
        if (vb == 4)
          SB_LAUNCH((mg_permute1d_push_kernel<I, uint32_t>), grid, 256, 0, st, c->peers, region,
                    (const int64_t *)d_bounds, (const uint32_t *)vals, (const I *)order, n_local);
        else
          SB_LAUNCH((mg_permute1d_push_kernel<I, uint64_t>), grid, 256, 0, st, c->peers, region,
                    (const int64_t *)d_bounds, (const uint64_t *)vals, (const I *)order, n_local);
      });
    }
    mg_barrier(c, st);
    if (n_local > 0)
      SB_CUDA(cudaMemcpyAsync(out, c->peers.data(rank) + region, (size_t)n_local * vb,
                              cudaMemcpyDeviceToDevice, st));
    mg_barrier(c, st);  // nobody overwrites the region before it has been copied out
  });
}

// ------------------------------------------------------------------ DegreeReorder
// row_ptr: block-local row_ptr of rows [h_bounds[rank], h_bounds[rank+1]); out_inv[n]: the FULL
// permutation (inv[old] = new) on every rank.  Synchronises the stream once (largest degree).
int sb200_mg_degree_reorder(sb200_mg_comm_t *c, int64_t n, const int64_t *h_bounds,
                            const void *row_ptr, int ascending, void *out_inv, int id_type,
                            int nnz_type, void *stream) {
  if (!c || !h_bounds) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    cudaStream_t st = (cudaStream_t)stream;
    const int world = c->peers.world, rank = c->peers.rank;
    const int64_t lo = h_bounds[rank], nl = h_bounds[rank + 1] - lo;
    const int ib = dtype_size(id_type);
    SB_REQUIRE(ib == 4 || ib == 8, SB200_ERR_BAD_DTYPE, "bad id_type");
    Workspace ws(c->device, st);
    MgTrace tr(c, st, "degree_reorder");
    // block-local order (degree ascending, id descending)
    void *local = ws.alloc_bytes((size_t)(nl > 0 ? nl : 1) * ib);
    SB_RC(sb200_degree_reorder(c->device, nl, row_ptr, 1, local, id_type, nnz_type, stream));
    tr.mark("local_order");
    // largest degree anywhere
    int64_t my_max = 0;
    SB_RC(sb200_max_degree(c->device, nl, row_ptr, nnz_type, &my_max, stream));
    std::vector<int64_t> all(world);
    SB_RC(sb200_mg_allgather_i64(c, my_max, all.data(), stream));
    int64_t maxdeg = 0;
    for (int r = 0; r < world; r++) maxdeg = std::max(maxdeg, all[r]);
    const int64_t nbins = maxdeg + 1;
    tr.mark("max_degree_allgather");
    MgLayout lay(c);
    const size_t hist_region = lay.take((size_t)world * nbins * sizeof(unsigned long long));
    const size_t inv_region = lay.take((size_t)n * ib);
    unsigned long long *hist = ws.alloc<unsigned long long>(nbins);
    SB_RC(sb200_degree_histogram(c->device, nl, row_ptr, nnz_type, nbins, hist, stream));
    mg_barrier(c, st);
    mg_bcast_slice(c, st, hist_region, hist, (size_t)rank * nbins * sizeof(unsigned long long),
                   (size_t)nbins * sizeof(unsigned long long));
    mg_barrier(c, st);
    tr.mark("histogram_allgather");
    const unsigned long long *all_hist =
        reinterpret_cast<const unsigned long long *>(c->peers.data(rank) + hist_region);
    int64_t *g_start = ws.alloc<int64_t>(nbins + 1), *l_start = ws.alloc<int64_t>(nbins + 1);
    int64_t *offset = ws.alloc<int64_t>(nbins);
    exclusive_scan<int64_t>(ws, MgHistTotalFn{all_hist, world, nbins}, g_start, nbins);
    exclusive_scan<int64_t>(ws, MgHistMineFn{all_hist, rank, nbins}, l_start, nbins);
    SB_LAUNCH(mg_degree_offset_kernel, (unsigned)ceil_div(nbins, 256), 256, 0, st, all_hist, rank,
              world, nbins, (const int64_t *)g_start, (const int64_t *)l_start, offset);
    void *part = ws.alloc_bytes((size_t)(nl > 0 ? nl : 1) * ib);
    SB_RC(sb200_degree_rank_combine(c->device, nl, row_ptr, local, offset,
                                    ascending ? -1 : n - 1, part, id_type, nnz_type, stream));
    tr.mark("offsets_and_rank");
    mg_bcast_slice(c, st, inv_region, part, (size_t)lo * ib, (size_t)nl * ib);
    tr.mark("bcast_slice");
    mg_barrier(c, st);
    tr.mark("barrier");
    SB_CUDA(cudaMemcpyAsync(out_inv, c->peers.data(rank) + inv_region, (size_t)n * ib,
                            cudaMemcpyDeviceToDevice, st));
    mg_barrier(c, st);
    tr.mark("copy_out");
  });
}

// ------------------------------------------------------------------ Permute2D
// Input: block-local CSR of rows [h_bounds[rank], h_bounds[rank+1]) + the FULL inverse
// permutations (replicated; NULL = identity).  The result is sharded by nnz-balanced blocks of
// the NEW rows.  Two calls:
//   sb200_mg_permute2d_run    does all the work; the permuted block is left in this rank's window.
//                             h_out_bounds[world+1] = the new row blocks, h_out2[3] = {rows, nnz,
//                             nnz of the blocks before} of this rank's new block.  Synchronises.
//   sb200_mg_permute2d_fetch  copies the block out of the window (row_ptr block-local) and
//                             releases the window for the next operator.
int sb200_mg_permute2d_run(sb200_mg_comm_t *c, int64_t n, int64_t m, int64_t nnz_total,
                           const int64_t *h_bounds, const void *row_ptr, const void *col,
                           const void *vals, const void *row_order, const void *col_order,
                           int64_t *h_out_bounds, int64_t *h_out2, int id_type, int nnz_type,
                           int val_type, void *stream) {
  if (!c || !h_bounds || !h_out_bounds || !h_out2) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    cudaStream_t st = (cudaStream_t)stream;
    const int world = c->peers.world, rank = c->peers.rank;
    const int64_t lo = h_bounds[rank], nl = h_bounds[rank + 1] - lo;
    const bool hv = vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      Workspace ws(c->device, st);
      MgTrace tr(c, st, "permute2d_run");
      MgLayout lay(c);
      const size_t deg_region = lay.take((size_t)(n + 1) * sizeof(N));
      // (the received block: nnz-balanced, so about nnz_total / world entries plus a row of
      // slack per boundary; it gets what is left of the window)
      const size_t left = c->window_bytes - kMgControlBytes - lay.used();
      SB_REQUIRE(left > 4096, SB200_ERR_BAD_ARG, "multi-GPU window too small");
      const size_t room = (left - 1024) / (sizeof(I) + (has_val<V> ? sizeof(V) : 0));
      const size_t col_region = lay.take(room * sizeof(I));
      const size_t val_region = has_val<V> ? lay.take(room * sizeof(V)) : 0;
      // ---- every row's degree, everywhere
      N *deg_local = ws.alloc<N>(nl > 0 ? nl : 1);
      SB_RC(sb200_degrees(c->device, nl, row_ptr, deg_local, nnz_type, nnz_type, stream));
      mg_barrier(c, st);  // the window is free on every rank
      mg_bcast_slice(c, st, deg_region, deg_local, (size_t)lo * sizeof(N), (size_t)nl * sizeof(N));
      mg_barrier(c, st);
      tr.mark("degrees_allgather");
      const N *deg = reinterpret_cast<const N *>(c->peers.data(rank) + deg_region);
      // ---- replicated: degrees in the new order, new row_ptr, nnz-balanced new row blocks
      N *new_deg = ws.alloc<N>(n > 0 ? n : 1), *new_ptr = ws.alloc<N>(n + 1);
      if (n > 0)
        SB_LAUNCH((mg_new_degree_kernel<I, N>), (unsigned)ceil_div(n, 256), 256, 0, st, deg,
                  (const I *)row_order, n, new_deg);
      exclusive_scan<N>(ws, LoadFn<N>{new_deg}, new_ptr, n);
      int64_t *nb = ws.alloc<int64_t>(world + 1), *nb_at = ws.alloc<int64_t>(world + 1);
      SB_LAUNCH((mg_balance_kernel<N>), 1, 32, 0, st, (const N *)new_ptr, n, nnz_total, world, nb,
                nb_at);
      // ---- my rows: renumber the columns and sort every row (rows stay in their old order)
      N h_nnz_local = 0;
      SB_CUDA(cudaMemcpyAsync(&h_nnz_local, (const N *)row_ptr + nl, sizeof(N),
                              cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaMemcpyAsync(h_out_bounds, nb, (world + 1) * sizeof(int64_t),
                              cudaMemcpyDeviceToHost, st));
      std::vector<int64_t> h_at(world + 1);
      SB_CUDA(cudaMemcpyAsync(h_at.data(), nb_at, (world + 1) * sizeof(int64_t),
                              cudaMemcpyDeviceToHost, st));
      tr.mark("new_ptr_and_blocks");
      SB_CUDA(cudaStreamSynchronize(st));
      const int64_t my_nnz = (int64_t)h_nnz_local;
      const int64_t new_rows = h_out_bounds[rank + 1] - h_out_bounds[rank];
      const int64_t new_nnz = h_at[rank + 1] - h_at[rank];
      SB_REQUIRE((size_t)new_nnz <= room, SB200_ERR_BAD_ARG,
                 "multi-GPU window too small for the permuted block (%lld entries, room for %zu)",
                 (long long)new_nnz, room);
      // ---- my rows: renumber the columns and sort every row (rows stay in their old order),
      // then every row goes to its final place in the owner's window.  The block is cut into
      // chunks of rows with about the same number of entries: while the SMs sort chunk k + 1,
      // the stores of chunk k are on the wire (second stream) -- the push of a block of 2-entry
      // rows is bound by the small NVLink writes, not by anything the sort needs.
      N *t_ptr = ws.alloc<N>(nl + 1 + kMgMaxChunks);
      I *t_col = ws.alloc<I>(my_nnz > 0 ? my_nnz : 1);
      V *t_val = nullptr;
      if constexpr (has_val<V>) t_val = ws.alloc<V>(my_nnz > 0 ? my_nnz : 1);
      int chunks = mg_push_chunks();
      if (world == 1 || my_nnz < (int64_t)chunks * mg_push_chunk_min()) chunks = 1;
      std::vector<int64_t> cb(chunks + 1, 0), cat(chunks + 1, 0);
      cb[chunks] = nl;
      cat[chunks] = my_nnz;
      if (chunks > 1) {
        int64_t *d_cb = ws.alloc<int64_t>(chunks + 1), *d_cat = ws.alloc<int64_t>(chunks + 1);
        SB_LAUNCH((mg_balance_kernel<N>), 1, 32, 0, st, (const N *)row_ptr, nl, my_nnz, chunks,
                  d_cb, d_cat);
        SB_CUDA(cudaMemcpyAsync(cb.data(), d_cb, (chunks + 1) * sizeof(int64_t),
                                cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaMemcpyAsync(cat.data(), d_cat, (chunks + 1) * sizeof(int64_t),
                                cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
      }
      cudaStream_t push_st = chunks > 1 ? mg_aux_stream(c) : st;
      cudaEvent_t ev = mg_aux_event(c);
      const int sms = device_info(c->device).sm_count;
      N *sub_ptr = t_ptr;
      for (int k = 0; k < chunks; k++) {
        const int64_t r0 = cb[k], rows = cb[k + 1] - cb[k];
        const int64_t e0 = cat[k], ents = cat[k + 1] - cat[k];
        if (rows <= 0) continue;
        // (the chunk as a CSR of its own: row_ptr rebased to the chunk's first entry)
        N *in_ptr = ws.alloc<N>(rows + 1);
        SB_LAUNCH((mg_rebase_ptr_kernel<N>), (unsigned)ceil_div(rows + 1, 256), 256, 0, st,
                  (const N *)row_ptr + r0, rows + 1, (N)e0, in_ptr);
        SB_RC(sb200_permute2d(c->device, rows, m, ents, in_ptr, (const I *)col + e0,
                              hv ? (const void *)((const V *)vals + e0) : nullptr, nullptr,
                              col_order, sub_ptr, t_col + e0, hv ? (void *)(t_val + e0) : nullptr,
                              id_type, nnz_type, hv ? val_type : SB200_VOID, stream));
        if (ents > 0) {
          if (push_st != st) {
            SB_CUDA(cudaEventRecord(ev, st));
            SB_CUDA(cudaStreamWaitEvent(push_st, ev, 0));
          }
          if (chunks == 1) tr.mark("local_renumber_sort");
          // one chunk: the rows are pushed in the order of their NEW ids (sorted local index,
          // scattered local reads of the sorted rows) -- consecutive rows then land at
          // increasing, mostly adjacent addresses of the same window, and the stores over NVLink
          // are long runs even when the rows hold 2 entries (source order: 180 GB/s from the
          // rank with the low-degree rows of R-MAT-26 at 4 GPUs)
          const I *sidx = nullptr, *dest_id = nullptr;
          const N *dptr = nullptr;
          if (chunks == 1 && row_order && mg_push_dest_order()) {
            using UI = typename std::make_unsigned<I>::type;
            UI *k0 = ws.alloc<UI>(rows), *i0 = ws.alloc<UI>(rows);
            UI *k1 = ws.alloc<UI>(rows), *i1 = ws.alloc<UI>(rows);
            UI *k2 = ws.alloc<UI>(rows), *i2 = ws.alloc<UI>(rows);
            SB_LAUNCH((mg_order_keys_kernel<I, UI>), (unsigned)ceil_div(rows, 256), 256, 0, st,
                      (const I *)row_order + lo, rows, k0, i0);
            radix_sort<UI, UI, NoVal>(ws, {k0, i0, nullptr}, {k1, i1, nullptr}, {k2, i2, nullptr},
                                      rows, {{0, bits_for((uint64_t)(n > 1 ? n - 1 : 1))}});
            N *dp = ws.alloc<N>(rows + 1);
            exclusive_scan<N>(ws, MgWalkLenFn<I, N>{(const N *)sub_ptr, (const I *)i1}, dp, rows);
            sidx = (const I *)i1;
            dest_id = (const I *)k1;
            dptr = dp;
          }
          if (chunks == 1) tr.mark("push_order");
          // (chunks overlapped with the next chunk's sort: a push that filled every SM with
          // warps waiting on NVLink kept the sort kernels out)
          const int64_t warps = std::min<int64_t>(ceil_div(ents, kMgPushChunk),
                                                  (int64_t)sms * (push_st != st ? 16 : 64));
          SB_LAUNCH((mg_push_rows_kernel<I, N, V>), (unsigned)ceil_div(warps * 32, 256), 256, 0,
                    push_st, c->peers, col_region, val_region, (const N *)sub_ptr,
                    (const I *)(t_col + e0), (const V *)(hv ? t_val + e0 : nullptr),
                    (const I *)row_order, lo + r0, rows, ents, (const N *)new_ptr,
                    (const int64_t *)nb, (const int64_t *)nb_at, sidx, dest_id, dptr);
        }
        sub_ptr += rows + 1;
      }
      if (chunks > 1) tr.mark("local_renumber_sort");
      if (push_st != st) {
        SB_CUDA(cudaEventRecord(ev, push_st));
        SB_CUDA(cudaStreamWaitEvent(st, ev, 0));
      }
      tr.mark("push_rows");
      mg_barrier(c, st);
      tr.mark("barrier");
      // the new block's row_ptr waits in the (now unused) degree region of my own window
      N *keep = reinterpret_cast<N *>(c->peers.data(rank) + deg_region);
      SB_LAUNCH((mg_local_ptr_kernel<N>), (unsigned)ceil_div(new_rows + 1, 256), 256, 0, st,
                (const N *)new_ptr, h_out_bounds[rank], new_rows, keep);
      SB_CUDA(cudaStreamSynchronize(st));
      check_barriers(c);
      h_out2[0] = new_rows;
      h_out2[1] = new_nnz;
      h_out2[2] = h_at[rank];
    });
  });
}

int sb200_mg_permute2d_fetch(sb200_mg_comm_t *c, int64_t n, int64_t new_rows, int64_t new_nnz,
                             void *out_row_ptr, void *out_col, void *out_vals, int id_type,
                             int nnz_type, int val_type, void *stream) {
  if (!c) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    cudaStream_t st = (cudaStream_t)stream;
    const int rank = c->peers.rank;
    const bool hv = out_vals != nullptr && val_type != SB200_VOID;
    const size_t ib = dtype_size(id_type), nb = dtype_size(nnz_type), vb = dtype_size(val_type);
    // the same carving as sb200_mg_permute2d_run
    MgLayout lay(c);
    const size_t deg_region = lay.take((size_t)(n + 1) * nb);
    const size_t left = c->window_bytes - kMgControlBytes - lay.used();
    const size_t room = (left - 1024) / (ib + (hv ? vb : 0));
    const size_t col_region = lay.take(room * ib);
    const size_t val_region = hv ? lay.take(room * vb) : 0;
    char *mine = c->peers.data(rank);
    MgTrace tr(c, st, "permute2d_fetch");
    SB_CUDA(cudaMemcpyAsync(out_row_ptr, mine + deg_region, (size_t)(new_rows + 1) * nb,
                            cudaMemcpyDeviceToDevice, st));
    if (new_nnz > 0) {
      SB_CUDA(cudaMemcpyAsync(out_col, mine + col_region, (size_t)new_nnz * ib,
                              cudaMemcpyDeviceToDevice, st));
      if (hv)
        SB_CUDA(cudaMemcpyAsync(out_vals, mine + val_region, (size_t)new_nnz * vb,
                                cudaMemcpyDeviceToDevice, st));
    }
    tr.mark("copy_out");
    mg_barrier(c, st);  // the window may be reused once every rank has copied its block out
    tr.mark("barrier");
  });
}

// ------------------------------------------------------------------ CSR -> CSC
// Input: block-local CSR of rows [h_bounds[rank], h_bounds[rank+1]).  The result is sharded by
// nnz-balanced blocks of COLUMNS.  Two calls, as for Permute2D:
//   sb200_mg_csr_to_csc_run    local transpose of the block, col_ptr tables all-gathered through
//                              the windows, column slices stored into their owners' windows.
//                              h_out_bounds[world+1] = the column blocks, h_out2[3] = {columns,
//                              nnz, nnz of the blocks before} of this rank's block.  Synchronises.
//   sb200_mg_csr_to_csc_fetch  interleaves the staged slices into the caller's arrays
//                              (col_ptr block-local, rows global ids, ascending inside a column).
int sb200_mg_csr_to_csc_run(sb200_mg_comm_t *c, int64_t n, int64_t m, int64_t nnz_total,
                            const int64_t *h_bounds, const void *row_ptr, const void *col,
                            const void *vals, int64_t *h_out_bounds, int64_t *h_out2,
                            int id_type, int nnz_type, int val_type, void *stream) {
  if (!c || !h_bounds || !h_out_bounds || !h_out2) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    cudaStream_t st = (cudaStream_t)stream;
    const int world = c->peers.world, rank = c->peers.rank;
    const int64_t lo = h_bounds[rank], nl = h_bounds[rank + 1] - lo;
    const bool hv = vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      Workspace ws(c->device, st);
      MgTrace tr(c, st, "csr_to_csc_run");
      MgLayout lay(c);
      const size_t cp_region = lay.take((size_t)world * (m + 1) * sizeof(N));
      const size_t gptr_region = lay.take((size_t)(m + 1) * sizeof(N));
      const size_t left = c->window_bytes - kMgControlBytes - lay.used();
      SB_REQUIRE(left > 4096, SB200_ERR_BAD_ARG, "multi-GPU window too small");
      const size_t room = (left - 1024) / (sizeof(I) + (has_val<V> ? sizeof(V) : 0));
      const size_t row_region = lay.take(room * sizeof(I));
      const size_t val_region = has_val<V> ? lay.take(room * sizeof(V)) : 0;
      // ---- local transpose of the block (rows ascending inside a column, global row ids)
      N h_nnz_local = 0;
      SB_CUDA(cudaMemcpyAsync(&h_nnz_local, (const N *)row_ptr + nl, sizeof(N),
                              cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaStreamSynchronize(st));
      const int64_t my_nnz = (int64_t)h_nnz_local;
      N *cp = ws.alloc<N>(m + 1);
      I *rows = ws.alloc<I>(my_nnz > 0 ? my_nnz : 1);
      V *tv = nullptr;
      if constexpr (has_val<V>) tv = ws.alloc<V>(my_nnz > 0 ? my_nnz : 1);
      SB_RC(sb200_csr_to_csc_block(c->device, lo, nl, m, my_nnz, row_ptr, col,
                                   hv ? vals : nullptr, cp, rows, tv, id_type, nnz_type,
                                   hv ? val_type : SB200_VOID, stream));
      tr.mark("local_transpose");
      // ---- every rank's col_ptr table, everywhere
      mg_barrier(c, st);  // the window is free on every rank
      mg_bcast_slice(c, st, cp_region, cp, (size_t)rank * (m + 1) * sizeof(N),
                     (size_t)(m + 1) * sizeof(N));
      mg_barrier(c, st);
      tr.mark("col_ptr_allgather");
      const N *all_cp = reinterpret_cast<const N *>(c->peers.data(rank) + cp_region);
      N *gptr = reinterpret_cast<N *>(c->peers.data(rank) + gptr_region);
      exclusive_scan<N>(ws, MgColTotalFn<N>{all_cp, world, m}, gptr, m);
      int64_t *cb = ws.alloc<int64_t>(world + 1), *cb_at = ws.alloc<int64_t>(world + 1);
      SB_LAUNCH((mg_balance_kernel<N>), 1, 32, 0, st, (const N *)gptr, m, nnz_total, world, cb,
                cb_at);
      tr.mark("global_ptr_and_blocks");
      // ---- my slice of every destination's columns, into its staging area
      if (my_nnz > 0) {
        dim3 grid((unsigned)std::max(1, device_info(c->device).sm_count * 4 / world),
                  (unsigned)world);
        SB_LAUNCH((mg_push_cols_kernel<I, N, V>), grid, 256, 0, st, c->peers, row_region,
                  val_region, all_cp, m, (const int64_t *)cb, (const I *)rows, (const V *)tv);
      }
      tr.mark("push_cols");
      mg_barrier(c, st);
      tr.mark("barrier");
      std::vector<int64_t> h_at(world + 1);
      SB_CUDA(cudaMemcpyAsync(h_out_bounds, cb, (world + 1) * sizeof(int64_t),
                              cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaMemcpyAsync(h_at.data(), cb_at, (world + 1) * sizeof(int64_t),
                              cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaStreamSynchronize(st));
      check_barriers(c);
      h_out2[0] = h_out_bounds[rank + 1] - h_out_bounds[rank];
      h_out2[1] = h_at[rank + 1] - h_at[rank];
      h_out2[2] = h_at[rank];
      SB_REQUIRE((size_t)h_out2[1] <= room, SB200_ERR_BAD_ARG,
                 "multi-GPU window too small for the transposed block (%lld entries, room for %zu)",
                 (long long)h_out2[1], room);
    });
  });
}

int sb200_mg_csr_to_csc_fetch(sb200_mg_comm_t *c, int64_t m, int64_t col_lo, int64_t n_cols,
                              void *out_col_ptr, void *out_row, void *out_vals, int id_type,
                              int nnz_type, int val_type, void *stream) {
  if (!c) return SB200_ERR_BAD_ARG;
  return guarded(c->device, [&] {
    check_comm(c);
    cudaStream_t st = (cudaStream_t)stream;
    const int world = c->peers.world, rank = c->peers.rank;
    const bool hv = out_vals != nullptr && val_type != SB200_VOID;
    dispatch_inv(id_type, nnz_type, val_type, hv, [&](auto I_, auto N_, auto V_) {
      using I = decltype(I_);
      using N = decltype(N_);
      using V = decltype(V_);
      MgLayout lay(c);  // the same carving as sb200_mg_csr_to_csc_run
      const size_t cp_region = lay.take((size_t)world * (m + 1) * sizeof(N));
      const size_t gptr_region = lay.take((size_t)(m + 1) * sizeof(N));
      const size_t left = c->window_bytes - kMgControlBytes - lay.used();
      const size_t room = (left - 1024) / (sizeof(I) + (has_val<V> ? sizeof(V) : 0));
      const size_t row_region = lay.take(room * sizeof(I));
      const size_t val_region = has_val<V> ? lay.take(room * sizeof(V)) : 0;
      char *mine = c->peers.data(rank);
      const N *all_cp = reinterpret_cast<const N *>(mine + cp_region);
      const N *gptr = reinterpret_cast<const N *>(mine + gptr_region);
      if (n_cols > 0) {
        const int sms = device_info(c->device).sm_count;
        Workspace ws(c->device, st);
        MgTrace tr(c, st, "csr_to_csc_fetch");
        // (a long column has >= kMgLongCol entries, so there are at most nnz / kMgLongCol)
        N h_ends[2] = {0, 0};  // the block's nnz bounds the number of long columns in it
        SB_CUDA(cudaMemcpyAsync(&h_ends[0], gptr + col_lo, sizeof(N), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaMemcpyAsync(&h_ends[1], gptr + col_lo + n_cols, sizeof(N),
                                cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        const int64_t long_cap = (int64_t)(h_ends[1] - h_ends[0]) / kMgLongCol + 1;
        int64_t *long_list = ws.alloc<int64_t>(long_cap);
        unsigned *long_count = ws.alloc<unsigned>(1);
        SB_CUDA(cudaMemsetAsync(long_count, 0, sizeof(unsigned), st));
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n_cols, 256), (int64_t)sms * 16);
        SB_LAUNCH((mg_interleave_kernel<I, N, V>), grid, 256, 0, st, world, all_cp, m, gptr, col_lo,
                  col_lo + n_cols, (const I *)(mine + row_region), (const V *)(mine + val_region),
                  (N *)out_col_ptr, (I *)out_row, (V *)out_vals, long_list, long_count);
        tr.mark("interleave_short");
        int64_t *first_chunk = ws.alloc<int64_t>(long_cap + 1);
        exclusive_scan<int64_t>(ws, MgColChunksFn<N>{long_list, long_count, gptr}, first_chunk,
                                long_cap);
        SB_LAUNCH((mg_interleave_long_kernel<I, N, V>), sms * 8, 256, 0, st, world, all_cp, m, gptr,
                  col_lo, col_lo + n_cols, (const I *)(mine + row_region),
                  (const V *)(mine + val_region), (I *)out_row, (V *)out_vals,
                  (const int64_t *)long_list, (const unsigned *)long_count,
                  (const int64_t *)first_chunk);
        tr.mark("interleave_long");
      } else {
        SB_CUDA(cudaMemsetAsync(out_col_ptr, 0, sizeof(N), st));
      }
      mg_barrier(c, st);  // the window may be reused once every rank is done with its staging
    });
  });
}

}  // extern "C"
