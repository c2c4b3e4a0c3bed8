// rcm.cu -- order-exact Reverse Cuthill-McKee on the GPU.
//
// Reproduces RCMReorder::GetReorderCSR + RCMReorder::peripheral
// (reorder/rcm_reorder.cc:22-166) bit for bit, level-synchronously:
//   * peripheral(): FIFO BFS.  The next level in queue order == newly reached vertices ordered
//     by the expansion slot (position of the parent in the current level, index inside the
//     parent's adjacency) of their FIRST discoverer -> atomicMin of the slot per vertex.
//   * Cuthill-McKee BFS: a popped vertex pushes its unvisited neighbours through a min-heap on
//     (degree, id) -> next level == newly reached vertices sorted by (queue position of the
//     first parent, degree, id) -> atomicMin of the parent position, then a per-parent sort.
//
// Two execution regimes share all global state (mark[], the queues, RcmState):
//   NARROW  one persistent 1024-thread CTA walks levels without returning to the host while
//           the frontier fits shared memory (high-diameter graphs: grids, bands -- thousands
//           to millions of levels); warp-cooperative neighbour expansion, claims by atomicMin
//           in L2, ordered compaction by warp prefix sums, sibling sort by enumeration.
//   WIDE    one level at a time driven by the host with grid-wide kernels (power-law / random
//           graphs: a handful of levels, millions of vertices): exclusive scan of frontier
//           degrees, claim sweep, collect sweep, radix sort on the level's order key, commit.
// The persistent kernel is a resumable state machine: when a level does not fit it stores its
// state and returns NEED_WIDE (or NEED_RESET / NEED_INVERT for bulk array work); the host runs
// that step with all SMs and relaunches it.
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {

constexpr unsigned kUnvisited = 0xffffffffu;
constexpr int kNwBlock = 1024;
constexpr int kNwWarps = kNwBlock / 32;
constexpr int kNwECap = 1 << 16;    // max expansion slots of a level handled by the narrow CTA
constexpr int kNwGroupCap = 96;     // max degree in a CM frontier (bounds sibling groups)
constexpr int64_t kBulkThreshold = 1 << 15;  // resets / inversions larger than this go wide

enum RcmPhase {
  PH_FIND = 0,
  PH_PBFS_INIT,
  PH_PBFS_LEVEL,
  PH_PBFS_END,
  PH_PBFS_AFTER_RESET,
  PH_CM_INIT,
  PH_CM_LEVEL,
  PH_CM_END,
  PH_CM_AFTER_INVERT,
  PH_DONE
};
enum RcmStatus { ST_RUNNING = 0, ST_DONE, ST_NEED_WIDE, ST_NEED_RESET, ST_NEED_INVERT };

struct RcmState {
  int64_t next_i;      // component scan position (rcm_reorder.cc:104)
  int64_t qwp;         // vertices placed so far (global CM queue length)
  int64_t qst;         // start of the current component in Q
  int64_t root;        // r of peripheral() / perv
  int64_t rlevel, qlevel;
  int64_t lvl_begin, lvl_end, prev_begin;  // frontier slice in the active queue
  int64_t depth;       // distance of the current frontier from the root
  int64_t new_root_pending;  // root chosen by PBFS_END, applied after the reset
  int32_t phase;
  int32_t status;
  int32_t next_phase_after_reset;
  int32_t pad;
  int64_t stat_levels_narrow, stat_levels_wide, stat_bfs, stat_components;
};

template <typename I, typename N>
struct RcmArgs {
  int64_t n;
  const N *xadj;
  const I *adj;
  unsigned *mark;
  I *Q;    // CM order (reference: Q), concatenated components, not yet reversed
  I *Qp;   // peripheral BFS queue (reference: Qp)
  I *inv;  // result: inv[Qp2[i]] = i
  RcmState *state;
  int force_wide;  // testing: never take the narrow path
};

// ------------------------------------------------------------------------------------
// Warp-cooperative expansion of a group of (up to) 32 frontier vertices.  Lane l holds the
// adjacency start xs and degree d of vertex g+l; the warp walks the concatenated adjacency
// lists 32 slots at a time.  f(slot_in_group, owner_lane, adjacency_position, valid) is called
// once per round by every lane (valid == false for the padding of the last round).
// ------------------------------------------------------------------------------------
template <typename Fn>
__device__ __forceinline__ void warp_expand(int64_t xs, unsigned d, Fn &&f) {
  const unsigned lane = lane_id();
  const unsigned incl = warp_inclusive_scan(d);
  const unsigned excl = incl - d;
  const unsigned tot = __shfl_sync(0xffffffffu, incl, 31);
  for (unsigned base = 0; base < tot; base += 32) {
    const unsigned s = base + lane;
    unsigned lo = 0;  // number of lanes whose inclusive end <= s  == owner lane
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const unsigned val = __shfl_sync(0xffffffffu, incl, (lo + step - 1) & 31);
      if (val <= s) lo += step;
    }
    const unsigned j = lo & 31;
    const int64_t xs_j = __shfl_sync(0xffffffffu, xs, j);
    const unsigned excl_j = __shfl_sync(0xffffffffu, excl, j);
    f(s, j, xs_j + (int64_t)(s - excl_j), s < tot);
  }
}

// ------------------------------------------------------------------------------------
// NARROW regime
// ------------------------------------------------------------------------------------
template <typename I>
struct NwCaps {
  static constexpr int kF = sizeof(I) == 4 ? 5120 : 4096;  // frontier / candidate capacity
};

template <typename I>
struct NwSmem {
  static constexpr int kF = NwCaps<I>::kF;
  I F[kF];          // frontier vertices
  int64_t Fx[kF];   // xadj[F[i]]
  unsigned Fd[kF];  // degree of F[i]
  unsigned off[kF + 1];
  I cv[kF];         // candidates: vertex
  unsigned ci[kF];  // parent position in the frontier
  unsigned cd[kF];  // degree
  int64_t cx[kF];   // xadj[cv]
  unsigned wtot[kNwWarps + 2];
  unsigned scratch[34];
  unsigned long long red[kNwWarps];
  RcmState S;
  int in_smem;
  int c_total;
  int flag;
};

template <typename I, typename N>
__device__ void nw_load_frontier(const RcmArgs<I, N> &a, NwSmem<I> &s, const I *queue) {
  const int f = (int)(s.S.lvl_end - s.S.lvl_begin);
  for (int i = threadIdx.x; i < f; i += kNwBlock) {
    const I v = queue[s.S.lvl_begin + i];
    const int64_t xs = (int64_t)a.xadj[v];
    s.F[i] = v;
    s.Fx[i] = xs;
    s.Fd[i] = (unsigned)((int64_t)a.xadj[v + 1] - xs);
  }
  __syncthreads();
}

// One BFS level in shared memory.  Returns the number of newly reached vertices, or -1 when
// the level has to be done by the wide path (nothing has been modified in that case).
template <typename I, typename N, bool CM>
__device__ int nw_level(const RcmArgs<I, N> &a, NwSmem<I> &s, I *queue) {
  constexpr int kF = NwSmem<I>::kF;
  constexpr int kPer = (kF + kNwBlock - 1) / kNwBlock;
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const int f = (int)(s.S.lvl_end - s.S.lvl_begin);
  if (!s.in_smem) nw_load_frontier(a, s, queue);

  // ---- exclusive scan of the frontier degrees -> expansion slot offsets ----
  unsigned local[kPer], sum = 0, mx = 0;
#pragma unroll
  for (int k = 0; k < kPer; k++) {
    const int i = threadIdx.x * kPer + k;
    local[k] = i < f ? s.Fd[i] : 0u;
    sum += local[k];
    mx = local[k] > mx ? local[k] : mx;
  }
  unsigned total;
  unsigned run = block_exclusive_scan(sum, s.scratch, &total);
  mx = warp_reduce_max(mx);
  if (lane == 0) s.wtot[wid] = mx;
#pragma unroll
  for (int k = 0; k < kPer; k++) {
    const int i = threadIdx.x * kPer + k;
    if (i < f) s.off[i] = run;
    run += local[k];
  }
  __syncthreads();
  unsigned maxdeg = 0;
  for (int w = 0; w < kNwWarps; w++) maxdeg = s.wtot[w] > maxdeg ? s.wtot[w] : maxdeg;
  __syncthreads();
  if (total > (unsigned)kNwECap || (CM && maxdeg > (unsigned)kNwGroupCap)) return -1;

  // each warp owns a contiguous range of the frontier (keeps slot order inside the warp)
  int vpw = (f + kNwWarps - 1) / kNwWarps;
  vpw = (vpw + 31) & ~31;
  const int wbeg = wid * vpw < f ? wid * vpw : f;
  const int wend = wbeg + vpw < f ? wbeg + vpw : f;

  // ---- sweep 1: claims ----
  for (int g = wbeg; g < wend; g += 32) {
    const int i = g + lane;
    const int64_t xs = i < wend ? s.Fx[i] : 0;
    const unsigned d = i < wend ? s.Fd[i] : 0u;
    const unsigned slot0 = s.off[g];
    warp_expand(xs, d, [&](unsigned sl, unsigned j, int64_t p, bool valid) {
      if (valid) {
        const I v = a.adj[p];
        const unsigned key = CM ? (unsigned)(g + j) + 1u : slot0 + sl + 1u;
        atomicMin(&a.mark[v], key);
      }
    });
  }
  __syncthreads();

  // ---- sweep 2: who won?  (winner bits of the first 64 rounds are kept in registers) ----
  unsigned long long wbits = 0;
  unsigned wcount = 0, round = 0;
  for (int g = wbeg; g < wend; g += 32) {
    const int i = g + lane;
    const int64_t xs = i < wend ? s.Fx[i] : 0;
    const unsigned d = i < wend ? s.Fd[i] : 0u;
    const unsigned slot0 = s.off[g];
    warp_expand(xs, d, [&](unsigned sl, unsigned j, int64_t p, bool valid) {
      bool win = false;
      if (valid) {
        const I v = a.adj[p];
        const unsigned key = CM ? (unsigned)(g + j) + 1u : slot0 + sl + 1u;
        win = __ldcg(&a.mark[v]) == key;
      }
      wcount += __popc(__ballot_sync(0xffffffffu, win));
      if (win && round < 64) wbits |= 1ull << round;
      round++;
    });
  }
  if (lane == 0) s.wtot[wid] = wcount;
  __syncthreads();
  if (wid == 0) {
    const unsigned v = s.wtot[lane];
    const unsigned inc = warp_inclusive_scan(v);
    s.wtot[lane] = inc - v;
    if (lane == 31) s.c_total = (int)inc;
  }
  __syncthreads();
  const int c = s.c_total;
  const bool overflow = c > kF;

  // ---- sweep 3: ordered compaction of the winners (or roll the claims back) ----
  {
    unsigned pos = s.wtot[wid];
    round = 0;
    for (int g = wbeg; g < wend; g += 32) {
      const int i = g + lane;
      const int64_t xs = i < wend ? s.Fx[i] : 0;
      const unsigned d = i < wend ? s.Fd[i] : 0u;
      const unsigned slot0 = s.off[g];
      warp_expand(xs, d, [&](unsigned sl, unsigned j, int64_t p, bool valid) {
        bool win = false;
        I v = 0;
        if (valid) {
          v = a.adj[p];
          if (round < 64) {
            win = (wbits >> round) & 1ull;
          } else {
            const unsigned key = CM ? (unsigned)(g + j) + 1u : slot0 + sl + 1u;
            win = __ldcg(&a.mark[v]) == key;
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, win);
        if (win) {
          if (overflow) {
            atomicExch(&a.mark[v], kUnvisited);  // undo: the wide path redoes this level
          } else {
            const unsigned k = pos + __popc(bal & lanemask_lt());
            s.cv[k] = v;
            s.ci[k] = (unsigned)(g + j);
          }
        }
        pos += __popc(bal);
        round++;
      });
    }
  }
  __syncthreads();
  if (overflow) return -1;

  // ---- new vertices: mark visited, fetch their adjacency extent ----
  for (int k = threadIdx.x; k < c; k += kNwBlock) {
    const I v = s.cv[k];
    atomicExch(&a.mark[v], 0u);
    const int64_t xs = (int64_t)a.xadj[v];
    s.cx[k] = xs;
    s.cd[k] = (unsigned)((int64_t)a.xadj[v + 1] - xs);
  }
  __syncthreads();

  // ---- next frontier: slot order (peripheral) or (parent, degree, id) order (CM) ----
  const int64_t out0 = s.S.lvl_end;
  for (int k = threadIdx.x; k < c; k += kNwBlock) {
    int pos = k;
    const I v = s.cv[k];
    if (CM) {
      const unsigned par = s.ci[k], dg = s.cd[k];
      int rank = 0, left = 0;
      for (int q = k - 1; q >= 0 && s.ci[q] == par; q--) {
        left++;
        rank += (s.cd[q] < dg || (s.cd[q] == dg && s.cv[q] < v)) ? 1 : 0;
      }
      for (int q = k + 1; q < c && s.ci[q] == par; q++)
        rank += (s.cd[q] < dg || (s.cd[q] == dg && s.cv[q] < v)) ? 1 : 0;
      pos = k - left + rank;
    }
    s.F[pos] = v;
    s.Fx[pos] = s.cx[k];
    s.Fd[pos] = s.cd[k];
    queue[out0 + pos] = v;
  }
  __syncthreads();
  return c;
}

template <typename I, typename N>
__global__ void __launch_bounds__(kNwBlock, 1) rcm_narrow_kernel(RcmArgs<I, N> a) {
  extern __shared__ __align__(16) unsigned char nw_smem_raw[];
  NwSmem<I> &s = *reinterpret_cast<NwSmem<I> *>(nw_smem_raw);
  constexpr int kF = NwSmem<I>::kF;
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s.S = *a.state;
    s.S.status = ST_RUNNING;
    s.in_smem = 0;
  }
  __syncthreads();

  for (;;) {
    const int phase = s.S.phase;  // uniform: S is only written by thread 0 between barriers
    if (phase == PH_DONE) {
      if (threadIdx.x == 0) s.S.status = ST_DONE;
      break;
    }

    // ================================================================ next component
    if (phase == PH_FIND) {
      // rcm_reorder.cc:104-116: scan for the next unvisited vertex; isolated vertices on the
      // way are placed immediately, in index order.
      const int64_t base = s.S.next_i;
      if (base >= a.n || s.S.qwp >= a.n) {  // everything placed: nothing left to scan
        __syncthreads();
        if (threadIdx.x == 0) s.S.phase = PH_DONE;
        __syncthreads();
        continue;
      }
      constexpr int kFindPer = 4;  // consecutive vertices per thread (keeps index order)
      bool unvis[kFindPer], isolated[kFindPer];
      unsigned stopper = 0xffffffffu;
#pragma unroll
      for (int k = 0; k < kFindPer; k++) {
        const int64_t i = base + (int64_t)threadIdx.x * kFindPer + k;
        unvis[k] = false;
        isolated[k] = false;
        if (i < a.n) {
          unvis[k] = __ldcg(&a.mark[i]) != 0u;
          isolated[k] = a.xadj[i] == a.xadj[i + 1];
        }
      }
#pragma unroll
      for (int k = kFindPer - 1; k >= 0; k--)
        if (unvis[k] && !isolated[k]) stopper = threadIdx.x * kFindPer + k;
      // first unvisited, non-isolated vertex of this chunk
      unsigned m = warp_reduce_min(stopper);
      if (lane == 0) s.wtot[wid] = m;
      __syncthreads();
      m = 0xffffffffu;
      for (int w = 0; w < kNwWarps; w++) m = s.wtot[w] < m ? s.wtot[w] : m;
      unsigned take = 0;
#pragma unroll
      for (int k = 0; k < kFindPer; k++)
        take += (unvis[k] && isolated[k] && threadIdx.x * kFindPer + k < m) ? 1u : 0u;
      unsigned total;
      unsigned off = block_exclusive_scan(take, s.scratch, &total);
#pragma unroll
      for (int k = 0; k < kFindPer; k++) {
        if (unvis[k] && isolated[k] && threadIdx.x * kFindPer + k < m) {
          const int64_t i = base + (int64_t)threadIdx.x * kFindPer + k;
          const int64_t pos = s.S.qwp + off++;
          a.Q[pos] = (I)i;
          a.inv[i] = (I)pos;  // singleton component: reversed slice == itself
          atomicExch(&a.mark[i], 0u);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        s.S.qwp += total;
        if (m != 0xffffffffu) {
          s.S.next_i = base + m;  // re-examined (then visited) after the component is done
          s.S.root = base + m;
          s.S.rlevel = -1;
          s.S.qlevel = 0;
          s.S.phase = PH_PBFS_INIT;
          s.S.stat_components++;
        } else {
          s.S.next_i = base + (int64_t)kNwBlock * kFindPer;
        }
      }
      __syncthreads();
      continue;
    }

    // ================================================================ peripheral(): BFS start
    if (phase == PH_PBFS_INIT) {  // rcm_reorder.cc:34-40
      if (threadIdx.x == 0) {
        const I r = (I)s.S.root;
        s.S.rlevel = s.S.qlevel;
        a.Qp[0] = r;
        atomicExch(&a.mark[r], 0u);
        s.S.lvl_begin = 0;
        s.S.lvl_end = 1;
        s.S.prev_begin = 0;
        s.S.depth = 0;
        s.S.phase = PH_PBFS_LEVEL;
        s.S.stat_bfs++;
        s.in_smem = 0;
      }
      __syncthreads();
      continue;
    }

    if (phase == PH_PBFS_LEVEL || phase == PH_CM_LEVEL) {
      const int64_t f = s.S.lvl_end - s.S.lvl_begin;
      if (f == 0) {
        __syncthreads();
        if (threadIdx.x == 0) s.S.phase = phase == PH_PBFS_LEVEL ? PH_PBFS_END : PH_CM_END;
        __syncthreads();
        continue;
      }
      int c = -1;
      if (f <= kF && !a.force_wide)
        c = phase == PH_PBFS_LEVEL ? nw_level<I, N, false>(a, s, a.Qp)
                                   : nw_level<I, N, true>(a, s, a.Q);
      if (c < 0) {
        __syncthreads();
        if (threadIdx.x == 0) s.S.status = ST_NEED_WIDE;
        break;
      }
      if (threadIdx.x == 0) {
        s.S.prev_begin = s.S.lvl_begin;
        s.S.lvl_begin = s.S.lvl_end;
        s.S.lvl_end += c;
        s.S.depth++;
        s.S.stat_levels_narrow++;
        s.in_smem = 1;
      }
      __syncthreads();
      continue;
    }

    // ================================================================ peripheral(): BFS end
    if (phase == PH_PBFS_END) {
      // the BFS visited lvl_end vertices; its last non-empty level is Qp[prev_begin, lvl_begin)
      // at distance depth-1 (rcm_reorder.cc:42-78)
      const int64_t visited = s.S.lvl_end;
      const int64_t ecc = s.S.depth - 1;
      const int64_t qlevel = ecc > s.S.qlevel ? ecc : s.S.qlevel;
      int next;                      // phase after the mark reset
      int64_t new_root = s.S.root;
      if (visited == qlevel + 1) {
        next = PH_CM_INIT;           // :58  path-like component: r is the root
      } else if (s.S.rlevel != qlevel) {
        // :62-78  eccentricity grew: min degree among the last level, first in queue order
        unsigned long long best = ~0ull;
        for (int64_t k = s.S.prev_begin + threadIdx.x; k < s.S.lvl_begin; k += kNwBlock) {
          const I v = a.Qp[k];
          const unsigned long long dg = (unsigned long long)(a.xadj[v + 1] - a.xadj[v]);
          const unsigned long long key = (dg << 32) | (unsigned long long)(k - s.S.prev_begin);
          best = key < best ? key : best;
        }
        best = warp_reduce_min(best);
        if (lane == 0) s.red[wid] = best;
        __syncthreads();
        best = ~0ull;
        for (int w = 0; w < kNwWarps; w++) best = s.red[w] < best ? s.red[w] : best;
        new_root = (int64_t)a.Qp[s.S.prev_begin + (int64_t)(best & 0xffffffffull)];
        next = PH_PBFS_INIT;
      } else {
        next = PH_CM_INIT;           // :34  eccentricity did not grow: keep r
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        s.S.qlevel = qlevel;
        s.S.new_root_pending = new_root;
        s.S.next_phase_after_reset = next;
      }
      __syncthreads();
      // un-visit everything this BFS touched (mark[] doubles as distance[] and V[])
      if (visited > kBulkThreshold) {
        if (threadIdx.x == 0) {
          s.S.phase = PH_PBFS_AFTER_RESET;
          s.S.status = ST_NEED_RESET;
        }
        break;
      }
      for (int64_t k = threadIdx.x; k < visited; k += kNwBlock)
        atomicExch(&a.mark[a.Qp[k]], kUnvisited);
      __syncthreads();
      if (threadIdx.x == 0) s.S.phase = PH_PBFS_AFTER_RESET;
      __syncthreads();
      continue;
    }

    if (phase == PH_PBFS_AFTER_RESET) {
      if (threadIdx.x == 0) {
        s.S.root = s.S.new_root_pending;
        s.S.phase = s.S.next_phase_after_reset;
      }
      __syncthreads();
      continue;
    }

    // ================================================================ Cuthill-McKee BFS
    if (phase == PH_CM_INIT) {  // rcm_reorder.cc:119-123
      if (threadIdx.x == 0) {
        const I r = (I)s.S.root;
        s.S.qst = s.S.qwp;
        a.Q[s.S.qwp] = r;
        atomicExch(&a.mark[r], 0u);
        s.S.lvl_begin = s.S.qwp;
        s.S.lvl_end = s.S.qwp + 1;
        s.S.prev_begin = s.S.qwp;
        s.S.depth = 0;
        s.S.phase = PH_CM_LEVEL;
        s.S.stat_bfs++;
        s.in_smem = 0;
      }
      __syncthreads();
      continue;
    }

    if (phase == PH_CM_END) {
      // component = Q[qst, lvl_end); reversed slice + final inversion (:147-160):
      // inv[Q[k]] = qst + (end-1-k)
      const int64_t qst = s.S.qst, end = s.S.lvl_end;
      if (threadIdx.x == 0) s.S.qwp = end;
      if (end - qst > kBulkThreshold) {
        __syncthreads();
        if (threadIdx.x == 0) {
          s.S.phase = PH_CM_AFTER_INVERT;
          s.S.status = ST_NEED_INVERT;
        }
        break;
      }
      for (int64_t k = qst + threadIdx.x; k < end; k += kNwBlock)
        a.inv[a.Q[k]] = (I)(qst + (end - 1 - k));
      __syncthreads();
      if (threadIdx.x == 0) s.S.phase = PH_CM_AFTER_INVERT;
      __syncthreads();
      continue;
    }

    if (phase == PH_CM_AFTER_INVERT) {
      if (threadIdx.x == 0) s.S.phase = PH_FIND;
      __syncthreads();
      continue;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.state = s.S;
}

// ------------------------------------------------------------------------------------
// WIDE regime (host-driven, one level per call)
// ------------------------------------------------------------------------------------
template <typename I, typename N>
struct FrontierDegFn {
  const I *frontier;
  const N *xadj;
  __device__ int64_t operator()(int64_t i) const {
    const I v = frontier[i];
    return (int64_t)(xadj[v + 1] - xadj[v]);
  }
};

// sweep 1: claims.  key = expansion slot + 1 (peripheral) or parent position + 1 (CM)
template <typename I, typename N, bool CM>
__global__ void __launch_bounds__(256)
    rcm_wide_claim_kernel(const I *__restrict__ frontier, int64_t f, const int64_t *__restrict__ off,
                          const N *__restrict__ xadj, const I *__restrict__ adj,
                          unsigned *__restrict__ mark) {
  const unsigned lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp * 32; g < f; g += nwarps * 32) {
    const int64_t i = g + lane;
    int64_t xs = 0;
    unsigned d = 0;
    if (i < f) {
      const I v = frontier[i];
      xs = (int64_t)xadj[v];
      d = (unsigned)((int64_t)xadj[v + 1] - xs);
    }
    const int64_t slot0 = off[g];
    warp_expand(xs, d, [&](unsigned sl, unsigned j, int64_t p, bool valid) {
      if (valid) {
        const I v = adj[p];
        const unsigned key = CM ? (unsigned)(g + j) + 1u : (unsigned)(slot0 + sl) + 1u;
        if (__ldcg(&mark[v]) != 0u) atomicMin(&mark[v], key);
      }
    });
  }
}

// sweep 2: winners are appended (unordered) with their order key.
//   peripheral: key = slot                      CM: key = (parent position << 32) | degree
template <typename I, typename N, bool CM>
__global__ void __launch_bounds__(256)
    rcm_wide_collect_kernel(const I *__restrict__ frontier, int64_t f,
                            const int64_t *__restrict__ off, const N *__restrict__ xadj,
                            const I *__restrict__ adj, const unsigned *__restrict__ mark,
                            uint64_t *__restrict__ ckey, uint32_t *__restrict__ cval,
                            unsigned long long *__restrict__ counter) {
  const unsigned lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp * 32; g < f; g += nwarps * 32) {
    const int64_t i = g + lane;
    int64_t xs = 0;
    unsigned d = 0;
    if (i < f) {
      const I v = frontier[i];
      xs = (int64_t)xadj[v];
      d = (unsigned)((int64_t)xadj[v + 1] - xs);
    }
    const int64_t slot0 = off[g];
    warp_expand(xs, d, [&](unsigned sl, unsigned j, int64_t p, bool valid) {
      bool win = false;
      I v = 0;
      unsigned key = 0;
      if (valid) {
        v = adj[p];
        key = CM ? (unsigned)(g + j) + 1u : (unsigned)(slot0 + sl) + 1u;
        win = __ldcg(&mark[v]) == key;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, win);
      if (bal) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (win) {
          const unsigned long long k = base + __popc(bal & lanemask_lt());
          uint64_t ok;
          if (CM)
            ok = ((uint64_t)(key - 1u) << 32) | (uint64_t)(unsigned)(xadj[v + 1] - xadj[v]);
          else
            ok = (uint64_t)(key - 1u);
          ckey[k] = ok;
          cval[k] = (uint32_t)v;
        }
      }
    });
  }
}

template <typename I>
__global__ void rcm_wide_commit_kernel(const uint32_t *__restrict__ sorted, int64_t c,
                                       I *__restrict__ queue_out, unsigned *__restrict__ mark) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < c) {
    const uint32_t v = sorted[k];
    queue_out[k] = (I)v;
    mark[v] = 0u;
  }
}

template <typename I>
__global__ void rcm_reset_kernel(const I *__restrict__ q, int64_t cnt, unsigned *__restrict__ mark) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < cnt) mark[q[k]] = kUnvisited;
}

template <typename I>
__global__ void rcm_invert_kernel(const I *__restrict__ Q, int64_t qst, int64_t end,
                                  I *__restrict__ inv) {
  const int64_t k = qst + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < end) inv[Q[k]] = (I)(qst + (end - 1 - k));
}

template <typename N>
__global__ void max_degree_kernel(const N *__restrict__ xadj, int64_t n,
                                  unsigned long long *__restrict__ out) {
  unsigned long long m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long d = (unsigned long long)(xadj[i + 1] - xadj[i]);
    m = d > m ? d : m;
  }
  m = warp_reduce_max(m);
  if (lane_id() == 0 && m) atomicMax(out, m);
}

template <typename I, typename N>
struct WideCtx {
  int64_t *off;
  uint64_t *k0, *k1, *k2;
  uint32_t *v0, *v1, *v2;
  unsigned long long *counter;
  int deg_bits, id_bits;
};

// One BFS level with all SMs.  Reads/updates the host copy of the state.
template <typename I, typename N>
void rcm_wide_level(Workspace &ws, const RcmArgs<I, N> &a, WideCtx<I, N> &w, RcmState &S,
                    bool cm) {
  cudaStream_t st = ws.stream();
  I *queue = cm ? a.Q : a.Qp;
  const I *frontier = queue + S.lvl_begin;
  const int64_t f = S.lvl_end - S.lvl_begin;
  const int grid = device_info(ws.device()).sm_count * 8;
  exclusive_scan<int64_t>(ws, FrontierDegFn<I, N>{frontier, a.xadj}, w.off, f);
  SB_CUDA(cudaMemsetAsync(w.counter, 0, sizeof(unsigned long long), st));
  if (cm) {
    SB_LAUNCH((rcm_wide_claim_kernel<I, N, true>), grid, 256, 0, st, frontier, f,
              (const int64_t *)w.off, a.xadj, a.adj, a.mark);
    SB_LAUNCH((rcm_wide_collect_kernel<I, N, true>), grid, 256, 0, st, frontier, f,
              (const int64_t *)w.off, a.xadj, a.adj, (const unsigned *)a.mark, w.k0, w.v0,
              w.counter);
  } else {
    SB_LAUNCH((rcm_wide_claim_kernel<I, N, false>), grid, 256, 0, st, frontier, f,
              (const int64_t *)w.off, a.xadj, a.adj, a.mark);
    SB_LAUNCH((rcm_wide_collect_kernel<I, N, false>), grid, 256, 0, st, frontier, f,
              (const int64_t *)w.off, a.xadj, a.adj, (const unsigned *)a.mark, w.k0, w.v0,
              w.counter);
  }
  unsigned long long c = 0;
  int64_t slots = 0;
  SB_CUDA(cudaMemcpyAsync(&c, w.counter, sizeof(c), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&slots, w.off + f, sizeof(slots), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  SB_REQUIRE(slots < 0xfffffffell, SB200_ERR_BAD_ARG,
             "RCM level with %lld expansion slots exceeds the 32-bit claim key", (long long)slots);
  if (c > 0) {
    const uint32_t *sorted_v;
    if (cm) {
      // order (parent position, degree, id): LSD = sort by id first, then by (parent, degree)
      radix_sort<uint32_t, uint64_t, NoVal>(ws, {w.v0, w.k0, nullptr}, {w.v1, w.k1, nullptr},
                                            {w.v2, w.k2, nullptr}, (int64_t)c,
                                            {{0, w.id_bits}});
      radix_sort<uint64_t, uint32_t, NoVal>(
          ws, {w.k1, w.v1, nullptr}, {w.k0, w.v0, nullptr}, {w.k2, w.v2, nullptr}, (int64_t)c,
          {{0, w.deg_bits}, {32, 32 + bits_for((uint64_t)(f - 1))}});
      sorted_v = w.v0;
    } else {
      radix_sort<uint64_t, uint32_t, NoVal>(ws, {w.k0, w.v0, nullptr}, {w.k1, w.v1, nullptr},
                                            {w.k2, w.v2, nullptr}, (int64_t)c,
                                            {{0, bits_for((uint64_t)slots)}});
      sorted_v = w.v1;
    }
    SB_LAUNCH((rcm_wide_commit_kernel<I>), (unsigned)ceil_div((int64_t)c, 256), 256, 0, st,
              sorted_v, (int64_t)c, queue + S.lvl_end, a.mark);
  }
  S.prev_begin = S.lvl_begin;
  S.lvl_begin = S.lvl_end;
  S.lvl_end += (int64_t)c;
  S.depth++;
  S.stat_levels_wide++;
}

template <typename I, typename N>
void rcm_impl(Workspace &ws, int64_t n, int64_t nnz, const N *xadj, const I *adj, I *out_inv,
              int force_wide, RcmState *h_stats) {
  if (n <= 0) return;
  cudaStream_t st = ws.stream();
  SB_REQUIRE(n < (1ll << 31), SB200_ERR_BAD_ARG, "RCM supports n < 2^31 (got %lld)", (long long)n);
  RcmArgs<I, N> a;
  a.n = n;
  a.xadj = xadj;
  a.adj = adj;
  a.mark = ws.alloc<unsigned>(n);
  a.Q = ws.alloc<I>(n);
  a.Qp = ws.alloc<I>(n);
  a.inv = out_inv;
  a.state = ws.alloc<RcmState>(1);
  a.force_wide = force_wide;
  SB_CUDA(cudaMemsetAsync(a.mark, 0xff, n * sizeof(unsigned), st));
  RcmState S;
  memset(&S, 0, sizeof(S));
  S.phase = PH_FIND;
  SB_CUDA(cudaMemcpyAsync(a.state, &S, sizeof(S), cudaMemcpyHostToDevice, st));

  WideCtx<I, N> w;
  memset(&w, 0, sizeof(w));
  bool wide_ready = false;
  auto prepare_wide = [&]() {
    if (wide_ready) return;
    w.off = ws.alloc<int64_t>(n + 1);
    w.k0 = ws.alloc<uint64_t>(n);
    w.k1 = ws.alloc<uint64_t>(n);
    w.k2 = ws.alloc<uint64_t>(n);
    w.v0 = ws.alloc<uint32_t>(n);
    w.v1 = ws.alloc<uint32_t>(n);
    w.v2 = ws.alloc<uint32_t>(n);
    w.counter = ws.alloc<unsigned long long>(1);
    unsigned long long *md = ws.alloc<unsigned long long>(1);
    SB_CUDA(cudaMemsetAsync(md, 0, sizeof(*md), st));
    SB_LAUNCH((max_degree_kernel<N>), device_info(ws.device()).sm_count * 8, 256, 0, st, xadj, n,
              md);
    unsigned long long h = 0;
    SB_CUDA(cudaMemcpyAsync(&h, md, sizeof(h), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_REQUIRE(h < (1ull << 32), SB200_ERR_BAD_ARG, "vertex degree %llu exceeds 32 bits", h);
    w.deg_bits = bits_for(h);
    w.id_bits = bits_for((uint64_t)(n - 1));
    wide_ready = true;
  };

  auto kern = rcm_narrow_kernel<I, N>;
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(NwSmem<I>)));
  for (;;) {
    SB_LAUNCH(kern, 1, kNwBlock, sizeof(NwSmem<I>), st, a);
    SB_CUDA(cudaMemcpyAsync(&S, a.state, sizeof(S), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (S.status == ST_DONE) break;
    if (S.status == ST_NEED_WIDE) {
      prepare_wide();
      const bool cm = S.phase == PH_CM_LEVEL;
      // keep going wide while the frontier is far beyond the narrow capacity
      do {
        rcm_wide_level<I, N>(ws, a, w, S, cm);
      } while (S.lvl_end - S.lvl_begin > (force_wide ? 0 : 4 * NwCaps<I>::kF));
    } else if (S.status == ST_NEED_RESET) {
      const int64_t cnt = S.lvl_end;
      SB_LAUNCH((rcm_reset_kernel<I>), (unsigned)ceil_div(cnt, 256), 256, 0, st, (const I *)a.Qp,
                cnt, a.mark);
    } else if (S.status == ST_NEED_INVERT) {
      const int64_t cnt = S.lvl_end - S.qst;
      SB_LAUNCH((rcm_invert_kernel<I>), (unsigned)ceil_div(cnt, 256), 256, 0, st, (const I *)a.Q,
                S.qst, S.lvl_end, a.inv);
    } else {
      SB_REQUIRE(false, SB200_ERR_INTERNAL, "RCM state machine returned status %d in phase %d",
                 S.status, S.phase);
    }
    S.status = ST_RUNNING;
    SB_CUDA(cudaMemcpyAsync(a.state, &S, sizeof(S), cudaMemcpyHostToDevice, st));
  }
  if (h_stats) *h_stats = S;
}

}  // namespace sb200

using namespace sb200;

static thread_local RcmState g_last_rcm_stats;

extern "C" {

int sb200_rcm_reorder(int device, int64_t n, int64_t nnz, const void *row_ptr, const void *col,
                      void *out_inv, int id_type, int nnz_type, void *stream) {
  return guarded(device, [&] {
    SB_REQUIRE(n >= 0 && nnz >= 0 && (n == 0 || (row_ptr && out_inv)), SB200_ERR_BAD_ARG,
               "bad argument");
    SB_REQUIRE(nnz == 0 || col, SB200_ERR_BAD_ARG, "col is null");
    Workspace ws(device, (cudaStream_t)stream);
    const char *fw = getenv("SB200_RCM_FORCE_WIDE");
    const int force_wide = fw && fw[0] == '1';
    dispatch_inv(id_type, nnz_type, SB200_VOID, false, [&](auto I_, auto N_, auto) {
      using I = decltype(I_);
      using N = decltype(N_);
      rcm_impl<I, N>(ws, n, nnz, (const N *)row_ptr, (const I *)col, (I *)out_inv, force_wide,
                     &g_last_rcm_stats);
    });
  });
}

// Diagnostics of the last sb200_rcm_reorder call on this thread:
// out[0..3] = levels done by the persistent CTA, levels done wide, BFS count, components.
int sb200_rcm_last_stats(int64_t *h_out4) {
  if (!h_out4) return SB200_ERR_BAD_ARG;
  h_out4[0] = g_last_rcm_stats.stat_levels_narrow;
  h_out4[1] = g_last_rcm_stats.stat_levels_wide;
  h_out4[2] = g_last_rcm_stats.stat_bfs;
  h_out4[3] = g_last_rcm_stats.stat_components;
  return SB200_OK;
}

}  // extern "C"
