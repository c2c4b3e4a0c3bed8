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
constexpr int kNwBlock = 512;
constexpr int kNwWarps = kNwBlock / 32;
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
  int32_t spec;  // 1: the running CM BFS stands in for a BFS of peripheral() (PH_PBFS_END)
  int64_t frontier_maxdeg;  // max degree over the current frontier (for the narrow caps)
  int64_t key_base;    // peripheral BFS: expansion slots of all earlier levels (monotone claim keys)
  int64_t inv_qst, inv_end;  // pending bulk inversion
  int64_t reset_cm;    // pending bulk reset walks Q[qst, lvl_end) instead of Qp[0, lvl_end)
  int64_t stat_spec_ok, stat_spec_fail, stat_spec_chain;
  int64_t stat_levels_narrow, stat_levels_wide, stat_bfs, stat_components;
  int64_t cyc[8];  // narrow-level phase cycle counters (CTA 0): load, claim, check, finalize, write
};

__device__ __forceinline__ int bits_for_dev(unsigned long long v) {
  return v ? 64 - __clzll((long long)v) : 0;
}

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
  int no_spec;     // testing: never run a peripheral BFS speculatively as the CM BFS
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
// NARROW regime: one thread-block CLUSTER walks the levels.
//
// A single SM cannot issue the ~3 scattered L2 accesses per edge of a 4096-wide level fast
// enough (measured: 36 us per level, LSU-bound), so the level is split over the CTAs of one
// cluster (16 where the device allows it, else 8).  Every CTA keeps ITS share of the frontier
// in shared memory from one level to the next: the vertices it discovers (already sorted,
// with their adjacency extents) are its share of the next level, so a level needs no global
// re-read and only two cluster barriers:
//     claims (atomicMin in L2)            -> barrier 1 ->
//     recheck, compaction, sibling sort   -> exchange of the per-CTA counts (barrier 2)
// The global queue is still written every level (it is the result), but nobody waits for it.
// Shares drift apart over time (a BFS starts with everything in CTA 0); when the largest
// share exceeds twice the even share, or shared memory, the CTAs re-split the frontier evenly
// from the global queue (cl_reload, one more barrier).
// All CTAs run the same state machine on replicated state, so control flow is uniform.
// ------------------------------------------------------------------------------------
constexpr int kClMax = 16;
// expansion slots (and provisional / final winners) per CTA and level; 8-byte ids get fewer so
// that the staging arrays stay inside the 227 KB of shared memory
constexpr int kEl = 4096;
constexpr int kEl64 = 3072;
template <typename I>
constexpr int cl_el() {
  return sizeof(I) == 4 ? kEl : kEl64;
}
constexpr int kFl = 1024;  // frontier vertices per CTA and level
constexpr int kClBatch = 4;  // 32-slot rounds whose loads are in flight together

template <typename I>
struct ClSmem {
  // this CTA's share of the current frontier: vertex, adjacency start, degree
  I F[kFl];
  int64_t Fx[kFl];
  unsigned Fd[kFl];
  // provisional winners of sweep 1, one region per warp (region base = prefix of wslot[])
  I pv[cl_el<I>()];
  unsigned pkey[cl_el<I>()];  // claim key; 0 after sweep 2 when the claim lost
  unsigned pi[cl_el<I>()];    // local index of the parent in F
  int64_t px[cl_el<I>()];     // adjacency start / degree of the claimed vertex (filled by sweep 2)
  unsigned pd[cl_el<I>()];
  // final winners in slot order
  I cv[cl_el<I>()];
  unsigned ci[cl_el<I>()];
  unsigned cd[cl_el<I>()];
  int64_t cx[cl_el<I>()];
  unsigned long long xch[2][kClMax][4];
  unsigned long long mine[4];
  unsigned wtot[kNwWarps + 2];
  unsigned wslot[kNwWarps];  // expansion slots of each warp's part of the share
  unsigned wprov[kNwWarps];
  unsigned wwin[kNwWarps];
  unsigned wmax[kNwWarps];
  unsigned wsum[kNwWarps];
  unsigned scratch[34];
  unsigned long long red[kNwWarps];
  int fl;        // share size
  long long fb;  // global frontier position of F[0]
  unsigned sbase;  // expansion slots of the shares of the lower-ranked CTAs (this level)
  // replicated: largest number of expansion slots of any share, slots of the whole level;
  // need_reload = 1: the shares must be re-split from the global queue
  unsigned long long max_slots;
  unsigned long long level_slots;
  int need_reload;
  RcmState S;
};

}  // namespace sb200
#include <cooperative_groups.h>
namespace sb200 {
namespace cg = cooperative_groups;

// Every CTA contributes s.mine[0..3]; afterwards s.xch[par][k][*] holds CTA k's values in every
// CTA.  Contains a cluster barrier.  `par` alternates so that a fast CTA never overwrites
// values a slow CTA has not read yet.
template <typename I>
__device__ __forceinline__ void cl_exchange(cg::cluster_group &cluster, ClSmem<I> &s, int &par) {
  __syncthreads();
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  if (threadIdx.x < C) {
    unsigned long long *dst = cluster.map_shared_rank(&s.xch[par][rank][0], threadIdx.x);
    dst[0] = s.mine[0];
    dst[1] = s.mine[1];
    dst[2] = s.mine[2];
    dst[3] = s.mine[3];
  }
  cluster.sync();
  par ^= 1;
}

// Warp w expands the share's vertices [fl*w/W, fl*(w+1)/W).
__device__ __forceinline__ int cl_part_begin(int fl, unsigned w) {
  return (int)(((long long)fl * w) / kNwWarps);
}

// Per-warp expansion-slot totals of the share F[0..fl) -> s.wslot[]; returns the CTA total.
// Contains __syncthreads().
template <typename I>
__device__ __forceinline__ unsigned cl_count_slots(ClSmem<I> &s, int fl) {
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  unsigned sum = 0;
  for (int i = cl_part_begin(fl, wid) + (int)lane; i < cl_part_begin(fl, wid + 1); i += 32)
    sum += s.Fd[i];
  sum = warp_reduce_sum(sum);
  if (lane == 0) s.wslot[wid] = sum;
  __syncthreads();
  unsigned total = 0;
#pragma unroll
  for (int w = 0; w < kNwWarps; w++) total += s.wslot[w];
  return total;
}

// Re-split the frontier queue[lvl_begin, lvl_end) evenly over the CTAs (also the way a BFS
// starts and the way the kernel resumes after a host-driven step).  Contains cluster barriers.
template <typename I, typename N>
__device__ void cl_reload(cg::cluster_group &cluster, const RcmArgs<I, N> &a, ClSmem<I> &s,
                          const I *queue, int &par) {
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  cluster.sync();  // every CTA's queue writes of the previous level are visible
  const int64_t f = s.S.lvl_end - s.S.lvl_begin;
  const int64_t max_share = (f + C - 1) / C;
  int fl = 0;
  long long fb = 0;
  unsigned total = 0;
  if (max_share <= kFl) {
    fb = f * rank / C;
    fl = (int)(f * (rank + 1) / C - fb);
    for (int i = threadIdx.x; i < fl; i += kNwBlock) {
      const I v = __ldcg(queue + s.S.lvl_begin + fb + i);
      const int64_t xs = (int64_t)a.xadj[v];
      s.F[i] = v;
      s.Fx[i] = xs;
      s.Fd[i] = (unsigned)((int64_t)a.xadj[v + 1] - xs);
    }
    __syncthreads();
    total = cl_count_slots(s, fl);
  }
  if (threadIdx.x == 0) {
    s.fl = fl;
    s.fb = fb;
    s.mine[0] = max_share <= kFl ? total : (unsigned long long)cl_el<I>() + 1;  // too wide: go wide
    s.mine[1] = s.mine[2] = s.mine[3] = 0;
  }
  cl_exchange(cluster, s, par);
  unsigned long long mx = 0, sum = 0, below = 0;
  for (unsigned k = 0; k < C; k++) {
    const unsigned long long tk = s.xch[par ^ 1][k][0];
    mx = tk > mx ? tk : mx;
    sum += tk;
    below += k < rank ? tk : 0ull;
  }
  if (threadIdx.x == 0) {
    s.max_slots = mx;
    s.level_slots = sum;
    s.sbase = (unsigned)below;
    s.need_reload = 0;
  }
  __syncthreads();
}

// One BFS level across the cluster.  Returns the number of newly reached vertices (cluster
// total), or -1 when the level has to be done by the wide path (nothing modified then).
// On return *next_maxdeg holds the maximum degree among the new vertices.
//
//   sweep 1   every warp walks the concatenated adjacency lists of its part of the share 32
//             slots at a time (kClBatch rounds in flight), claims with atomicMin and appends
//             the claims that were the minimum when they landed to its provisional list
//   barrier 1 all claims of the level have landed
//   sweep 2   every warp re-reads the marks of its provisional claims (a claim survived iff
//             the mark still equals its key), fetches the winners' adjacency extents in the
//             same round trip and marks them visited
//   compaction into slot order, counts exchanged through distributed shared memory (barrier 2)
//   ordering  slot order (peripheral) or per-parent (degree, id) order (CM), written to the
//             global queue and -- unless the shares drifted apart -- kept as the next share
template <typename I, typename N, bool CM>
__device__ long long cl_level(cg::cluster_group &cluster, const RcmArgs<I, N> &a, ClSmem<I> &s,
                              I *queue, int &par, unsigned cur_maxdeg, unsigned *next_maxdeg) {
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  const int64_t f = s.S.lvl_end - s.S.lvl_begin;
  // uniform feasibility checks (every CTA evaluates the same replicated numbers)
  if (s.max_slots > (unsigned long long)cl_el<I>()) return -1;
  if (CM && cur_maxdeg > (unsigned)kNwGroupCap) return -1;
  // Claim keys grow from level to level (CM: queue position of the parent; peripheral: running
  // expansion-slot number of the BFS), so a vertex reached earlier always holds a SMALLER mark
  // than any later claim: visited vertices need no separate "visited" store, and barrier 2 has
  // no global writes to wait for.
  if (!CM && (unsigned long long)s.S.key_base + s.level_slots >= 0xfffffff0ull) return -1;
  const unsigned kb = CM ? (unsigned)s.S.lvl_begin : (unsigned)s.S.key_base + s.sbase;
  (void)f;
  const int fl = s.fl;
  const long long fb = s.fb;
  long long t0 = clock64(), t1;
#define SB_TICK(slot)                                   \
  do {                                                  \
    t1 = clock64();                                     \
    if (threadIdx.x == 0) s.S.cyc[slot] += t1 - t0;     \
    t0 = t1;                                            \
  } while (0)

  unsigned pbase = 0;  // my warp's provisional region
#pragma unroll
  for (int w = 0; w < kNwWarps; w++) pbase += (unsigned)w < wid ? s.wslot[w] : 0u;
  unsigned pcount = 0;
  unsigned goff = pbase;  // expansion-slot number (inside this CTA's share) of the group's slot 0
  SB_TICK(0);

  // ---- sweep 1: claims.  key orders (global frontier position[, adjacency index]) ----
  const int vb = cl_part_begin(fl, wid), ve = cl_part_begin(fl, wid + 1);
  for (int g = vb; g < ve; g += 32) {
    const int i = g + (int)lane;
    int64_t xs = 0;
    unsigned d = 0;
    if (i < ve) {
      xs = s.Fx[i];
      d = s.Fd[i];
    }
    const unsigned incl = warp_inclusive_scan(d);
    const unsigned excl = incl - d;
    const unsigned tot = __shfl_sync(0xffffffffu, incl, 31);
    for (unsigned b0 = 0; b0 < tot; b0 += 32 * kClBatch) {
      I v[kClBatch];
      unsigned key[kClBatch], own[kClBatch], old[kClBatch];
      bool valid[kClBatch];
#pragma unroll
      for (int u = 0; u < kClBatch; u++) {
        const unsigned sl = b0 + u * 32 + lane;
        unsigned lo = 0;  // number of lanes whose inclusive end <= sl  == owner lane
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const unsigned val = __shfl_sync(0xffffffffu, incl, (lo + step - 1) & 31);
          if (val <= sl) lo += step;
        }
        const unsigned j = lo & 31;
        const int64_t xs_j = __shfl_sync(0xffffffffu, xs, j);
        const unsigned excl_j = __shfl_sync(0xffffffffu, excl, j);
        valid[u] = sl < tot;
        own[u] = (unsigned)g + j;
        key[u] = CM ? kb + (unsigned)(fb + g + j) + 1u : kb + goff + sl + 1u;
        if (valid[u]) v[u] = a.adj[xs_j + (int64_t)(sl - excl_j)];
      }
#pragma unroll
      for (int u = 0; u < kClBatch; u++)
        if (valid[u]) old[u] = atomicMin(&a.mark[v[u]], key[u]);
#pragma unroll
      for (int u = 0; u < kClBatch; u++) {
        const bool prov = valid[u] && old[u] > key[u];
        const unsigned bal = __ballot_sync(0xffffffffu, prov);
        if (prov) {
          const unsigned at = pbase + pcount + __popc(bal & lanemask_lt());
          s.pv[at] = v[u];
          s.pkey[at] = key[u];
          s.pi[at] = own[u];
        }
        pcount += __popc(bal);
      }
    }
    goff += tot;
  }
  SB_TICK(1);
  cluster.sync();
  SB_TICK(2);

  // ---- sweep 2: which provisional claims survived?  (warp-local lists) ----
  unsigned wins = 0, mymax = 0, mysum = 0;
  for (unsigned k0 = 0; k0 < pcount; k0 += 32 * kClBatch) {
    I v[kClBatch];
    unsigned m[kClBatch];
    N xs[kClBatch], xe[kClBatch];
#pragma unroll
    for (int u = 0; u < kClBatch; u++) {
      const unsigned k = k0 + u * 32 + lane;
      if (k < pcount) {
        v[u] = s.pv[pbase + k];
        m[u] = __ldcg(&a.mark[v[u]]);
        xs[u] = a.xadj[v[u]];
        xe[u] = a.xadj[v[u] + 1];
      }
    }
#pragma unroll
    for (int u = 0; u < kClBatch; u++) {
      const unsigned k = k0 + u * 32 + lane;
      if (k < pcount) {
        if (m[u] == s.pkey[pbase + k]) {
          const unsigned dg = (unsigned)(xe[u] - xs[u]);
          // the winner is expanded in the next level, after two cluster barriers and the
          // ordering: pull its adjacency into L2 meanwhile (the lists are read once per
          // traversal, so the expansion would otherwise wait for HBM)
          if (dg > 0) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + xs[u]));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + xe[u] - 1));
          }
          s.px[pbase + k] = (int64_t)xs[u];
          s.pd[pbase + k] = dg;
          mymax = dg > mymax ? dg : mymax;
          mysum += dg;
          wins++;
        } else {
          s.pkey[pbase + k] = 0u;
        }
      }
    }
  }
  wins = warp_reduce_sum(wins);
  mymax = warp_reduce_max(mymax);
  mysum = warp_reduce_sum(mysum);
  if (lane == 0) {
    s.wwin[wid] = wins;
    s.wmax[wid] = mymax;
    s.wsum[wid] = mysum;
  }
  SB_TICK(3);
  __syncthreads();
  // ---- ordered compaction of the winners (warp regions are already in slot order) ----
  unsigned wb = 0, c_local = 0;
#pragma unroll
  for (int w = 0; w < kNwWarps; w++) {
    const unsigned cw = s.wwin[w];
    wb += (unsigned)w < wid ? cw : 0u;
    c_local += cw;
  }
  for (unsigned k0 = 0; k0 < pcount; k0 += 32) {
    const unsigned k = k0 + lane;
    const bool win = k < pcount && s.pkey[pbase + k] != 0u;
    const unsigned bal = __ballot_sync(0xffffffffu, win);
    if (win) {
      const unsigned at = wb + __popc(bal & lanemask_lt());
      s.cv[at] = s.pv[pbase + k];
      s.ci[at] = s.pi[pbase + k];
      s.cx[at] = s.px[pbase + k];
      s.cd[at] = s.pd[pbase + k];
    }
    wb += __popc(bal);
  }
  if (threadIdx.x == 0) {
    unsigned mx = 0, sm = 0;
#pragma unroll
    for (int w = 0; w < kNwWarps; w++) {
      mx = s.wmax[w] > mx ? s.wmax[w] : mx;
      sm += s.wsum[w];
    }
    s.mine[0] = c_local;
    s.mine[1] = mx;
    s.mine[2] = sm;
    s.mine[3] = 0;
  }
  SB_TICK(4);
  cl_exchange(cluster, s, par);
  const int c = (int)c_local;
  // every warp reduces the C records with its first C lanes
  unsigned ck = 0, mk = 0, sk = 0;
  if (lane < C) {
    ck = (unsigned)s.xch[par ^ 1][lane][0];
    mk = (unsigned)s.xch[par ^ 1][lane][1];
    sk = (unsigned)s.xch[par ^ 1][lane][2];
  }
  const long long cbase = warp_reduce_sum(lane < rank ? ck : 0u);
  const long long ctotal = warp_reduce_sum(ck);
  const long long cmax = warp_reduce_max(ck);
  const unsigned nmax = warp_reduce_max(mk);
  const unsigned long long smax = warp_reduce_max(sk);
  const unsigned sbase_next = warp_reduce_sum(lane < rank ? sk : 0u);
  const unsigned long long slots_next = warp_reduce_sum(sk);
  *next_maxdeg = nmax;
  // keep the shares where they are unless they have drifted too far apart (uniform decision)
  const bool keep = cmax <= kFl && smax <= (unsigned long long)cl_el<I>() &&
                    cmax <= 2 * ((ctotal + C - 1) / C) + 32;
  SB_TICK(5);

  // ---- next frontier: slot order (peripheral) or (parent, degree, id) order (CM); the
  //      sibling groups of a parent never leave the CTA that owns the parent ----
  const int64_t out0 = s.S.lvl_end + cbase;
  for (int k = threadIdx.x; k < c; k += kNwBlock) {
    int dst = k;
    const I w = s.cv[k];
    const unsigned dg = s.cd[k];
    if (CM) {
      const unsigned par_i = s.ci[k];
      int rank_in = 0, left = 0;
      for (int q = k - 1; q >= 0 && s.ci[q] == par_i; q--) {
        left++;
        rank_in += (s.cd[q] < dg || (s.cd[q] == dg && s.cv[q] < w)) ? 1 : 0;
      }
      for (int q = k + 1; q < c && s.ci[q] == par_i; q++)
        rank_in += (s.cd[q] < dg || (s.cd[q] == dg && s.cv[q] < w)) ? 1 : 0;
      dst = k - left + rank_in;
    }
    queue[out0 + dst] = w;
    if (keep) {  // c <= kFl
      s.F[dst] = w;
      s.Fx[dst] = s.cx[k];
      s.Fd[dst] = dg;
    }
    // the next level starts by reading this vertex's adjacency list: pull it towards L2 now
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + s.cx[k]));
  }
  if (threadIdx.x == 0) {
    if (!CM) s.S.key_base += (int64_t)s.level_slots;  // this level's slots are used up
    if (keep) {
      s.fl = c;
      s.fb = cbase;
      s.sbase = sbase_next;
      s.max_slots = smax;
      s.level_slots = slots_next;
    } else {
      s.need_reload = 1;
    }
  }
  __syncthreads();
  SB_TICK(6);
  if (keep) cl_count_slots(s, c);  // ends with a barrier
  SB_TICK(7);
#undef SB_TICK
  return ctotal;
}

// min over queue[b, e) of (degree << 32 | position - b); *ties = how many vertices of the slice
// have that minimum degree.  Every CTA scans the whole slice (it is one BFS level, and this
// keeps the replicated state identical).  Contains __syncthreads().
template <typename I, typename N>
__device__ unsigned long long cl_last_level_min(const RcmArgs<I, N> &a, ClSmem<I> &s,
                                                const I *queue, int64_t b, int64_t e,
                                                unsigned *ties) {
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  unsigned long long best = ~0ull;
  for (int64_t k = b + threadIdx.x; k < e; k += kNwBlock) {
    const I v = __ldcg(queue + k);
    const unsigned long long dg = (unsigned long long)(a.xadj[v + 1] - a.xadj[v]);
    const unsigned long long key = (dg << 32) | (unsigned long long)(k - b);
    best = key < best ? key : best;
  }
  best = warp_reduce_min(best);
  __syncthreads();
  if (lane == 0) s.red[wid] = best;
  __syncthreads();
  best = ~0ull;
  for (int w = 0; w < kNwWarps; w++) best = s.red[w] < best ? s.red[w] : best;
  const unsigned long long mind = best >> 32;
  unsigned cnt = 0;
  for (int64_t k = b + threadIdx.x; k < e; k += kNwBlock) {
    const I v = __ldcg(queue + k);
    cnt += (unsigned long long)(a.xadj[v + 1] - a.xadj[v]) == mind ? 1u : 0u;
  }
  cnt = warp_reduce_sum(cnt);
  if (lane == 0) s.wtot[wid] = cnt;
  __syncthreads();
  cnt = 0;
  for (int w = 0; w < kNwWarps; w++) cnt += s.wtot[w];
  __syncthreads();
  *ties = cnt;
  return best;
}

template <typename I, typename N>
__global__ void __launch_bounds__(kNwBlock, 1) rcm_narrow_kernel(RcmArgs<I, N> a) {
  extern __shared__ __align__(16) unsigned char nw_smem_raw[];
  ClSmem<I> &s = *reinterpret_cast<ClSmem<I> *>(nw_smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  int par = 0;
  unsigned cur_maxdeg = 0;  // max degree over the current frontier (replicated)
  if (threadIdx.x == 0) {
    s.S = *a.state;
    s.S.status = ST_RUNNING;
    s.fl = 0;
    s.fb = 0;
    s.max_slots = 0;
    s.need_reload = 1;  // nothing is resident yet: the first level splits the queue
  }
  __syncthreads();
  cur_maxdeg = (unsigned)s.S.frontier_maxdeg;

  for (;;) {
    const int phase = s.S.phase;  // replicated state: identical in every CTA
    if (phase == PH_DONE) {
      if (threadIdx.x == 0) s.S.status = ST_DONE;
      break;
    }

    // ================================================================ next component
    if (phase == PH_FIND) {
      // rcm_reorder.cc:104-116: scan for the next unvisited vertex; isolated vertices on the
      // way are placed immediately, in index order.  CTA 0 scans, the result is broadcast.
      const int64_t base = s.S.next_i;
      if (base >= a.n || s.S.qwp >= a.n) {  // everything placed: nothing left to scan
        __syncthreads();
        if (threadIdx.x == 0) s.S.phase = PH_DONE;
        __syncthreads();
        continue;
      }
      constexpr int kFindPer = 4;  // consecutive vertices per thread (keeps index order)
      if (rank == 0) {
        bool unvis[kFindPer], isolated[kFindPer];
        unsigned stopper = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < kFindPer; k++) {
          const int64_t i = base + (int64_t)threadIdx.x * kFindPer + k;
          unvis[k] = false;
          isolated[k] = false;
          if (i < a.n) {
            unvis[k] = __ldcg(&a.mark[i]) == kUnvisited;
            isolated[k] = a.xadj[i] == a.xadj[i + 1];
          }
        }
#pragma unroll
        for (int k = kFindPer - 1; k >= 0; k--)
          if (unvis[k] && !isolated[k]) stopper = threadIdx.x * kFindPer + k;
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
        if (threadIdx.x == 0) {
          s.mine[0] = m;
          s.mine[1] = total;
        }
      }
      cl_exchange(cluster, s, par);
      const unsigned m = (unsigned)s.xch[par ^ 1][0][0];
      const unsigned total = (unsigned)s.xch[par ^ 1][0][1];
      if (threadIdx.x == 0) {
        s.S.qwp += total;
        if (m != 0xffffffffu) {
          s.S.next_i = base + m;  // re-examined (then visited) after the component is done
          s.S.root = base + m;
          s.S.rlevel = -1;
          s.S.qlevel = 0;
          s.S.phase = PH_PBFS_INIT;
          s.S.spec = 0;
          s.S.stat_components++;
        } else {
          s.S.next_i = base + (int64_t)kNwBlock * kFindPer;
        }
      }
      __syncthreads();
      continue;
    }

    // ================================================================ BFS start
    if (phase == PH_PBFS_INIT || phase == PH_CM_INIT) {  // rcm_reorder.cc:34-40 / :119-123
      const bool cm = phase == PH_CM_INIT;
      const I r = (I)s.S.root;
      const int64_t at = cm ? s.S.qwp : 0;
      if (rank == 0 && threadIdx.x == 0) {
        (cm ? a.Q : a.Qp)[at] = r;
        atomicExch(&a.mark[r], 0u);
      }
      cur_maxdeg = (unsigned)((int64_t)a.xadj[r + 1] - (int64_t)a.xadj[r]);
      if (threadIdx.x == 0) {
        if (cm) {
          s.S.qst = s.S.qwp;
        } else {
          s.S.rlevel = s.S.qlevel;
          s.S.key_base = 0;
        }
        s.S.lvl_begin = at;
        s.S.lvl_end = at + 1;
        s.S.prev_begin = at;
        s.S.depth = 0;
        s.S.phase = cm ? PH_CM_LEVEL : PH_PBFS_LEVEL;
        s.S.stat_bfs++;
        s.need_reload = 1;  // the reload's barrier publishes the root to every CTA
      }
      __syncthreads();
      continue;
    }

    if (phase == PH_PBFS_LEVEL || phase == PH_CM_LEVEL) {
      const int64_t f = s.S.lvl_end - s.S.lvl_begin;
      if (f == 0) {
        cluster.sync();  // the queue is complete: the END phases read other CTAs' slices
        if (threadIdx.x == 0) s.S.phase = phase == PH_PBFS_LEVEL ? PH_PBFS_END : PH_CM_END;
        __syncthreads();
        continue;
      }
      long long c = -1;
      unsigned next_maxdeg = 0;
      if (!a.force_wide && s.need_reload)
        cl_reload<I, N>(cluster, a, s, phase == PH_PBFS_LEVEL ? a.Qp : a.Q, par);
      if (!a.force_wide)
        c = phase == PH_PBFS_LEVEL
                ? cl_level<I, N, false>(cluster, a, s, a.Qp, par, cur_maxdeg, &next_maxdeg)
                : cl_level<I, N, true>(cluster, a, s, a.Q, par, cur_maxdeg, &next_maxdeg);
      if (c < 0) {
        __syncthreads();
        if (threadIdx.x == 0) s.S.status = ST_NEED_WIDE;
        break;
      }
      cur_maxdeg = next_maxdeg;
      if (threadIdx.x == 0) {
        s.S.prev_begin = s.S.lvl_begin;
        s.S.lvl_begin = s.S.lvl_end;
        s.S.lvl_end += c;
        s.S.depth++;
        s.S.stat_levels_narrow++;
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
      int next;  // phase after the mark reset
      int64_t new_root = s.S.root;
      bool speculate = false;
      if (visited == qlevel + 1) {
        next = PH_CM_INIT;  // :58  path-like component: r is the root
      } else if (s.S.rlevel != qlevel) {
        // :62-78  eccentricity grew: min degree among the last level, first in queue order
        // (every CTA scans the whole level: it is short and this keeps the state replicated)
        unsigned ties;
        const unsigned long long best =
            cl_last_level_min<I, N>(a, s, a.Qp, s.S.prev_begin, s.S.lvl_begin, &ties);
        new_root = (int64_t)a.Qp[s.S.prev_begin + (int64_t)(best & 0xffffffffull)];
        // The next BFS of peripheral() starts from new_root.  If its eccentricity does not
        // exceed qlevel, peripheral() returns new_root and the Cuthill-McKee BFS walks the same
        // component from the same root: both traversals have the same level SETS, so the CM
        // BFS alone gives the eccentricity.  Run it first, speculatively; PH_CM_END checks the
        // outcome and, when the eccentricity did grow, continues the literal sequence.
        speculate = !a.no_spec;
        next = speculate ? PH_CM_INIT : PH_PBFS_INIT;
      } else {
        next = PH_CM_INIT;  // :34  eccentricity did not grow: keep r
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        s.S.spec = speculate ? 1 : 0;
        if (speculate) s.S.rlevel = qlevel;  // :35 of the iteration the CM BFS stands in for
        s.S.qlevel = qlevel;
        s.S.new_root_pending = new_root;
        s.S.next_phase_after_reset = next;
        s.S.phase = PH_PBFS_AFTER_RESET;
        s.S.reset_cm = 0;
      }
      __syncthreads();
      // un-visit everything this BFS touched (mark[] doubles as distance[] and V[])
      if (visited > kBulkThreshold) {
        if (threadIdx.x == 0) s.S.status = ST_NEED_RESET;
        break;
      }
      for (int64_t k = (int64_t)rank * kNwBlock + threadIdx.x; k < visited;
           k += (int64_t)C * kNwBlock)
        atomicExch(&a.mark[a.Qp[k]], kUnvisited);
      cluster.sync();
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

    // ================================================================ Cuthill-McKee end
    if (phase == PH_CM_END) {
      // component = Q[qst, lvl_end); reversed slice + final inversion (:147-160):
      // inv[Q[k]] = qst + (end-1-k)
      const int64_t qst = s.S.qst, end = s.S.lvl_end;
      if (s.S.spec == 1) {
        // this CM BFS stood in for a BFS of peripheral() (see PH_PBFS_END): rcm_reorder.cc:42-59
        const int64_t visited = end - qst, ecc = s.S.depth - 1;
        const int64_t qlevel = ecc > s.S.qlevel ? ecc : s.S.qlevel;
        const bool confirmed = visited == qlevel + 1 || qlevel == s.S.rlevel;
        if (!confirmed) {
          // The eccentricity grew: peripheral() picks the min-degree vertex of the last level,
          // FIRST IN FIFO ORDER on ties (:64-76).  The CM order of the level differs from the
          // FIFO order, but the level SET is the same: when one vertex alone has the minimum
          // degree it is the new root whatever the order, and the search continues (again
          // speculatively) from it.  On a tie the FIFO order is needed: replay this BFS literally.
          unsigned ties;
          const unsigned long long best =
              cl_last_level_min<I, N>(a, s, a.Q, s.S.prev_begin, s.S.lvl_begin, &ties);
          const int64_t new_root = (int64_t)a.Q[s.S.prev_begin + (int64_t)(best & 0xffffffffull)];
          __syncthreads();
          if (threadIdx.x == 0) {
            if (ties == 1) {
              s.S.stat_spec_chain++;
              s.S.qlevel = qlevel;
              s.S.rlevel = qlevel;
              s.S.new_root_pending = new_root;
              s.S.next_phase_after_reset = PH_CM_INIT;  // spec stays 1
            } else {
              s.S.stat_spec_fail++;
              s.S.spec = 0;
              s.S.new_root_pending = s.S.root;  // same root; rlevel/qlevel as before
              s.S.next_phase_after_reset = PH_PBFS_INIT;
            }
            s.S.phase = PH_PBFS_AFTER_RESET;
            s.S.reset_cm = 1;
          }
          __syncthreads();
          if (visited > kBulkThreshold) {
            if (threadIdx.x == 0) s.S.status = ST_NEED_RESET;
            break;
          }
          for (int64_t k = qst + (int64_t)rank * kNwBlock + threadIdx.x; k < end;
               k += (int64_t)C * kNwBlock)
            atomicExch(&a.mark[a.Q[k]], kUnvisited);
          cluster.sync();
          continue;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          s.S.spec = 0;
          s.S.stat_spec_ok++;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        s.S.qwp = end;
        s.S.phase = PH_FIND;
      }
      __syncthreads();
      if (end - qst > kBulkThreshold) {
        if (threadIdx.x == 0) s.S.status = ST_NEED_INVERT;
        break;
      }
      for (int64_t k = qst + (int64_t)rank * kNwBlock + threadIdx.x; k < end;
           k += (int64_t)C * kNwBlock)
        a.inv[a.Q[k]] = (I)(qst + (end - 1 - k));
      cluster.sync();
      continue;
    }
  }
  __syncthreads();
  if (rank == 0 && threadIdx.x == 0) {
    s.S.frontier_maxdeg = cur_maxdeg;
    *a.state = s.S;
  }
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
  // counter[0] = number of winners, counter[1] = max degree among them
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
          const unsigned dg = (unsigned)(xadj[v + 1] - xadj[v]);
          ckey[k] = CM ? (((uint64_t)(key - 1u) << 32) | (uint64_t)dg) : (uint64_t)(key - 1u);
          cval[k] = (uint32_t)v;
          atomicMax(counter + 1, (unsigned long long)dg);
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
__global__ void rcm_reset_kernel(const I *__restrict__ q, int64_t cnt, unsigned *__restrict__ mark,
                                 unsigned value) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < cnt) mark[q[k]] = value;
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
  SB_CUDA(cudaMemsetAsync(w.counter, 0, 2 * sizeof(unsigned long long), st));
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
  unsigned long long cnt2[2] = {0, 0};
  int64_t slots = 0;
  SB_CUDA(cudaMemcpyAsync(cnt2, w.counter, sizeof(cnt2), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&slots, w.off + f, sizeof(slots), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const unsigned long long c = cnt2[0];
  S.frontier_maxdeg = (int64_t)cnt2[1];
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
  {
    const char *ns = getenv("SB200_RCM_NO_SPEC");
    a.no_spec = ns && ns[0] == '1';
  }
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
    w.counter = ws.alloc<unsigned long long>(2);
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
                               (int)sizeof(ClSmem<I>)));
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  cfg.blockDim = dim3(kNwBlock, 1, 1);
  cfg.dynamicSmemBytes = sizeof(ClSmem<I>);
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int cluster = kClMax;
  if (const char *cs = getenv("SB200_RCM_CLUSTER")) {  // tuning aid: 1, 2, 4, 8 or 16
    const int want = atoi(cs);
    if (want >= 1 && want <= kClMax && (want & (want - 1)) == 0) cluster = want;
  }
  for (; cluster >= 1; cluster >>= 1) {  // largest cluster the device can co-schedule
    cfg.gridDim = dim3(cluster, 1, 1);
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    int nclusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (e == cudaSuccess && nclusters >= 1) break;
    cudaGetLastError();
  }
  SB_REQUIRE(cluster >= 1, SB200_ERR_CUDA, "cannot launch the RCM cluster kernel");
  const int64_t narrow_cap = (int64_t)cluster * kFl;
  for (;;) {
    SB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    launch_counter()++;
    SB_CUDA(cudaMemcpyAsync(&S, a.state, sizeof(S), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (S.status == ST_DONE) break;
    if (S.status == ST_NEED_WIDE) {
      prepare_wide();
      const bool cm = S.phase == PH_CM_LEVEL;
      {
        // the cluster kernel leaves the (monotone) claim key in mark[] of every vertex it
        // reached; the wide kernels use per-level keys and expect 0 = visited
        const int64_t from = cm ? S.qst : 0, cnt = S.lvl_end - from;
        if (cnt > 0)
          SB_LAUNCH((rcm_reset_kernel<I>), (unsigned)ceil_div(cnt, 256), 256, 0, st,
                    (const I *)(cm ? a.Q : a.Qp) + from, cnt, a.mark, 0u);
      }
      // keep going wide while the frontier is far beyond the narrow capacity
      do {
        rcm_wide_level<I, N>(ws, a, w, S, cm);
      } while (S.lvl_end - S.lvl_begin > (force_wide ? 0 : narrow_cap));
    } else if (S.status == ST_NEED_RESET) {
      const int64_t from = S.reset_cm ? S.qst : 0, cnt = S.lvl_end - from;
      SB_LAUNCH((rcm_reset_kernel<I>), (unsigned)ceil_div(cnt, 256), 256, 0, st,
                (const I *)(S.reset_cm ? a.Q : a.Qp) + from, cnt, a.mark, kUnvisited);
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

// {speculative CM traversals confirmed, continued from a unique new root, replayed literally}
int sb200_rcm_last_speculation(int64_t *h_out3) {
  if (!h_out3) return SB200_ERR_BAD_ARG;
  h_out3[0] = g_last_rcm_stats.stat_spec_ok;
  h_out3[1] = g_last_rcm_stats.stat_spec_chain;
  h_out3[2] = g_last_rcm_stats.stat_spec_fail;
  return SB200_OK;
}

// Cycle counters of the cluster kernel's level phases (CTA 0): load, claim, check, finalize,
// write -- a profiling aid, not part of the reference-facing surface.
int sb200_rcm_last_cycles(int64_t *h_out8) {
  if (!h_out8) return SB200_ERR_BAD_ARG;
  for (int i = 0; i < 8; i++) h_out8[i] = g_last_rcm_stats.cyc[i];
  return SB200_OK;
}

}  // extern "C"
