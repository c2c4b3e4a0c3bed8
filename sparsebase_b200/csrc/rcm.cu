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
// mark[v] (64 bits) is the claim key of v's first discoverer: (level << 32) | position.  Keys
// grow from level to level, so a vertex reached earlier always holds a SMALLER mark than any
// later claim: no separate "visited" store exists, and both regimes share the array.
//
// Two execution regimes share all global state (mark[], the queues, RcmState):
//   NARROW  one persistent thread-block cluster (1, 2, 4, 8 or 16 CTAs, chosen from the frontier
//           width) walks levels without returning to the host while the frontier fits shared
//           memory (high-diameter graphs: grids, bands -- thousands to millions of levels).
//   WIDE    one level at a time driven by the host with grid-wide kernels (power-law / random
//           graphs: a handful of levels, millions of vertices): exclusive scan of frontier
//           degrees, claim sweep, collect sweep, radix sort on the level's order key.
// The persistent kernel is a resumable state machine: when a level does not fit it stores its
// state and returns NEED_WIDE (or NEED_RESET / NEED_INVERT for bulk array work, NEED_RESIZE to be
// relaunched with another cluster size); the host runs that step and relaunches it.
#include <cooperative_groups.h>

#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace sb200 {
namespace cg = cooperative_groups;

constexpr unsigned long long kUnvisited = ~0ull;
constexpr int kNwBlock = 512;
constexpr int kNwWarps = kNwBlock / 32;
constexpr int kRounds = 4;                      // 32-slot rounds per warp and level, at most
constexpr int kPadCap = kNwBlock * kRounds;     // padded expansion slots per CTA and level
constexpr int64_t kBulkThreshold = 1 << 15;     // resets / inversions larger than this go wide
constexpr int kClMax = 16;

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
enum RcmStatus {
  ST_RUNNING = 0,
  ST_DONE,
  ST_NEED_WIDE,
  ST_NEED_RESET,
  ST_NEED_INVERT,
  ST_NEED_RESIZE
};

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
  int64_t frontier_maxdeg;  // max degree over the current frontier (sizes the narrow regime)
  int64_t resize_to;   // ST_NEED_RESIZE: cluster size the kernel asks for
  int64_t inv_qst, inv_end;  // pending bulk inversion
  int64_t reset_cm;    // pending bulk reset walks Q[qst, lvl_end) instead of Qp[0, lvl_end)
  int64_t stat_spec_ok, stat_spec_fail, stat_spec_chain;
  int64_t stat_levels_narrow, stat_levels_wide, stat_bfs, stat_components;
  int64_t stat_reloads, stat_resizes;
  int64_t cyc[8];  // narrow-level phase cycle counters (CTA 0): claim, barrier, recheck,
                   // compaction, sibling sort + state
};

template <typename I, typename N>
struct RcmArgs {
  int64_t n;
  const N *xadj;
  const I *adj;
  unsigned long long *mark;
  I *Q;    // CM order (reference: Q), concatenated components, not yet reversed
  I *Qp;   // peripheral BFS queue (reference: Qp)
  I *inv;  // result: inv[Qp2[i]] = i
  RcmState *state;
  int force_wide;  // testing: never take the narrow path
  int no_spec;     // testing: never run a peripheral BFS speculatively as the CM BFS
  // cluster sizing (0 = fixed size): grow / shrink when the running mean of the padded
  // expansion slots per CTA and level leaves [shrink_below, grow_above]; aim at `target`
  int max_cluster;
  int grow_above, shrink_below, target;
  int profile;  // accumulate the per-phase cycle counters (SB200_RCM_PROFILE=1)
};

// ------------------------------------------------------------------------------------
// Warp-cooperative expansion of a group of (up to) 32 frontier vertices (WIDE regime).  Lane l
// holds the adjacency start xs and degree d of vertex g+l; the warp walks the concatenated
// adjacency lists 32 slots at a time.  f(slot_in_group, owner_lane, adjacency_position, valid)
// is called once per round by every lane (valid == false for the padding of the last round).
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
// NARROW regime: one thread-block CLUSTER walks the levels, ONE cluster barrier per level.
//
// Every CTA keeps ITS share of the frontier in shared memory from one level to the next: the
// vertices it discovers (already ordered, with their adjacency extents) are its share of the
// next level, and the shares of the CTAs, taken in rank order, are the frontier in queue order.
// A level is
//     claims     every thread owns up to kRounds expansion slots of the share (slot = vertex i
//                of the share x adjacency index j, padded to the share's largest degree D so
//                that i = slot / D needs no search), loads the neighbour and claims it with
//                atomicMin(mark, key).  key = (level << 32) | (CTA rank << 16) | position is
//                order-isomorphic to the reference's queue position and needs nothing from the
//                other CTAs.  A thread REMEMBERS its claims in registers.
//     barrier    all claims of the level have landed.  The same barrier carries the exchange of
//                the share sizes (one remote shared-memory store per CTA pair before it).
//     recheck    a claim survived iff the mark still equals its key; the same round trip fetches
//                the winner's adjacency extent.  Winners are compacted in slot order (ballots +
//                one cross-warp prefix) and, for Cuthill-McKee, ranked inside their sibling
//                group by (degree, id): that IS the CTA's share of the next level.
// The share is written to the global queue (the result) one level late, when the share sizes of
// the lower-ranked CTAs are known; nobody waits for it.
// Shares drift apart over time (a BFS starts with everything in CTA 0); when the largest share
// exceeds twice the even share, or a share no longer fits kPadCap slots, the CTAs re-split the
// frontier evenly from the global queue (cl_reload).  A share that does not fit announces it in
// the exchange and makes no claims; the other CTAs take their claims of that level back
// (cl_level abort path), so the level can be repeated after the re-split or by the WIDE regime.
// All CTAs run the same state machine on replicated state, so control flow is uniform.
// ------------------------------------------------------------------------------------
template <typename I>
struct ClSmem {
  // this CTA's share of the current frontier: vertex, adjacency start, degree
  I Fv[kPadCap];
  int64_t Fx[kPadCap];
  unsigned Fd[kPadCap];
  // Cuthill-McKee: the winners in slot order, before the sibling sort (+ parent's share index)
  I Sv[kPadCap];
  int64_t Sx[kPadCap];
  unsigned Sd[kPadCap];
  unsigned Sp[kPadCap];
  unsigned long long lx[2][kClMax];      // per-level exchange: share size | D << 16 | flags
  __align__(8) unsigned long long lxbar[2];  // mbarriers: all records of a level have arrived
  unsigned long long xch[2][kClMax][4];  // generic exchange (component search)
  unsigned long long mine[4];
  unsigned wwin[kNwWarps];
  unsigned wmax[kNwWarps];
  unsigned wtot[kNwWarps + 2];
  unsigned scratch[34];
  unsigned long long red[kNwWarps];
  int fl;           // share size
  unsigned D;       // largest degree in the share (>= 1): padded slots per vertex
  unsigned invD;    // ceil(2^32 / D), for slot / D (unused when D == 1)
  unsigned rounds;  // 32-slot rounds per warp this level; > kRounds: the share does not fit
  // replicated (identical in every CTA):
  int need_reload;    // re-split the frontier from the global queue before the next level
  int share_written;  // the current frontier is in the global queue and S.lvl_end is known
  int just_reloaded;  // the shares are an even split: if they still do not fit, go wide
  int want_resize;    // ask the host for this cluster size before the next level
  int since_resize;
  unsigned long long work_acc;  // 16 x running mean of the padded slots per level
  RcmState S;
};

constexpr unsigned long long kLxNoFit = 1ull << 48;

__device__ __forceinline__ void cl_sync(cg::cluster_group &cluster, unsigned C) {
  if (C == 1)
    __syncthreads();
  else
    cluster.sync();
}

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
  cl_sync(cluster, C);
  par ^= 1;
}

// thread 0: the share now has fl vertices of degree <= dmax
template <typename I>
__device__ __forceinline__ void cl_set_shape(ClSmem<I> &s, int fl, unsigned dmax) {
  const unsigned D = dmax > 0u ? dmax : 1u;
  s.fl = fl;
  s.D = D;
  s.invD = D > 1u ? 0xffffffffu / D + 1u : 0u;
  const unsigned long long total = (unsigned long long)fl * D;
  s.rounds = total > (unsigned long long)kPadCap ? (unsigned)kRounds + 1u
                                                 : (unsigned)((total + kNwBlock - 1) / kNwBlock);
}

struct LxSum {
  unsigned cbase, ctotal, cmax, dmax;
  bool nofit;
};

__device__ __forceinline__ unsigned cl_smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}

// Per-level exchange AND synchronisation of the cluster in one step.  Every CTA sends its record
// (share size | D << 16 | no-fit flag) into slot [rank] of every CTA's lx[par] with st.async,
// which completes bytes on the RECEIVER's mbarrier; a CTA passes when the records of all C CTAs
// have landed in its own shared memory.  A CTA sends after its claims of the level have been
// performed (their results are consumed before the block barrier below), so passing also means
// "all claims of the level have landed" -- without the GPU-scope fence and L1 invalidation that
// barrier.cluster costs (MEMBAR.ALL.GPU + CCTL.IVALL, 1600 cycles per level measured).
// lxstate: bit 0 = buffer / barrier to use, bits 1-2 = phase parity of the two barriers.
template <typename I, typename Between>
__device__ __forceinline__ LxSum cl_lx_exchange(ClSmem<I> &s, unsigned C, unsigned rank,
                                                unsigned fl, unsigned D, bool fits,
                                                unsigned &lxstate, Between &&between) {
  const unsigned d = D > 0xffffffu ? 0xffffffu : D;
  // (share sizes never exceed kPadCap < 2^16)
  const unsigned long long rec =
      (unsigned long long)fl | ((unsigned long long)d << 16) | (fits ? 0ull : kLxNoFit);
  LxSum r = {0u, 0u, 0u, 0u, false};
  if (C == 1u) {  // nobody to talk to
    __syncthreads();
    between();
    r.ctotal = r.cmax = fl;
    r.dmax = fl > 0u ? d : 0u;
    r.nofit = !fits;
    return r;
  }
  const unsigned par = lxstate & 1u;
  __syncthreads();  // every warp's claims are done
  if (threadIdx.x < C) {
    unsigned raddr, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                 : "=r"(raddr)
                 : "r"(cl_smem_u32(&s.lx[par][rank])), "r"(threadIdx.x));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                 : "=r"(rbar)
                 : "r"(cl_smem_u32(&s.lxbar[par])), "r"(threadIdx.x));
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr),
        "l"(rec), "r"(rbar)
        : "memory");
  } else if (threadIdx.x == 32) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     cl_smem_u32(&s.lxbar[par])),
                 "r"(C * 8u)
                 : "memory");
  }
  {
    const unsigned phase = (lxstate >> (1u + par)) & 1u;
    unsigned ok;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(cl_smem_u32(&s.lxbar[par])), "r"(phase)
          : "memory");
    } while (!ok);
  }
  lxstate ^= 1u | (2u << par);
  between();
  // lane k holds CTA k's record
  const unsigned lane = lane_id();
  unsigned cnt = 0, dd = 0, nf = 0;
  if (lane < C) {
    const unsigned long long x = s.lx[par][lane];
    cnt = (unsigned)(x & 0xffffull);
    dd = cnt > 0u ? (unsigned)((x >> 16) & 0xffffffffull) : 0u;
    nf = (x & kLxNoFit) != 0ull ? 1u : 0u;
  }
  r.cbase = __reduce_add_sync(0xffffffffu, lane < rank ? cnt : 0u);
  r.ctotal = __reduce_add_sync(0xffffffffu, cnt);
  r.cmax = __reduce_max_sync(0xffffffffu, cnt);
  r.dmax = __reduce_max_sync(0xffffffffu, dd);
  r.nofit = __any_sync(0xffffffffu, nf != 0u);
  return r;
}

// The shares hold a frontier that is not in the global queue yet: exchange the sizes, write it.
// Contains a cluster barrier; the queue writes are visible to the other CTAs after the next one.
template <typename I>
__device__ void cl_flush_share(cg::cluster_group &cluster, ClSmem<I> &s, I *queue,
                               unsigned &lxstate) {
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const LxSum x = cl_lx_exchange(s, C, rank, (unsigned)s.fl, s.D, true, lxstate, [] {});
  for (int k = threadIdx.x; k < s.fl; k += kNwBlock) queue[s.S.lvl_begin + x.cbase + k] = s.Fv[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    s.S.lvl_end = s.S.lvl_begin + x.ctotal;
    s.S.frontier_maxdeg = x.dmax;
    s.share_written = 1;
  }
  __syncthreads();
}

// Re-split the frontier queue[lvl_begin, lvl_end) evenly over the CTAs (also the way the kernel
// resumes after a host-driven step).  Contains cluster barriers.
template <typename I, typename N>
__device__ void cl_reload(cg::cluster_group &cluster, const RcmArgs<I, N> &a, ClSmem<I> &s,
                          I *queue, unsigned &lxstate) {
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  if (!s.share_written) cl_flush_share(cluster, s, queue, lxstate);
  cl_sync(cluster, C);  // every CTA's queue writes are visible
  const int64_t f = s.S.lvl_end - s.S.lvl_begin;
  const int64_t fb = f * rank / C;
  int64_t fl = f * (rank + 1) / C - fb;
  unsigned dmax = 0;
  if (fl > kPadCap) {
    fl = 0;
    dmax = 0xffffffffu;  // does not fit whatever the degrees are
  } else {
    for (int i = threadIdx.x; i < (int)fl; i += kNwBlock) {
      const I v = __ldcg(queue + s.S.lvl_begin + fb + i);
      const int64_t xs = (int64_t)a.xadj[v];
      const unsigned d = (unsigned)((int64_t)a.xadj[v + 1] - xs);
      s.Fv[i] = v;
      s.Fx[i] = xs;
      s.Fd[i] = d;
      dmax = d > dmax ? d : dmax;
    }
  }
  dmax = __reduce_max_sync(0xffffffffu, dmax);
  if (lane == 0) s.wmax[wid] = dmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned m = 0;
    for (int w = 0; w < kNwWarps; w++) m = s.wmax[w] > m ? s.wmax[w] : m;
    cl_set_shape(s, (int)fl, m);
    if (m == 0xffffffffu) s.rounds = (unsigned)kRounds + 1u;
    s.need_reload = 0;
    s.just_reloaded = 1;
    s.S.stat_reloads++;
  }
  __syncthreads();
}

// BFS levels across the cluster, until something other than "next level" has to happen.  The
// level state (share shape, queue positions, depth, sizing statistics) lives in registers between
// levels -- every thread computes the same replicated values from the exchanged records, so no
// thread has a serial section -- and is stored back on the way out.  Returns
//    1  stopped after a level because the shares want a re-split or the cluster a new size
//    0  the frontier was empty: the BFS is complete, the queue is complete and visible
//   -1  some share did not fit: claims taken back, frontier in the queue; re-split and repeat
//   -2  an even split does not fit either: the level belongs to the WIDE regime
template <typename I, typename N, bool CM>
__device__ __forceinline__ int cl_levels(cg::cluster_group &cluster, const RcmArgs<I, N> &a,
                                         ClSmem<I> &s, I *queue, unsigned &lxstate) {
  const unsigned C = cluster.num_blocks(), rank = cluster.block_rank();
  const unsigned lane = lane_id(), wid = threadIdx.x >> 5;
  int fl = s.fl;
  unsigned D = s.D, invD = s.invD, R = s.rounds;
  int64_t lvl_begin = s.S.lvl_begin, prev_begin = s.S.prev_begin, lvl_end = s.S.lvl_end;
  int64_t depth = s.S.depth, frontier_maxdeg = s.S.frontier_maxdeg;
  bool written = s.share_written != 0, reloaded = s.just_reloaded != 0;
  unsigned long long acc = s.work_acc;
  int since = s.since_resize, nlevels = 0, need_reload = 0, want_resize = 0, ret;
  long long t0 = a.profile ? clock64() : 0, t1;
#define SB_TICK(slot)                                   \
  do {                                                  \
    if (a.profile) {                                    \
      t1 = clock64();                                   \
      if (threadIdx.x == 0) s.S.cyc[slot] += t1 - t0;   \
      t0 = t1;                                          \
    }                                                   \
  } while (0)

  for (;;) {
    const bool fits = R <= (unsigned)kRounds;
    const unsigned total = fits ? (unsigned)fl * D : 0u;
    // level << 32 | rank << 16 | position inside the share: grows with the queue position
    const unsigned long long klevel =
        ((unsigned long long)(depth + 1) << 32) | ((unsigned long long)rank << 16);

    // ---- claims: slot p = i * D + j of the share; warp w owns slots [w * 32 R, (w + 1) * 32 R)
    I v[kRounds];
    unsigned valid = 0, prov = 0;
    const unsigned wbase = wid * 32u * R + lane;
    // claim key of round r: the position part is the parent's index in the share (CM) or the
    // slot (peripheral); recomputed where needed instead of being kept in registers
    auto key_of = [&](int r) -> unsigned long long {
      const unsigned p = wbase + (unsigned)r * 32u;
      const unsigned i = D == 1u ? p : __umulhi(p, invD);
      return klevel | (unsigned long long)(CM ? i : p);
    };
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
      if ((unsigned)r < R) {
        const unsigned p = wbase + (unsigned)r * 32u;
        if (p < total) {
          const unsigned i = D == 1u ? p : __umulhi(p, invD);
          const unsigned j = p - i * D;
          if (j < s.Fd[i]) {
            valid |= 1u << r;
            v[r] = a.adj[s.Fx[i] + (int64_t)j];
          }
        }
      }
    }
    unsigned long long old[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; r++)
      if ((valid >> r) & 1u) old[r] = atomicMin(&a.mark[v[r]], key_of(r));
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
      if (((valid >> r) & 1u) && old[r] > key_of(r)) {
        prov |= 1u << r;  // the smallest claim so far: a candidate
        // its adjacency extent is read right after the exchange: pull the line towards L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.xadj + v[r]));
      }
    }
    SB_TICK(0);

    // ---- exchange (= all claims have landed), then the recheck loads first of all ----
    unsigned long long m[kRounds];
    N xs[kRounds], xe[kRounds];
    const LxSum x = cl_lx_exchange(s, C, rank, (unsigned)fl, D, fits, lxstate, [&] {
#pragma unroll
      for (int r = 0; r < kRounds; r++) {
        if ((prov >> r) & 1u) {
          m[r] = __ldcg(&a.mark[v[r]]);
          xs[r] = a.xadj[v[r]];
          xe[r] = a.xadj[v[r] + 1];
        }
      }
    });
    SB_TICK(1);
    if (x.nofit) {  // uniform: take the claims back, publish the frontier, let the caller decide
#pragma unroll
      for (int r = 0; r < kRounds; r++)
        if ((prov >> r) & 1u) atomicExch(&a.mark[v[r]], kUnvisited);
      if (!written) {
        for (int k = threadIdx.x; k < fl; k += kNwBlock) queue[lvl_begin + x.cbase + k] = s.Fv[k];
        lvl_end = lvl_begin + x.ctotal;
        written = true;
      }
      frontier_maxdeg = x.dmax;
      cl_sync(cluster, C);
      ret = reloaded ? -2 : -1;
      break;
    }
    if (!written)
      for (int k = threadIdx.x; k < fl; k += kNwBlock) queue[lvl_begin + x.cbase + k] = s.Fv[k];
    if (x.ctotal == 0u) {
      lvl_end = lvl_begin;
      written = true;
      cl_sync(cluster, C);  // every CTA's queue writes of the last levels are visible
      ret = 0;
      break;
    }

    // ---- winners, in slot order inside the warp ----
    unsigned win = 0, run = 0, wmaxd = 0;
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
      if ((unsigned)r < R) {
        const bool w = ((prov >> r) & 1u) && m[r] == key_of(r);
        if (w) {
          win |= 1u << r;
          const unsigned dg = (unsigned)(xe[r] - xs[r]);
          wmaxd = dg > wmaxd ? dg : wmaxd;
        }
        run += __popc(__ballot_sync(0xffffffffu, w));
      }
    }
    wmaxd = __reduce_max_sync(0xffffffffu, wmaxd);
    if (lane == 0) {
      s.wwin[wid] = run;
      s.wmax[wid] = wmaxd;
    }
    SB_TICK(2);
    __syncthreads();  // the share has been read by everybody (claims, queue write)
    // lane w holds warp w's totals
    const unsigned cw = lane < (unsigned)kNwWarps ? s.wwin[lane] : 0u;
    const unsigned mw = lane < (unsigned)kNwWarps ? s.wmax[lane] : 0u;
    unsigned wb = __reduce_add_sync(0xffffffffu, lane < wid ? cw : 0u);
    const unsigned c = __reduce_add_sync(0xffffffffu, cw);
    const unsigned dmax = __reduce_max_sync(0xffffffffu, mw);
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
      if ((unsigned)r < R) {
        const bool w = (win >> r) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, w);
        if (w) {
          const unsigned at = wb + __popc(bal & lanemask_lt());
          const unsigned dg = (unsigned)(xe[r] - xs[r]);
          if (CM) {
            s.Sv[at] = v[r];
            s.Sp[at] = (unsigned)key_of(r) & 0xffffu;
            s.Sx[at] = (int64_t)xs[r];
            s.Sd[at] = dg;
          } else {
            s.Fv[at] = v[r];
            s.Fx[at] = (int64_t)xs[r];
            s.Fd[at] = dg;
          }
          // the winner is expanded in the next level: pull its adjacency towards L2 meanwhile
          // (the lists are read once per traversal: the expansion would otherwise wait for HBM)
          if (dg > 0u) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + xs[r]));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + xe[r] - 1));
          }
        }
        wb += __popc(bal);
      }
    }
    SB_TICK(3);
    if (CM) {
      __syncthreads();
      // (parent, degree, id) order: the winners of one parent are adjacent in slot order and
      // never leave the CTA that owns the parent; rank inside the sibling group by enumeration
      if (D <= 8u) {  // small sibling groups: one thread per winner
        for (int k = threadIdx.x; k < (int)c; k += kNwBlock) {
          const unsigned par_i = s.Sp[k], dgk = s.Sd[k];
          const I vk = s.Sv[k];
          const int64_t xk = s.Sx[k];
          int left = 0, before = 0;
#pragma unroll
          for (int u = 1; u < 8; u++) {  // a group has at most D <= 8 members
            const int ql = k - u, qr = k + u;
            if (ql >= 0 && s.Sp[ql] == par_i) {
              left++;
              before += (s.Sd[ql] < dgk || (s.Sd[ql] == dgk && s.Sv[ql] < vk)) ? 1 : 0;
            }
            if (qr < (int)c && s.Sp[qr] == par_i)
              before += (s.Sd[qr] < dgk || (s.Sd[qr] == dgk && s.Sv[qr] < vk)) ? 1 : 0;
          }
          const int dst = k - left + before;
          s.Fv[dst] = vk;
          s.Fx[dst] = xk;
          s.Fd[dst] = dgk;
        }
      } else {  // wide sibling groups (bands): one warp per winner, 32 siblings per step
        for (int k = wid; k < (int)c; k += kNwWarps) {
          const unsigned par_i = s.Sp[k], dgk = s.Sd[k];
          const I vk = s.Sv[k];
          int left = 0, before = 0;
          for (int base = k - 1; base >= 0; base -= 32) {
            const int q = base - (int)lane;
            const bool mt = q >= 0 && s.Sp[q] == par_i;
            const unsigned bal = __ballot_sync(0xffffffffu, mt);
            const int cnt = bal == 0xffffffffu ? 32 : __ffs(~bal) - 1;  // matches are a prefix
            const bool less = (int)lane < cnt &&
                              (s.Sd[q] < dgk || (s.Sd[q] == dgk && s.Sv[q] < vk));
            before += __popc(__ballot_sync(0xffffffffu, less));
            left += cnt;
            if (cnt < 32) break;
          }
          for (int base = k + 1; base < (int)c; base += 32) {
            const int q = base + (int)lane;
            const bool mt = q < (int)c && s.Sp[q] == par_i;
            const unsigned bal = __ballot_sync(0xffffffffu, mt);
            const int cnt = bal == 0xffffffffu ? 32 : __ffs(~bal) - 1;
            const bool less = (int)lane < cnt &&
                              (s.Sd[q] < dgk || (s.Sd[q] == dgk && s.Sv[q] < vk));
            before += __popc(__ballot_sync(0xffffffffu, less));
            if (cnt < 32) break;
          }
          if (lane == 0) {
            const int dst = k - left + before;
            s.Fv[dst] = vk;
            s.Fx[dst] = s.Sx[k];
            s.Fd[dst] = dgk;
          }
        }
      }
    }
    // ---- the next level's state: the same arithmetic in every thread of every CTA ----
    {
      const unsigned nd = dmax > 0u ? dmax : 1u;
      if (nd != D) {
        D = nd;
        invD = D > 1u ? 0xffffffffu / D + 1u : 0u;
      }
      fl = (int)c;
      const unsigned tot = c * D;  // c <= kPadCap, D < 2^31 / kPadCap whenever it matters
      R = (D > (unsigned)kPadCap || tot > (unsigned)kPadCap) ? (unsigned)kRounds + 1u
                                                             : (tot + kNwBlock - 1) / kNwBlock;
      prev_begin = lvl_begin;
      lvl_begin += x.ctotal;
      depth++;
      nlevels++;
      written = false;
      reloaded = false;
      // the shares of the level just expanded: re-split when they have drifted apart (or when
      // one of them nears the capacity an even split would stay well below)
      const unsigned even = (x.ctotal + C - 1u) / C, dd = x.dmax > 0u ? x.dmax : 1u;
      if (C > 1u && (x.cmax > 2u * even + 32u ||
                     ((unsigned long long)x.cmax * dd > (unsigned long long)kPadCap * 3 / 4 &&
                      (unsigned long long)even * dd <= (unsigned long long)kPadCap / 2)))
        need_reload = 1;
      if (a.max_cluster > 0) {  // cluster sizing on the running mean of the padded slots
        const unsigned long long work = (unsigned long long)x.ctotal * dd;
        acc = acc == ~0ull ? work * 16ull : acc - acc / 16ull + work;
        if (++since >= 32) {
          const unsigned long long mean = acc / 16ull;
          const bool grow = (int)C < a.max_cluster && mean > (unsigned long long)a.grow_above * C;
          const bool shrink = C > 1u && mean < (unsigned long long)a.shrink_below * C;
          if (grow || shrink) {
            int want = 1;
            while (want < a.max_cluster && (unsigned long long)want * a.target < mean) want <<= 1;
            if (want != (int)C) want_resize = want;
          }
        }
      }
    }
    if (need_reload || want_resize) {
      ret = 1;
      break;
    }
    __syncthreads();  // the next share is complete
    SB_TICK(4);
  }
#undef SB_TICK
  __syncthreads();
  if (threadIdx.x == 0) {
    s.fl = fl;
    s.D = D;
    s.invD = invD;
    s.rounds = R;
    s.S.lvl_begin = lvl_begin;
    s.S.prev_begin = prev_begin;
    s.S.lvl_end = lvl_end;
    s.S.depth = depth;
    s.S.frontier_maxdeg = frontier_maxdeg;
    s.S.stat_levels_narrow += nlevels;
    s.share_written = written ? 1 : 0;
    s.just_reloaded = reloaded ? 1 : 0;
    if (need_reload) s.need_reload = 1;
    if (want_resize) s.want_resize = want_resize;
    s.work_acc = acc;
    s.since_resize = since;
  }
  __syncthreads();
  return ret;
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
  unsigned lxstate = 0;  // see cl_lx_exchange
  if (threadIdx.x == 0) {
    s.S = *a.state;
    s.S.status = ST_RUNNING;
    cl_set_shape(s, 0, 1);
    s.need_reload = 1;    // nothing is resident yet: a level in progress re-splits the queue
    s.share_written = 1;  // ... which holds the whole frontier (S.lvl_begin .. S.lvl_end)
    s.just_reloaded = 0;
    s.want_resize = 0;
    s.since_resize = 0;
    s.work_acc = ~0ull;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cl_smem_u32(&s.lxbar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cl_smem_u32(&s.lxbar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cl_sync(cluster, C);  // the barriers exist before any CTA sends to them

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
            atomicExch(&a.mark[i], 0ull);
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
      const int64_t rxs = (int64_t)a.xadj[r];
      const unsigned rd = (unsigned)((int64_t)a.xadj[r + 1] - rxs);
      if (rank == 0 && threadIdx.x == 0) {
        (cm ? a.Q : a.Qp)[at] = r;
        atomicExch(&a.mark[r], 0ull);
      }
      if (threadIdx.x == 0) {
        if (cm) {
          s.S.qst = s.S.qwp;
        } else {
          s.S.rlevel = s.S.qlevel;
        }
        s.S.lvl_begin = at;
        s.S.lvl_end = at + 1;
        s.S.prev_begin = at;
        s.S.depth = 0;
        s.S.frontier_maxdeg = rd;
        s.S.phase = cm ? PH_CM_LEVEL : PH_PBFS_LEVEL;
        s.S.stat_bfs++;
        // the root is the whole frontier and lives in CTA 0's share
        if (rank == 0) {
          s.Fv[0] = r;
          s.Fx[0] = rxs;
          s.Fd[0] = rd;
          cl_set_shape(s, 1, rd);
        } else {
          cl_set_shape(s, 0, 1);
        }
        s.share_written = 1;
        s.need_reload = 0;
        s.just_reloaded = 0;
      }
      __syncthreads();
      continue;
    }

    if (phase == PH_PBFS_LEVEL || phase == PH_CM_LEVEL) {
      I *queue = phase == PH_PBFS_LEVEL ? a.Qp : a.Q;
      if (a.force_wide) {  // (the frontier is always in the queue here)
        if (s.S.lvl_end == s.S.lvl_begin) {  // the wide regime emptied it: the BFS is complete
          __syncthreads();
          if (threadIdx.x == 0) s.S.phase = phase == PH_PBFS_LEVEL ? PH_PBFS_END : PH_CM_END;
          __syncthreads();
          continue;
        }
        if (threadIdx.x == 0) s.S.status = ST_NEED_WIDE;
        break;
      }
      if (s.want_resize) {
        if (!s.share_written) cl_flush_share(cluster, s, queue, lxstate);
        cl_sync(cluster, C);
        if (threadIdx.x == 0) {
          s.S.status = ST_NEED_RESIZE;
          s.S.resize_to = s.want_resize;
          s.S.stat_resizes++;
        }
        break;
      }
      if (s.need_reload) cl_reload<I, N>(cluster, a, s, queue, lxstate);
      const int r = phase == PH_PBFS_LEVEL ? cl_levels<I, N, false>(cluster, a, s, queue, lxstate)
                                           : cl_levels<I, N, true>(cluster, a, s, queue, lxstate);
      if (r == 1) continue;  // (re-split or resize pending)
      if (r == 0) {
        if (threadIdx.x == 0) s.S.phase = phase == PH_PBFS_LEVEL ? PH_PBFS_END : PH_CM_END;
        __syncthreads();
        continue;
      }
      if (r == -1) {
        if (threadIdx.x == 0) s.need_reload = 1;
        __syncthreads();
        continue;
      }
      __syncthreads();
      if (threadIdx.x == 0) s.S.status = ST_NEED_WIDE;
      break;
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
        new_root = (int64_t)__ldcg(a.Qp + s.S.prev_begin + (int64_t)(best & 0xffffffffull));
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
        atomicExch(&a.mark[__ldcg(a.Qp + k)], kUnvisited);
      cl_sync(cluster, C);
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
          const int64_t new_root =
              (int64_t)__ldcg(a.Q + s.S.prev_begin + (int64_t)(best & 0xffffffffull));
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
            atomicExch(&a.mark[__ldcg(a.Q + k)], kUnvisited);
          cl_sync(cluster, C);
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
        a.inv[__ldcg(a.Q + k)] = (I)(qst + (end - 1 - k));
      cl_sync(cluster, C);
      continue;
    }
  }
  __syncthreads();
  if (rank == 0 && threadIdx.x == 0) *a.state = s.S;
  cl_sync(cluster, C);  // no CTA leaves while a peer may still address its shared memory
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

// sweep 1: claims.  key = level << 32 | expansion slot (peripheral) or parent position (CM)
template <typename I, typename N, bool CM>
__global__ void __launch_bounds__(256)
    rcm_wide_claim_kernel(const I *__restrict__ frontier, int64_t f, const int64_t *__restrict__ off,
                          const N *__restrict__ xadj, const I *__restrict__ adj,
                          unsigned long long *__restrict__ mark, unsigned long long klevel) {
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
        const unsigned long long key =
            klevel | (CM ? (unsigned long long)(g + j) : (unsigned long long)(slot0 + sl));
        if (__ldcg(&mark[v]) > key) atomicMin(&mark[v], key);
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
                            const I *__restrict__ adj, const unsigned long long *__restrict__ mark,
                            unsigned long long klevel, uint64_t *__restrict__ ckey,
                            uint32_t *__restrict__ cval, unsigned long long *__restrict__ counter) {
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
      unsigned long long pk = 0;  // position part of the key
      if (valid) {
        v = adj[p];
        pk = CM ? (unsigned long long)(g + j) : (unsigned long long)(slot0 + sl);
        win = __ldcg(&mark[v]) == (klevel | pk);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, win);
      if (bal) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (win) {
          const unsigned long long k = base + __popc(bal & lanemask_lt());
          const unsigned dg = (unsigned)(xadj[v + 1] - xadj[v]);
          ckey[k] = CM ? ((uint64_t)pk << 32) | (uint64_t)dg : (uint64_t)pk;
          cval[k] = (uint32_t)v;
          atomicMax(counter + 1, (unsigned long long)dg);
        }
      }
    });
  }
}

template <typename I>
__global__ void rcm_wide_commit_kernel(const uint32_t *__restrict__ sorted, int64_t c,
                                       I *__restrict__ queue_out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < c) queue_out[k] = (I)sorted[k];
}

template <typename I>
__global__ void rcm_reset_kernel(const I *__restrict__ q, int64_t cnt,
                                 unsigned long long *__restrict__ mark, unsigned long long value) {
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
  const unsigned long long klevel = (unsigned long long)(S.depth + 1) << 32;
  exclusive_scan<int64_t>(ws, FrontierDegFn<I, N>{frontier, a.xadj}, w.off, f);
  SB_CUDA(cudaMemsetAsync(w.counter, 0, 2 * sizeof(unsigned long long), st));
  if (cm) {
    SB_LAUNCH((rcm_wide_claim_kernel<I, N, true>), grid, 256, 0, st, frontier, f,
              (const int64_t *)w.off, a.xadj, a.adj, a.mark, klevel);
    SB_LAUNCH((rcm_wide_collect_kernel<I, N, true>), grid, 256, 0, st, frontier, f,
              (const int64_t *)w.off, a.xadj, a.adj, (const unsigned long long *)a.mark, klevel,
              w.k0, w.v0, w.counter);
  } else {
    SB_LAUNCH((rcm_wide_claim_kernel<I, N, false>), grid, 256, 0, st, frontier, f,
              (const int64_t *)w.off, a.xadj, a.adj, a.mark, klevel);
    SB_LAUNCH((rcm_wide_collect_kernel<I, N, false>), grid, 256, 0, st, frontier, f,
              (const int64_t *)w.off, a.xadj, a.adj, (const unsigned long long *)a.mark, klevel,
              w.k0, w.v0, w.counter);
  }
  unsigned long long cnt2[2] = {0, 0};
  int64_t slots = 0;
  SB_CUDA(cudaMemcpyAsync(cnt2, w.counter, sizeof(cnt2), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&slots, w.off + f, sizeof(slots), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const unsigned long long c = cnt2[0];
  S.frontier_maxdeg = (int64_t)cnt2[1];
  SB_REQUIRE(slots < 0xfffffffell, SB200_ERR_BAD_ARG,
             "RCM level with %lld expansion slots exceeds the 32-bit claim position",
             (long long)slots);
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
              sorted_v, (int64_t)c, queue + S.lvl_end);
  }
  S.prev_begin = S.lvl_begin;
  S.lvl_begin = S.lvl_end;
  S.lvl_end += (int64_t)c;
  S.depth++;
  S.stat_levels_wide++;
}

static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
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
  a.mark = ws.alloc<unsigned long long>(n);
  a.Q = ws.alloc<I>(n);
  a.Qp = ws.alloc<I>(n);
  a.inv = out_inv;
  a.state = ws.alloc<RcmState>(1);
  a.force_wide = force_wide;
  a.no_spec = env_int("SB200_RCM_NO_SPEC", 0) == 1;
  a.profile = env_int("SB200_RCM_PROFILE", 0) == 1;
  SB_CUDA(cudaMemsetAsync(a.mark, 0xff, n * sizeof(unsigned long long), st));
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
  auto set_cluster = [&](int c) {
    cfg.gridDim = dim3(c, 1, 1);
    attr[0].val.clusterDim.x = c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
  };
  // largest cluster the device can co-schedule
  int max_cluster = kClMax;
  for (; max_cluster >= 1; max_cluster >>= 1) {
    set_cluster(max_cluster);
    int nclusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (e == cudaSuccess && nclusters >= 1) break;
    cudaGetLastError();
  }
  SB_REQUIRE(max_cluster >= 1, SB200_ERR_CUDA, "cannot launch the RCM cluster kernel");
  // SB200_RCM_CLUSTER=c pins the cluster size (tuning aid: 1, 2, 4, 8 or 16); otherwise the
  // kernel asks to be resized from the running mean of the frontier width
  int cluster = max_cluster;
  bool pinned = false;
  {
    const int want = env_int("SB200_RCM_CLUSTER", 0);
    if (want >= 1 && want <= max_cluster && (want & (want - 1)) == 0) {
      cluster = want;
      pinned = true;
    }
  }
  a.max_cluster = pinned ? 0 : max_cluster;
  a.grow_above = env_int("SB200_RCM_GROW", kPadCap / 2);
  a.shrink_below = env_int("SB200_RCM_SHRINK", kPadCap / 16);
  a.target = env_int("SB200_RCM_TARGET", kPadCap / 4);
  if (!pinned) cluster = 1;  // a BFS starts with one vertex
  // an even split of f vertices of degree <= d over c CTAs fits the narrow regime
  auto fits_narrow = [&](int64_t f, int64_t d, int c) {
    return (f / c + 1) * (d > 0 ? d : 1) <= (int64_t)kPadCap;
  };
  for (;;) {
    set_cluster(cluster);
    SB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    launch_counter()++;
    SB_CUDA(cudaMemcpyAsync(&S, a.state, sizeof(S), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (S.status == ST_DONE) break;
    if (S.status == ST_NEED_WIDE) {
      const bool cm = S.phase == PH_CM_LEVEL;
      const int64_t f = S.lvl_end - S.lvl_begin;
      // a larger cluster may hold the level that did not fit this one
      int bigger = cluster;
      while (!pinned && !force_wide && bigger < max_cluster &&
             !fits_narrow(f, S.frontier_maxdeg, bigger))
        bigger <<= 1;
      if (!force_wide && bigger != cluster && fits_narrow(f, S.frontier_maxdeg, bigger)) {
        cluster = bigger;
      } else {
        prepare_wide();
        // keep going wide while the frontier is beyond the narrow capacity
        const int cap_cluster = pinned ? cluster : max_cluster;
        do {
          rcm_wide_level<I, N>(ws, a, w, S, cm);
        } while (S.lvl_end > S.lvl_begin &&
                 (force_wide ||
                  !fits_narrow(S.lvl_end - S.lvl_begin, S.frontier_maxdeg, cap_cluster)));
        if (!pinned) cluster = cap_cluster;
      }
    } else if (S.status == ST_NEED_RESIZE) {
      int c = (int)S.resize_to;
      if (c < 1) c = 1;
      if (c > max_cluster) c = max_cluster;
      cluster = c;
    } else if (S.status == ST_NEED_RESET) {
      const int64_t from = S.reset_cm ? S.qst : 0, cnt = S.lvl_end - from;
      SB_LAUNCH((rcm_reset_kernel<I>), (unsigned)ceil_div(cnt, 256), 256, 0, st,
                (const I *)(S.reset_cm ? a.Q : a.Qp) + from, cnt, a.mark, kUnvisited);
      if (!pinned) cluster = 1;  // the next BFS starts from a single vertex
    } else if (S.status == ST_NEED_INVERT) {
      const int64_t cnt = S.lvl_end - S.qst;
      SB_LAUNCH((rcm_invert_kernel<I>), (unsigned)ceil_div(cnt, 256), 256, 0, st, (const I *)a.Q,
                S.qst, S.lvl_end, a.inv);
      if (!pinned) cluster = 1;
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
    const int force_wide = env_int("SB200_RCM_FORCE_WIDE", 0) == 1;
    dispatch_inv(id_type, nnz_type, SB200_VOID, false, [&](auto I_, auto N_, auto) {
      using I = decltype(I_);
      using N = decltype(N_);
      rcm_impl<I, N>(ws, n, nnz, (const N *)row_ptr, (const I *)col, (I *)out_inv, force_wide,
                     &g_last_rcm_stats);
    });
  });
}

// Diagnostics of the last sb200_rcm_reorder call on this thread:
// out[0..3] = levels done by the persistent cluster kernel, levels done wide, BFS count,
// components; out[4..5] (sb200_rcm_last_stats6) = share re-splits, cluster resizes.
int sb200_rcm_last_stats(int64_t *h_out4) {
  if (!h_out4) return SB200_ERR_BAD_ARG;
  h_out4[0] = g_last_rcm_stats.stat_levels_narrow;
  h_out4[1] = g_last_rcm_stats.stat_levels_wide;
  h_out4[2] = g_last_rcm_stats.stat_bfs;
  h_out4[3] = g_last_rcm_stats.stat_components;
  return SB200_OK;
}

int sb200_rcm_last_resplits(int64_t *h_out2) {
  if (!h_out2) return SB200_ERR_BAD_ARG;
  h_out2[0] = g_last_rcm_stats.stat_reloads;
  h_out2[1] = g_last_rcm_stats.stat_resizes;
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

// Cycle counters of the cluster kernel's level phases (CTA 0): claims, barrier, recheck,
// compaction, sibling sort + state update -- a profiling aid, not part of the reference-facing
// surface.
int sb200_rcm_last_cycles(int64_t *h_out8) {
  if (!h_out8) return SB200_ERR_BAD_ARG;
  for (int i = 0; i < 8; i++) h_out8[i] = g_last_rcm_stats.cyc[i];
  return SB200_OK;
}

}  // extern "C"
