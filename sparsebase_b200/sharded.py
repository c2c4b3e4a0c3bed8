"""Row-block sharded operators: one process per GPU, ``torch.distributed`` for the exchanges.

The reference has no distributed code; this module is the multi-GPU form
of the same operators.  A matrix is split into
``world`` contiguous row blocks with (nearly) equal nnz; rank r owns rows
``[bounds[r], bounds[r+1])``.  All arithmetic on nonzeros runs in the local CUDA kernels of
libsb200.so (``ops``, by default :mod:`sparsebase_b200.lib`); the collectives carry

    COO->CSR            all_gather of one int64 per rank (nnz totals -> row_ptr base offsets)
    Degrees / DegreeDistribution   nothing (global nnz comes with the shard)
    DegreeReorder       all_reduce(max) of the max degree, all_gather of the per-rank degree
                        histograms, all_gather of the permutation slices
    Permute2D           all_gather of the degrees (new-row balancing) + ONE personalised
                        all_to_all of whole rows (ids already renumbered and sorted by the sender)
    CSR->CSC            all_reduce of the column counts + ONE personalised all_to_all of whole
                        column pieces (a transpose IS an exchange)
    RCMReorder          not sharded: levels serialise (replicas only)

Results are bit-identical to the single-GPU operators (tests/test_sharded_cpu.py with gloo and
an oracle-backed stand-in for ``ops``; tests/test_sharded_gpu.py on 2 GPUs with NCCL).
``ops`` is an explicit parameter only so that the host logic can be exercised without a GPU;
the product path never runs without the CUDA library.
"""
import os
import time
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------- plumbing
class _Trace:
    """SB200_SHARD_TRACE=1: per-stage wall time (device synchronised) printed by rank 0."""
    on = os.environ.get("SB200_SHARD_TRACE", "0") == "1"

    def __init__(self, op):
        self.op, self.t, self.rows = op, None, []
        if self.on:
            torch.cuda.synchronize()
            self.t = time.perf_counter()

    def mark(self, stage):
        if self.on:
            torch.cuda.synchronize()
            now = time.perf_counter()
            self.rows.append((stage, (now - self.t) * 1e3))
            self.t = now

    def done(self):
        if self.on and _world()[0] == 0:
            print(f"[shard-trace] {self.op}: " +
                  "  ".join(f"{k}={v:.3f}ms" for k, v in self.rows), flush=True)


def _world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _all_gather_i64(value, device, group=None):
    """One int64 per rank -> tensor[world] (on the host)."""
    rank, world = _world(group)
    if world == 1:
        return torch.tensor([int(value)], dtype=torch.int64)
    mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.cpu()


def _all_gather_var(t, counts, group=None):
    """Concatenation over ranks of 1-D tensors with per-rank lengths `counts` (host list)."""
    rank, world = _world(group)
    if world == 1:
        return t
    mx = int(max(counts))
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    out = torch.empty(world * mx, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx: r * mx + int(counts[r])] for r in range(world)])


def _all_to_all_var(t, send_counts, recv_counts, group=None):
    """Personalised exchange of contiguous slices of a 1-D tensor."""
    rank, world = _world(group)
    if world == 1:
        return t
    out = torch.empty(int(sum(recv_counts)), dtype=t.dtype, device=t.device)
    dist.all_to_all_single(out, t.contiguous(), [int(x) for x in recv_counts],
                           [int(x) for x in send_counts], group=group)
    return out


def _exchange_counts(send_counts, device, group=None):
    """send_counts[q] = what I send to q  ->  recv_counts[s] = what s sends to me."""
    rank, world = _world(group)
    if world == 1:
        return list(send_counts)
    s = torch.tensor([int(x) for x in send_counts], dtype=torch.int64, device=device)
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s, group=group)
    return [int(x) for x in r.cpu()]


_ORDER_CACHE = {}


def _interleave_order(world, ncl, dtype, device):
    """new position c * world + s of segment (source s, column c); a function of the shape only,
    kept between calls (three elementwise kernels over world * ncl entries otherwise)."""
    key = (world, ncl, dtype, str(device))
    if key not in _ORDER_CACHE:
        if len(_ORDER_CACHE) > 4:
            _ORDER_CACHE.clear()
        k = torch.arange(world * ncl, dtype=torch.int64, device=device)
        _ORDER_CACHE[key] = ((k % ncl) * world + k // ncl).to(dtype)
    return _ORDER_CACHE[key]


def even_bounds(n, world):
    return [n * k // world for k in range(world + 1)]


# ------------------------------------------------------------------------------- data model
@dataclass
class ShardedCSR:
    """Rows [row_lo, row_hi) of an n x m CSR.  row_ptr is block-local (starts at 0); col holds
    global column ids; nnz_base is the number of nonzeros in the blocks before this one."""
    n: int
    m: int
    nnz: int            # global
    bounds: list        # world+1 row boundaries (host ints)
    row_ptr: torch.Tensor
    col: torch.Tensor
    vals: Optional[torch.Tensor]
    nnz_base: int

    def block(self, rank):
        return self.bounds[rank], self.bounds[rank + 1]

    def global_row_ptr(self, group=None):
        """The replicated global row_ptr (n+1 entries): all_gather of the shifted segments."""
        rank, world = _world(group)
        seg = self.row_ptr[:-1] + self.nnz_base
        counts = [self.bounds[r + 1] - self.bounds[r] for r in range(world)]
        full = _all_gather_var(seg, counts, group)
        tail = torch.tensor([self.nnz], dtype=self.row_ptr.dtype, device=self.row_ptr.device)
        return torch.cat([full, tail])


@dataclass
class ShardedCSC:
    """Columns [col_lo, col_hi) of an n x m CSC; col_ptr block-local, row holds global ids."""
    n: int
    m: int
    nnz: int
    bounds: list
    col_ptr: torch.Tensor
    row: torch.Tensor
    vals: Optional[torch.Tensor]
    nnz_base: int


def shard_csr(ops, n, m, row_ptr, col, vals, rank, world):
    """Cut a replicated/global CSR into this rank's nnz-balanced row block (test/bench setup)."""
    nnz = int(col.numel())
    bounds = ops.partition_rows(n, nnz, row_ptr, world) if world > 1 else [0, n]
    lo, hi = bounds[rank], bounds[rank + 1]
    base = int(row_ptr[lo])
    end = int(row_ptr[hi])
    return ShardedCSR(n, m, nnz, bounds, (row_ptr[lo:hi + 1] - base).contiguous(),
                      col[base:end].contiguous(),
                      None if vals is None else vals[base:end].contiguous(), base)


# ------------------------------------------------------------------------------- COO -> CSR
def coo_to_csr(ops, n, m, bounds, row, col, vals, group=None, nnz_dtype=torch.int32, copy=True):
    """`row/col/vals` are this rank's entries: all nonzeros of rows [bounds[rank], bounds[rank+1])
    in any order.  COO-constructor sort + histogram/scan/copy run locally; one 8-byte
    all_gather assembles the global offsets.  With ``copy=False`` the caller's arrays are sorted
    in place, as the reference's COO constructor does (format/coo.cc:110-157)."""
    rank, world = _world(group)
    lo, hi = bounds[rank], bounds[rank + 1]
    if copy:
        row, col = row.clone(), col.clone()
        vals = None if vals is None else vals.clone()
    ops.coo_sort_(n, m, row, col, vals)                      # format::COO constructor
    row_ptr, ocol, ovals = ops.coo_to_csr_block(lo, hi - lo, m, row, col, vals,
                                                nnz_dtype=nnz_dtype)
    totals = _all_gather_i64(row.numel(), row.device, group)
    return ShardedCSR(n, m, int(totals.sum()), list(bounds), row_ptr, ocol, ovals,
                      int(totals[:rank].sum()))


# ------------------------------------------------------------------------------- features
def degrees(ops, s: ShardedCSR, id_dtype=torch.int32, group=None):
    lo, hi = s.block(_world(group)[0])
    return ops.degrees(hi - lo, s.row_ptr, id_dtype=id_dtype)


def degree_distribution(ops, s: ShardedCSR, feature_dtype=torch.float32, group=None):
    rank, _ = _world(group)
    lo, hi = s.block(rank)
    return ops.degree_distribution(hi - lo, s.nnz, s.row_ptr, feature_dtype=feature_dtype)


# ------------------------------------------------------------------------------- DegreeReorder
def degree_reorder(ops, s: ShardedCSR, ascending=True, group=None, id_dtype=torch.int32):
    """Global DegreeReorder (inv[old] = new, ties by descending id / ascending when descending)
    = block-local rank + an offset per degree value that only needs the per-rank degree
    histograms.  Returns the FULL permutation on every rank."""
    rank, world = _world(group)
    lo, hi = s.block(rank)
    nl = hi - lo
    dev = s.row_ptr.device
    tr = _Trace("degree_reorder")
    local = ops.degree_reorder(nl, s.row_ptr, True, id_dtype=id_dtype)  # (deg asc, id desc)
    tr.mark("local")
    maxdeg = int(_all_gather_i64(ops.max_degree(nl, s.row_ptr), dev, group).max())
    nbins = maxdeg + 1
    hist = ops.degree_histogram(nl, s.row_ptr, nbins)                   # int64[nbins]
    if world > 1:
        allh = torch.empty(world * nbins, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allh, hist, group=group)
        allh = allh.view(world, nbins)
    else:
        allh = hist.view(1, nbins)
    tr.mark("hist_gather")
    # O(world * maxdeg) bookkeeping on the histograms (tiny next to n):
    total = allh.sum(0)
    g_start = torch.cumsum(total, 0) - total          # vertices with a smaller degree, anywhere
    later = allh[rank + 1:].sum(0)                    # same degree, larger ids (ranked first)
    l_start = torch.cumsum(hist, 0) - hist            # smaller degree inside this block
    offset = (g_start + later - l_start).contiguous()
    part = ops.degree_rank_combine(nl, s.row_ptr, local, offset,
                                   -1 if ascending else s.n - 1)
    tr.mark("combine")
    counts = [s.bounds[r + 1] - s.bounds[r] for r in range(world)]
    out = _all_gather_var(part, counts, group)
    tr.mark("allgather_perm")
    tr.done()
    return out


# ------------------------------------------------------------------------------- Permute2D
def permute2d(ops, s: ShardedCSR, row_order, col_order, group=None):
    """Sharded PermuteOrderTwo + CSR-constructor row sort.  row_order / col_order are the FULL
    inverse permutations (replicated); the result is sharded by nnz-balanced NEW-row blocks.

    sender    renumber + sort the columns of its rows and order the rows by new id (one local
              sb200_permute2d), so every destination's rows are one contiguous slice
    exchange  all_to_all of (new id, length) per row and (col, val) per nonzero
    receiver  place the received rows (one local sb200_permute2d with the row order only)
    """
    rank, world = _world(group)
    lo, hi = s.block(rank)
    nl = hi - lo
    dev = s.col.device
    idt = s.col.dtype
    # ---- nnz-balanced blocks of the NEW row space (needs every row's degree once)
    tr = _Trace("permute2d")
    deg_local = ops.degrees(nl, s.row_ptr, id_dtype=s.row_ptr.dtype)
    counts = [s.bounds[r + 1] - s.bounds[r] for r in range(world)]
    deg = _all_gather_var(deg_local, counts, group)
    new_deg = ops.permute1d(deg, row_order) if row_order is not None else deg
    new_ptr = ops.exclusive_scan(new_deg)
    nb = ops.partition_rows(s.n, s.nnz, new_ptr, world) if world > 1 else [0, s.n]
    tr.mark("balance")
    # ---- sender
    my_new = row_order[lo:hi].contiguous() if row_order is not None else \
        torch.arange(lo, hi, dtype=idt, device=dev)
    local_rank = ops.rank_keys(my_new, s.n)                        # order my rows by new id
    p_ptr, p_col, p_val = ops.permute2d(nl, s.m, s.row_ptr, s.col, s.vals, local_rank, col_order)
    tr.mark("local_permute")
    sorted_new = ops.permute1d(my_new, local_rank)                 # ascending new ids
    p_len = (p_ptr[1:] - p_ptr[:-1]).contiguous()
    cut = torch.searchsorted(sorted_new, torch.tensor(nb, dtype=idt, device=dev)).cpu().tolist()
    ptr_at = p_ptr[torch.tensor(cut, device=dev)].cpu().tolist()
    send_rows = [cut[q + 1] - cut[q] for q in range(world)]
    send_nnz = [ptr_at[q + 1] - ptr_at[q] for q in range(world)]
    recv_rows = _exchange_counts(send_rows, dev, group)
    recv_nnz = _exchange_counts(send_nnz, dev, group)
    tr.mark("cuts_counts")
    # ---- exchange
    r_new = _all_to_all_var(sorted_new, send_rows, recv_rows, group)
    r_len = _all_to_all_var(p_len, send_rows, recv_rows, group)
    r_col = _all_to_all_var(p_col, send_nnz, recv_nnz, group)
    r_val = None if p_val is None else _all_to_all_var(p_val, send_nnz, recv_nnz, group)
    tr.mark("all_to_all")
    # ---- receiver
    nlo, nhi = nb[rank], nb[rank + 1]
    r_ptr = ops.exclusive_scan(r_len)
    order = (r_new - nlo).to(idt)
    o_ptr, o_col, o_val = ops.permute2d(nhi - nlo, s.m, r_ptr, r_col, r_val, order, None)
    base = int(new_ptr[nlo])
    tr.mark("place")
    tr.done()
    return ShardedCSR(s.n, s.m, s.nnz, list(nb), o_ptr, o_col, o_val, base)


# ------------------------------------------------------------------------------- CSR -> CSC
def csr_to_csc(ops, s: ShardedCSR, group=None):
    """Sharded transpose of the layout: row-block sharded CSR -> column-block sharded CSC.

    local     CSR block -> CSC of the block over ALL columns (rows ascending inside a column)
    exchange  all_reduce of the column counts (global col_ptr, column blocks balanced by nnz),
              all_to_all of every destination's contiguous column range
    receiver  interleave the pieces: inside a column the source ranks come in rank order, which
              is ascending row order because row blocks are ordered
    """
    rank, world = _world(group)
    lo, hi = s.block(rank)
    dev = s.col.device
    tr = _Trace("csr_to_csc")
    cp, rows, vals = ops.csr_to_csc_block(lo, hi - lo, s.m, s.row_ptr, s.col, s.vals)
    tr.mark("local_csc")
    # column counts in the pointer type (the global col_ptr has to fit it anyway): for 32-bit
    # pointers the all_reduce moves half the bytes of an int64 one
    cnt = cp[1:] - cp[:-1]
    tot = cnt.clone()
    if world > 1:
        dist.all_reduce(tot, group=group)
    tr.mark("allreduce_counts")
    gptr = ops.exclusive_scan(tot)                                   # int64[m+1], replicated
    cb = ops.partition_rows(s.m, s.nnz, gptr, world) if world > 1 else [0, s.m]
    at = cp[torch.tensor(cb, device=dev)].cpu().tolist()
    tr.mark("scan_partition")
    send_cols = [cb[q + 1] - cb[q] for q in range(world)]
    send_nnz = [at[q + 1] - at[q] for q in range(world)]
    clo, chi = cb[rank], cb[rank + 1]
    ncl = chi - clo
    recv_cols = [ncl] * world
    recv_nnz = _exchange_counts(send_nnz, dev, group)
    r_cnt = _all_to_all_var(cnt, send_cols, recv_cols, group)
    r_row = _all_to_all_var(rows, send_nnz, recv_nnz, group)
    r_val = None if vals is None else _all_to_all_var(vals, send_nnz, recv_nnz, group)
    tr.mark("all_to_all")
    if world == 1:
        return ShardedCSC(s.n, s.m, s.nnz, list(cb), cp, rows, vals, 0)
    # segments arrive ordered (source, column); the result needs (column, source)
    seg_ptr = ops.exclusive_scan(r_cnt)
    order = _interleave_order(world, ncl, s.col.dtype, dev)          # (s, c) -> c * world + s
    tr.mark("order")
    o_ptr, o_row, o_val = ops.permute2d(world * ncl, s.n, seg_ptr, r_row, r_val, order, None)
    tr.mark("interleave")
    col_ptr = o_ptr[::world].contiguous()
    tr.mark("col_ptr")
    tr.done()
    return ShardedCSC(s.n, s.m, s.nnz, list(cb), col_ptr, o_row, o_val, int(gptr[clo]))
