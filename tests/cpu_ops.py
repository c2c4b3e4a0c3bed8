"""TEST INFRASTRUCTURE: a CPU stand-in for sparsebase_b200.lib with the same function names,
built on the parity oracle (tests/oracle_lib.py) and numpy.  It exists so that the HOST logic
of sparsebase_b200/sharded.py (partitioning, offsets, the collectives) can run under gloo with
world_size 2 in a container without a GPU.  The product never imports this module."""
import numpy as np
import torch

import oracle_lib

_o = oracle_lib.restated()


def _np(t):
    return None if t is None else t.numpy()


def _t(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a))


def partition_rows(n, nnz, row_ptr, parts):
    rp = _np(row_ptr).astype(np.int64)
    b = [0]
    for k in range(1, parts):
        b.append(int(np.searchsorted(rp[:n], (nnz * k) // parts, side="left")))
    return b + [n]


def coo_sort_(n, m, row, col, vals=None):
    r, c, v = _o.coo_ctor_sort(n, m, _np(row), _np(col), _np(vals))
    row.copy_(_t(r)), col.copy_(_t(c))
    if vals is not None:
        vals.copy_(_t(v))


def coo_to_csr_block(row_lo, n_local, m, row, col, vals=None, nnz_dtype=torch.int32):
    r = _np(row) - np.int32(row_lo)
    dim = max(n_local, m, 1)  # the oracle follows the reference's square-ish layout
    rp, c, v = _o.coo_to_csr(dim, dim, r.astype(_np(row).dtype), _np(col), _np(vals))
    return _t(rp[: n_local + 1].copy()), _t(c), _t(v)


def degrees(n, row_ptr, id_dtype=torch.int32):
    return (row_ptr[1:] - row_ptr[:-1]).to(id_dtype)


def degree_distribution(n, nnz, row_ptr, feature_dtype=torch.float32):
    d = _np(row_ptr[1:] - row_ptr[:-1])
    ft = np.float32 if feature_dtype == torch.float32 else np.float64
    return _t(d.astype(ft) / ft(nnz))


def degree_reorder(n, row_ptr, ascending=True, id_dtype=torch.int32):
    # (degree ascending, id descending) == the reference's bucket fill.  The
    # oracle's counting sort indexes by degree and needs degree <= n, which does not hold for
    # a row BLOCK of a wider matrix, so the block-local stand-in sorts directly.
    deg = np.diff(_np(row_ptr).astype(np.int64))
    ids = np.arange(n)
    order = np.lexsort((-ids, deg)) if ascending else np.lexsort((ids, -deg))
    inv = np.empty(n, dtype=np.int32)
    inv[order] = np.arange(n, dtype=np.int32)
    return _t(inv)


def max_degree(n, row_ptr):
    return int((row_ptr[1:] - row_ptr[:-1]).max()) if n > 0 else 0


def degree_histogram(n, row_ptr, nbins):
    return torch.bincount((row_ptr[1:] - row_ptr[:-1]).to(torch.int64), minlength=nbins)[:nbins]


def degree_rank_combine(n, row_ptr, local_rank, offset, flip_from=-1):
    d = (row_ptr[1:] - row_ptr[:-1]).to(torch.int64)
    g = local_rank.to(torch.int64) + offset[d]
    if flip_from >= 0:
        g = flip_from - g
    return g.to(local_rank.dtype)


def permute1d(vals, order):
    out = torch.empty_like(vals)
    out[order.to(torch.int64)] = vals
    return out


def exclusive_scan(x):
    out = torch.zeros(x.numel() + 1, dtype=x.dtype)
    out[1:] = torch.cumsum(x, 0)
    return out


def rank_keys(keys, key_bound):
    rank = torch.empty_like(keys)
    rank[torch.argsort(keys.to(torch.int64))] = torch.arange(keys.numel(), dtype=keys.dtype)
    return rank


def permute2d(n, m, row_ptr, col, vals, row_order, col_order, out=None):
    ro = np.arange(n, dtype=_np(col).dtype) if row_order is None else _np(row_order)
    co = np.arange(m, dtype=_np(col).dtype) if col_order is None else _np(col_order)
    rp, c, v = _o.permute2d(n, m, _np(row_ptr), _np(col), _np(vals), ro, co)
    return _t(rp), _t(c), _t(v)


def csr_to_csc_block(row_lo, n_local, m, row_ptr, col, vals=None):
    rp, c = _np(row_ptr), _np(col)
    rows = np.repeat(np.arange(n_local, dtype=c.dtype), np.diff(rp)) + c.dtype.type(row_lo)
    order = np.argsort(c, kind="stable")
    cp = np.zeros(m + 1, dtype=rp.dtype)
    cp[1:] = np.cumsum(np.bincount(c, minlength=m))
    return _t(cp), _t(rows[order]), None if vals is None else _t(_np(vals)[order])
