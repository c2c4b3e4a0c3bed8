"""Multi-GPU operators over peer memory: thin torch wrappers of the sb200_mg_* C entry points.

One process per GPU.  ``Comm`` creates this rank's window, all-gathers the 64-byte CUDA IPC
handles with ``torch.distributed`` (any backend: it is plumbing, 64 bytes per rank, once) and
connects; after that every exchange of the operators below is a kernel storing straight into the
destination GPU's window over NVLink -- ``torch.distributed`` is not on the data path.

The data model is :class:`sparsebase_b200.sharded.ShardedCSR` / ``ShardedCSC`` (row / column
blocks with block-local pointers and global ids); results are bit-identical to the single-GPU
operators and to the collective-based implementation in :mod:`sparsebase_b200.sharded`.
"""
import ctypes

import torch
import torch.distributed as dist

from . import lib
from .lib import _DT, _check, _i64, _p, _stream, _vt
from .sharded import ShardedCSC, ShardedCSR


class Comm:
    """This rank's end of a peer-memory communicator (sb200_mg_comm_t)."""

    def __init__(self, window_bytes, device=None, group=None):
        self.device = torch.cuda.current_device() if device is None else int(device)
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.window_bytes = int(window_bytes)
        self._h = ctypes.c_void_p(None)
        handle = (ctypes.c_ubyte * 64)()
        _check(lib.load().sb200_mg_comm_create(
            ctypes.c_int(self.device), ctypes.c_int(self.rank), ctypes.c_int(self.world),
            ctypes.c_size_t(self.window_bytes), ctypes.byref(self._h), handle))
        if self.world > 1:
            every = [None] * self.world
            dist.all_gather_object(every, bytes(handle), group=group)
            blob = (ctypes.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(every))
            _check(lib.load().sb200_mg_comm_connect(self._h, blob))
        else:
            _check(lib.load().sb200_mg_comm_connect(self._h, None))

    def barrier(self):
        _check(lib.load().sb200_mg_barrier(self._h, _stream()))

    def allgather_i64(self, value):
        out = (ctypes.c_int64 * self.world)()
        _check(lib.load().sb200_mg_allgather_i64(self._h, _i64(value), out, _stream()))
        return list(out)

    def destroy(self):
        if self._h:
            _check(lib.load().sb200_mg_comm_destroy(self._h))
            self._h = ctypes.c_void_p(None)


def _bounds(b):
    return (ctypes.c_int64 * len(b))(*[int(x) for x in b])


def coo_to_csr(comm, n, m, bounds, row, col, vals, nnz_dtype=torch.int32, copy=True,
               presorted=False):
    """`row/col/vals`: all nonzeros of rows [bounds[rank], bounds[rank+1]) in any order
    (``presorted``: already (row, col)-sorted, i.e. a constructed format::COO -- the constructor's
    check / sort is skipped and the arrays are only read)."""
    lo, hi = bounds[comm.rank], bounds[comm.rank + 1]
    if not presorted:
        if copy:
            row, col = row.clone(), col.clone()
            vals = None if vals is None else vals.clone()
        lib.coo_sort_(n, m, row, col, vals)                   # format::COO constructor
    nnz_l = row.numel()
    row_ptr = torch.empty(hi - lo + 1, dtype=nnz_dtype, device=row.device)
    ocol = torch.empty_like(col)
    ovals = None if vals is None else torch.empty_like(vals)
    out2 = (ctypes.c_int64 * 2)()
    _check(lib.load().sb200_mg_coo_to_csr(
        comm._h, _i64(lo), _i64(hi - lo), _i64(m), _i64(nnz_l), _p(row), _p(col), _p(vals),
        _p(row_ptr), _p(ocol), _p(ovals), out2, _DT[row.dtype], _DT[nnz_dtype], _vt(vals),
        _stream(row)))
    return ShardedCSR(n, m, int(out2[0]), list(bounds), row_ptr, ocol, ovals, int(out2[1]))


def degree_reorder(comm, s: ShardedCSR, ascending=True, id_dtype=torch.int32):
    inv = torch.empty(s.n, dtype=id_dtype, device=s.row_ptr.device)
    _check(lib.load().sb200_mg_degree_reorder(
        comm._h, _i64(s.n), _bounds(s.bounds), _p(s.row_ptr), ctypes.c_int(1 if ascending else 0),
        _p(inv), _DT[id_dtype], _DT[s.row_ptr.dtype], _stream(s.row_ptr)))
    return inv


def permute2d(comm, s: ShardedCSR, row_order, col_order):
    """Sharded PermuteOrderTwo + CSR-constructor row sort: the result is sharded by nnz-balanced
    blocks of the NEW rows."""
    nb = (ctypes.c_int64 * (comm.world + 1))()
    out2 = (ctypes.c_int64 * 3)()
    vt = _vt(s.vals)
    _check(lib.load().sb200_mg_permute2d_run(
        comm._h, _i64(s.n), _i64(s.m), _i64(s.nnz), _bounds(s.bounds), _p(s.row_ptr), _p(s.col),
        _p(s.vals), _p(row_order), _p(col_order), nb, out2, _DT[s.col.dtype],
        _DT[s.row_ptr.dtype], vt, _stream(s.col)))
    rows, nnz_l = int(out2[0]), int(out2[1])
    dev = s.col.device
    orp = torch.empty(rows + 1, dtype=s.row_ptr.dtype, device=dev)
    ocol = torch.empty(nnz_l, dtype=s.col.dtype, device=dev)
    oval = None if s.vals is None else torch.empty(nnz_l, dtype=s.vals.dtype, device=dev)
    _check(lib.load().sb200_mg_permute2d_fetch(
        comm._h, _i64(s.n), _i64(rows), _i64(nnz_l), _p(orp), _p(ocol), _p(oval),
        _DT[s.col.dtype], _DT[s.row_ptr.dtype], vt, _stream(s.col)))
    return ShardedCSR(s.n, s.m, s.nnz, [int(x) for x in nb], orp, ocol, oval, int(out2[2]))


def csr_to_csc(comm, s: ShardedCSR):
    """Sharded transpose of the layout: column-block sharded CSC."""
    cb = (ctypes.c_int64 * (comm.world + 1))()
    out2 = (ctypes.c_int64 * 3)()
    vt = _vt(s.vals)
    _check(lib.load().sb200_mg_csr_to_csc_run(
        comm._h, _i64(s.n), _i64(s.m), _i64(s.nnz), _bounds(s.bounds), _p(s.row_ptr), _p(s.col),
        _p(s.vals), cb, out2, _DT[s.col.dtype], _DT[s.row_ptr.dtype], vt, _stream(s.col)))
    ncl, nnz_l = int(out2[0]), int(out2[1])
    dev = s.col.device
    cp = torch.empty(ncl + 1, dtype=s.row_ptr.dtype, device=dev)
    orow = torch.empty(nnz_l, dtype=s.col.dtype, device=dev)
    oval = None if s.vals is None else torch.empty(nnz_l, dtype=s.vals.dtype, device=dev)
    cb = [int(x) for x in cb]
    _check(lib.load().sb200_mg_csr_to_csc_fetch(
        comm._h, _i64(s.m), _i64(cb[comm.rank]), _i64(ncl), _p(cp), _p(orow), _p(oval),
        _DT[s.col.dtype], _DT[s.row_ptr.dtype], vt, _stream(s.col)))
    return ShardedCSC(s.n, s.m, s.nnz, cb, cp, orow, oval, int(out2[2]))


def permute1d(comm, bounds, vals, order):
    """vals / order: this rank's block of the arrays; returns this rank's block of
    out[order[i]] = vals[i]."""
    out = torch.empty_like(vals)
    _check(lib.load().sb200_mg_permute1d(comm._h, _bounds(bounds), _p(vals), _p(order), _p(out),
                                         _DT[order.dtype], _DT[vals.dtype], _stream(vals)))
    return out
