"""ctypes binding of libsb200.so + thin torch-tensor wrappers.

The wrappers take CUDA torch tensors (or raw device pointers), pass ``data_ptr()`` and the
current torch stream to the C ABI and return freshly allocated CUDA tensors.  They mirror the
reference's operator names (``coo_to_csr`` = ``COO::Convert<CSR>``, ``degree_reorder`` =
``DegreeReorder::GetReorder`` ...).  No computation happens in Python.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsb200.so")

VOID, I32, I64, U32, U64, F32, F64 = range(7)
_DT = {torch.int32: I32, torch.int64: I64, torch.float32: F32, torch.float64: F64}

# every symbol include/sb200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "sb200_abi_version", "sb200_last_error", "sb200_device_count", "sb200_can_access_peer",
    "sb200_enable_peer_access", "sb200_malloc", "sb200_free", "sb200_malloc_host",
    "sb200_free_host", "sb200_memcpy_h2d", "sb200_memcpy_d2h", "sb200_memcpy_d2d",
    "sb200_stream_synchronize", "sb200_trim", "sb200_coo_sort", "sb200_compressed_sort", "sb200_coo_to_csr",
    "sb200_csr_to_coo", "sb200_coo_to_csc", "sb200_csr_to_csc", "sb200_degree_reorder",
    "sb200_rcm_reorder", "sb200_rcm_last_stats", "sb200_rcm_last_cycles",
    "sb200_rcm_last_resplits",
    "sb200_rcm_last_speculation", "sb200_permute2d", "sb200_permute1d",
    "sb200_inverse_permutation",
    "sb200_degrees", "sb200_degree_distribution", "sb200_degree_features", "sb200_reorder_heatmap", "sb200_boba_reorder", "sb200_edges_to_coo", "sb200_partition_rows", "sb200_launch_count",
    "sb200_reset_launch_count", "sb200_coo_to_csr_block", "sb200_csr_to_csc_block",
    "sb200_exclusive_scan", "sb200_rank_keys", "sb200_max_degree", "sb200_degree_histogram",
    "sb200_degree_rank_combine",
    "sb200_mg_comm_create", "sb200_mg_comm_connect", "sb200_mg_comm_create_local",
    "sb200_mg_comm_destroy", "sb200_mg_comm_info", "sb200_mg_run_ranks", "sb200_mg_barrier", "sb200_mg_allgather_i64",
    "sb200_mg_coo_to_csr", "sb200_mg_degree_reorder", "sb200_mg_permute2d_run",
    "sb200_mg_permute2d_fetch", "sb200_mg_csr_to_csc_run", "sb200_mg_csr_to_csc_fetch",
    "sb200_mg_permute1d",
]


class Sb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsb200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libsb200.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m sparsebase_b200.build` "
                "(there is no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.sb200_last_error.restype = ctypes.c_char_p
        _lib.sb200_launch_count.restype = ctypes.c_int64
        _lib.sb200_reset_launch_count.restype = None
    return _lib


def _check(rc):
    if rc != 0:
        raise Sb200Error(rc, load().sb200_last_error().decode())


def _p(t):
    if t is None:
        return ctypes.c_void_p(None)
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    return ctypes.c_void_p(t.data_ptr())


def _i64(x):
    return ctypes.c_int64(int(x))


def _dev(t):
    assert t.is_cuda, "sparsebase_b200 operates on CUDA tensors only (no CPU fallback)"
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(t=None):
    """torch's current stream OF THE DEVICE THE OPERANDS LIVE ON (not of the current device)."""
    dev = torch.cuda.current_device() if t is None else _dev(t)
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def trim(device=None):
    """Return the library's cached scratch blocks on `device` to the driver."""
    _check(load().sb200_trim(torch.cuda.current_device() if device is None else int(device)))


def _vt(vals):
    return VOID if vals is None else _DT[vals.dtype]


def launch_count():
    return int(load().sb200_launch_count())


def reset_launch_count():
    load().sb200_reset_launch_count()


def device_count():
    c = ctypes.c_int(0)
    _check(load().sb200_device_count(ctypes.byref(c)))
    return c.value


# ------------------------------------------------------------------ format constructors
def coo_sort_(n, m, row, col, vals=None):
    """COO constructor semantics, in place.  Returns True when the input was already sorted."""
    flag = ctypes.c_int(0)
    _check(load().sb200_coo_sort(_dev(row), _i64(n), _i64(m), _i64(row.numel()), _p(row), _p(col),
                                 _p(vals), _DT[row.dtype], _vt(vals), ctypes.byref(flag),
                                 _stream(row)))
    return bool(flag.value)


def compressed_sort_(n_seg, n_idx, ptr, idx, vals=None):
    """CSR/CSC constructor semantics (check + per-segment sort), in place."""
    flag = ctypes.c_int(0)
    _check(load().sb200_compressed_sort(_dev(idx), _i64(n_seg), _i64(n_idx), _i64(idx.numel()),
                                        _p(ptr), _p(idx), _p(vals), _DT[idx.dtype],
                                        _DT[ptr.dtype], _vt(vals), ctypes.byref(flag), _stream(idx)))
    return bool(flag.value)


# ------------------------------------------------------------------ conversions
def coo_to_csr(n, m, row, col, vals=None, nnz_dtype=torch.int32):
    nnz = row.numel()
    row_ptr = torch.empty(n + 1, dtype=nnz_dtype, device=row.device)
    ocol = torch.empty_like(col)
    ovals = None if vals is None else torch.empty_like(vals)
    _check(load().sb200_coo_to_csr(_dev(row), _i64(n), _i64(m), _i64(nnz), _p(row), _p(col),
                                   _p(vals), _p(row_ptr), _p(ocol), _p(ovals), _DT[row.dtype],
                                   _DT[nnz_dtype], _vt(vals), _stream(row)))
    return row_ptr, ocol, ovals


def csr_to_coo(n, m, row_ptr, col, vals=None):
    nnz = col.numel()
    orow = torch.empty_like(col)
    ocol = torch.empty_like(col)
    ovals = None if vals is None else torch.empty_like(vals)
    _check(load().sb200_csr_to_coo(_dev(col), _i64(n), _i64(m), _i64(nnz), _p(row_ptr), _p(col),
                                   _p(vals), _p(orow), _p(ocol), _p(ovals), _DT[col.dtype],
                                   _DT[row_ptr.dtype], _vt(vals), _stream(col)))
    return orow, ocol, ovals


def coo_to_csc(n, m, row, col, vals=None, nnz_dtype=torch.int32):
    nnz = row.numel()
    col_ptr = torch.empty(n + 1, dtype=nnz_dtype, device=row.device)
    orow = torch.empty_like(row)
    ovals = None if vals is None else torch.empty_like(vals)
    _check(load().sb200_coo_to_csc(_dev(row), _i64(n), _i64(m), _i64(nnz), _p(row), _p(col),
                                   _p(vals), _p(col_ptr), _p(orow), _p(ovals), _DT[row.dtype],
                                   _DT[nnz_dtype], _vt(vals), _stream(row)))
    return col_ptr, orow, ovals


def csr_to_csc(n, m, row_ptr, col, vals=None, out=None):
    nnz = col.numel()
    if out is None:
        col_ptr = torch.empty(n + 1, dtype=row_ptr.dtype, device=col.device)
        orow = torch.empty_like(col)
        ovals = None if vals is None else torch.empty_like(vals)
    else:
        col_ptr, orow, ovals = out
    _check(load().sb200_csr_to_csc(_dev(col), _i64(n), _i64(m), _i64(nnz), _p(row_ptr), _p(col),
                                   _p(vals), _p(col_ptr), _p(orow), _p(ovals), _DT[col.dtype],
                                   _DT[row_ptr.dtype], _vt(vals), _stream(col)))
    return col_ptr, orow, ovals


# ------------------------------------------------------------------ reorderings
def degree_reorder(n, row_ptr, ascending=True, id_dtype=torch.int32):
    inv = torch.empty(n, dtype=id_dtype, device=row_ptr.device)
    _check(load().sb200_degree_reorder(_dev(row_ptr), _i64(n), _p(row_ptr),
                                       ctypes.c_int(1 if ascending else 0), _p(inv),
                                       _DT[id_dtype], _DT[row_ptr.dtype], _stream(row_ptr)))
    return inv


def rcm_reorder(n, row_ptr, col):
    inv = torch.empty(n, dtype=col.dtype, device=col.device)
    _check(load().sb200_rcm_reorder(_dev(col), _i64(n), _i64(col.numel()), _p(row_ptr), _p(col),
                                    _p(inv), _DT[col.dtype], _DT[row_ptr.dtype], _stream(col)))
    return inv


def rcm_last_stats():
    out = (ctypes.c_int64 * 4)()
    _check(load().sb200_rcm_last_stats(out))
    d = dict(zip(("levels_narrow", "levels_wide", "bfs", "components"), list(out)))
    cyc = (ctypes.c_int64 * 8)()
    _check(load().sb200_rcm_last_cycles(cyc))
    d["phase_cycles"] = dict(zip(("claim", "barrier", "recheck", "compact", "sort_state"),
                                 list(cyc)[:5]))
    rs = (ctypes.c_int64 * 2)()
    _check(load().sb200_rcm_last_resplits(rs))
    d["resplits"], d["resizes"] = int(rs[0]), int(rs[1])
    spec = (ctypes.c_int64 * 3)()
    _check(load().sb200_rcm_last_speculation(spec))
    d["spec_confirmed"], d["spec_continued"], d["spec_replayed"] = (int(x) for x in spec)
    return d


# ------------------------------------------------------------------ permutation
def permute2d(n, m, row_ptr, col, vals, row_order, col_order, out=None):
    nnz = col.numel()
    if out is None:
        orp = torch.empty_like(row_ptr)
        ocol = torch.empty_like(col)
        ovals = None if vals is None else torch.empty_like(vals)
    else:
        orp, ocol, ovals = out
    _check(load().sb200_permute2d(_dev(col), _i64(n), _i64(m), _i64(nnz), _p(row_ptr), _p(col),
                                  _p(vals), _p(row_order), _p(col_order), _p(orp), _p(ocol),
                                  _p(ovals), _DT[col.dtype], _DT[row_ptr.dtype], _vt(vals),
                                  _stream(col)))
    return orp, ocol, ovals


def permute1d(vals, order):
    out = torch.empty_like(vals)
    _check(load().sb200_permute1d(_dev(vals), _i64(vals.numel()), _p(vals), _p(order), _p(out),
                                  _DT[order.dtype], _DT[vals.dtype], _stream(vals)))
    return out


def inverse_permutation(perm):
    out = torch.empty_like(perm)
    _check(load().sb200_inverse_permutation(_dev(perm), _i64(perm.numel()), _p(perm), _p(out),
                                            _DT[perm.dtype], _stream(perm)))
    return out


# ------------------------------------------------------------------ features
def degrees(n, row_ptr, id_dtype=torch.int32):
    out = torch.empty(n, dtype=id_dtype, device=row_ptr.device)
    _check(load().sb200_degrees(_dev(row_ptr), _i64(n), _p(row_ptr), _p(out), _DT[id_dtype],
                                _DT[row_ptr.dtype], _stream(row_ptr)))
    return out


def degree_distribution(n, nnz, row_ptr, feature_dtype=torch.float32):
    out = torch.empty(n, dtype=feature_dtype, device=row_ptr.device)
    _check(load().sb200_degree_distribution(_dev(row_ptr), _i64(n), _i64(nnz), _p(row_ptr),
                                            _p(out), _DT[row_ptr.dtype], _DT[feature_dtype],
                                            _stream(row_ptr)))
    return out


def degree_features(n, nnz, row_ptr, col=None, id_dtype=torch.int32,
                    feature_dtype=torch.float32, want_arrays=True):
    """Fused Degrees_DegreeDistribution + min/max/avg degree + Bandwidth + Profile.
    Returns (degrees, dist, dict(min_degree, max_degree, bandwidth, profile, avg_degree))."""
    deg = torch.empty(n, dtype=id_dtype, device=row_ptr.device) if want_arrays else None
    dist = torch.empty(n, dtype=feature_dtype, device=row_ptr.device) if want_arrays else None
    sc = (ctypes.c_int64 * 4)()
    avg = ctypes.c_double(0.0)
    _check(load().sb200_degree_features(_dev(row_ptr), _i64(n), _i64(nnz), _p(row_ptr), _p(col),
                                        _p(deg), _p(dist), sc, ctypes.byref(avg),
                                        _DT[id_dtype], _DT[row_ptr.dtype], _DT[feature_dtype],
                                        _stream(row_ptr)))
    out = dict(zip(("min_degree", "max_degree", "bandwidth", "profile"), (int(x) for x in sc)))
    out["avg_degree"] = avg.value
    return deg, dist, out


def boba_reorder(n, m, row, col):
    """BOBAReorder on a (row, col)-sorted device COO: inv[max(n, m)] (inv[v] = new position)."""
    inv = torch.empty(max(n, m), dtype=row.dtype, device=row.device)
    _check(load().sb200_boba_reorder(_dev(row), _i64(n), _i64(m), _i64(row.numel()), _p(row),
                                     _p(col), _p(inv), _DT[row.dtype], _stream(row)))
    return inv


def reorder_heatmap(n, m, row_ptr, col, order_r, order_c, num_parts=3,
                    feature_dtype=torch.float32):
    """ReorderHeatmap: num_parts x num_parts grid (row-major, device tensor) of the shares of the
    nonzeros per cell after renumbering rows / columns by order_r / order_c (None = identity).
    Raises Sb200Error (BAD_ARG) where the reference throws (num_parts > n or m)."""
    heat = torch.empty(max(1, num_parts * num_parts), dtype=feature_dtype, device=row_ptr.device)
    _check(load().sb200_reorder_heatmap(_dev(row_ptr), _i64(n), _i64(m), _i64(col.numel()),
                                        _p(row_ptr), _p(col), _p(order_r), _p(order_c),
                                        ctypes.c_int(num_parts), _p(heat), _DT[col.dtype],
                                        _DT[row_ptr.dtype], _DT[feature_dtype],
                                        _stream(row_ptr)))
    return heat[:num_parts * num_parts]


def edges_to_coo(u, v, w=None, remove_duplicates=True, remove_self_edges=False,
                 read_undirected=False, square=False):
    """EdgeListReader::ReadCOO on a device edge list: returns (n, m, row, col, vals), sorted by
    (row, col) and de-duplicated."""
    cap = max(1, u.numel() * (2 if read_undirected else 1))
    orow = torch.empty(cap, dtype=u.dtype, device=u.device)
    ocol = torch.empty(cap, dtype=u.dtype, device=u.device)
    oval = None if w is None else torch.empty(cap, dtype=w.dtype, device=u.device)
    out3 = (ctypes.c_int64 * 3)()
    _check(load().sb200_edges_to_coo(_dev(u), _i64(u.numel()), _p(u), _p(v), _p(w),
                                     ctypes.c_int(int(remove_duplicates)),
                                     ctypes.c_int(int(remove_self_edges)),
                                     ctypes.c_int(int(read_undirected)), ctypes.c_int(int(square)),
                                     _p(orow), _p(ocol), _p(oval), out3, _DT[u.dtype], _vt(w),
                                     _stream(u)))
    n, m, nnz = (int(x) for x in out3)
    return n, m, orow[:nnz], ocol[:nnz], None if oval is None else oval[:nnz]


def partition_rows(n, nnz, row_ptr, parts):
    bounds = (ctypes.c_int64 * (parts + 1))()
    _check(load().sb200_partition_rows(_dev(row_ptr), _i64(n), _i64(nnz), _p(row_ptr),
                                       _DT[row_ptr.dtype], ctypes.c_int(parts), bounds, _stream(row_ptr)))
    return list(bounds)


# ------------------------------------------------------------------ multi-GPU building blocks
def coo_to_csr_block(row_lo, n_local, m, row, col, vals=None, nnz_dtype=torch.int32):
    nnz = row.numel()
    row_ptr = torch.empty(n_local + 1, dtype=nnz_dtype, device=row.device)
    ocol = torch.empty_like(col)
    ovals = None if vals is None else torch.empty_like(vals)
    _check(load().sb200_coo_to_csr_block(_dev(row), _i64(row_lo), _i64(n_local), _i64(m),
                                         _i64(nnz), _p(row), _p(col), _p(vals), _p(row_ptr),
                                         _p(ocol), _p(ovals), _DT[row.dtype], _DT[nnz_dtype],
                                         _vt(vals), _stream(row)))
    return row_ptr, ocol, ovals


def csr_to_csc_block(row_lo, n_local, m, row_ptr, col, vals=None):
    nnz = col.numel()
    col_ptr = torch.empty(m + 1, dtype=row_ptr.dtype, device=col.device)
    orow = torch.empty_like(col)
    ovals = None if vals is None else torch.empty_like(vals)
    _check(load().sb200_csr_to_csc_block(_dev(col), _i64(row_lo), _i64(n_local), _i64(m),
                                         _i64(nnz), _p(row_ptr), _p(col), _p(vals), _p(col_ptr),
                                         _p(orow), _p(ovals), _DT[col.dtype], _DT[row_ptr.dtype],
                                         _vt(vals), _stream(col)))
    return col_ptr, orow, ovals


def exclusive_scan(x):
    out = torch.empty(x.numel() + 1, dtype=x.dtype, device=x.device)
    _check(load().sb200_exclusive_scan(_dev(x), _i64(x.numel()), _p(x), _p(out), _DT[x.dtype],
                                       _stream(x)))
    return out


def rank_keys(keys, key_bound):
    out = torch.empty_like(keys)
    _check(load().sb200_rank_keys(_dev(keys), _i64(keys.numel()), _p(keys), _i64(key_bound),
                                  _p(out), _DT[keys.dtype], _stream(keys)))
    return out


def max_degree(n, row_ptr):
    out = ctypes.c_int64(0)
    _check(load().sb200_max_degree(_dev(row_ptr), _i64(n), _p(row_ptr), _DT[row_ptr.dtype],
                                   ctypes.byref(out), _stream(row_ptr)))
    return out.value


def degree_histogram(n, row_ptr, nbins):
    out = torch.empty(nbins, dtype=torch.int64, device=row_ptr.device)
    _check(load().sb200_degree_histogram(_dev(row_ptr), _i64(n), _p(row_ptr), _DT[row_ptr.dtype],
                                         _i64(nbins), _p(out), _stream(row_ptr)))
    return out


def degree_rank_combine(n, row_ptr, local_rank, offset, flip_from=-1):
    out = torch.empty_like(local_rank)
    _check(load().sb200_degree_rank_combine(_dev(row_ptr), _i64(n), _p(row_ptr), _p(local_rank),
                                            _p(offset), _i64(flip_from), _p(out),
                                            _DT[local_rank.dtype], _DT[row_ptr.dtype], _stream(row_ptr)))
    return out
