"""TEST INFRASTRUCTURE: ctypes bindings for the parity oracle.

* ``oracle/liboracle.so``      -- plain-C restatement (prefix ``sbo_``), always available.
* ``oracle/_ref/libsbref.so``  -- the unmodified reference compiled header-only
  (prefix ``sbref_``); present when it was built in the dev container.

Both expose the same entry points, so a test can run either through ``Oracle``.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_TAG = {np.dtype(np.int32): "i32", np.dtype(np.int64): "i64", np.dtype(np.float32): "f32",
        np.dtype(np.float64): "f64"}
SUPPORTED_TAGS = ("i32_i32_f32", "i32_i64_f32", "i64_i64_f64", "i32_i32_i32", "i32_i32_void")
_FEAT = {"i32_i32_f32": np.float32, "i32_i64_f32": np.float32, "i64_i64_f64": np.float64,
         "i32_i32_i32": np.float32, "i32_i32_void": np.float32}


def _ptr(a):
    return ctypes.c_void_p(None) if a is None else ctypes.c_void_p(a.ctypes.data)


def _c(a):
    return None if a is None else np.ascontiguousarray(a)


def tag_of(idt, nt, vt):
    t = f"{_TAG[np.dtype(idt)]}_{_TAG[np.dtype(nt)]}_" + ("void" if vt is None else _TAG[np.dtype(vt)])
    if t not in SUPPORTED_TAGS:
        raise ValueError(f"oracle has no instantiation for {t}")
    return t


class Oracle:
    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        self.path = path

    def _fn(self, name):
        f = getattr(self.lib, f"{self.prefix}_{name}")
        f.restype = ctypes.c_int
        return f

    @staticmethod
    def _i64(x):
        return ctypes.c_int64(int(x))

    # ---- format constructors -------------------------------------------------------
    def coo_ctor_sort(self, n, m, row, col, vals, nnz_dtype=np.int32):
        """COO ctor (in place on copies). Returns (row, col, vals)."""
        row, col = _c(row).copy(), _c(col).copy()
        vals = None if vals is None else _c(vals).copy()
        t = tag_of(row.dtype, nnz_dtype, None if vals is None else vals.dtype)
        rc = self._fn(f"coo_ctor_sort_{t}")(self._i64(n), self._i64(m), self._i64(len(row)),
                                            _ptr(row), _ptr(col), _ptr(vals))
        assert rc == 0
        return row, col, vals

    def csr_ctor_sort(self, n, m, row_ptr, col, vals):
        col = _c(col).copy()
        vals = None if vals is None else _c(vals).copy()
        row_ptr = _c(row_ptr)
        t = tag_of(col.dtype, row_ptr.dtype, None if vals is None else vals.dtype)
        rc = self._fn(f"csr_ctor_sort_{t}")(self._i64(n), self._i64(m), _ptr(row_ptr), _ptr(col),
                                            _ptr(vals))
        assert rc == 0
        return col, vals

    # ---- conversions ---------------------------------------------------------------
    def coo_to_csr(self, n, m, row, col, vals, nnz_dtype=np.int32):
        row, col = _c(row).copy(), _c(col).copy()
        vals = None if vals is None else _c(vals).copy()
        nnz = len(row)
        t = tag_of(row.dtype, nnz_dtype, None if vals is None else vals.dtype)
        orp = np.empty(n + 1, nnz_dtype)
        oc = np.empty(nnz, row.dtype)
        ov = None if vals is None else np.empty(nnz, vals.dtype)
        rc = self._fn(f"coo_to_csr_{t}")(self._i64(n), self._i64(m), self._i64(nnz), _ptr(row),
                                         _ptr(col), _ptr(vals), _ptr(orp), _ptr(oc), _ptr(ov))
        assert rc == 0
        return orp, oc, ov

    def coo_to_csc(self, n, m, row, col, vals, nnz_dtype=np.int32):
        row, col = _c(row).copy(), _c(col).copy()
        vals = None if vals is None else _c(vals).copy()
        nnz = len(row)
        t = tag_of(row.dtype, nnz_dtype, None if vals is None else vals.dtype)
        ocp = np.empty(n + 1, nnz_dtype)
        orow = np.empty(nnz, row.dtype)
        ov = None if vals is None else np.empty(nnz, vals.dtype)
        rc = self._fn(f"coo_to_csc_{t}")(self._i64(n), self._i64(m), self._i64(nnz), _ptr(row),
                                         _ptr(col), _ptr(vals), _ptr(ocp), _ptr(orow), _ptr(ov))
        assert rc == 0
        return ocp, orow, ov

    def csr_to_csc(self, n, m, row_ptr, col, vals):
        row_ptr, col, vals = _c(row_ptr), _c(col), _c(vals)
        nnz = int(row_ptr[n])
        t = tag_of(col.dtype, row_ptr.dtype, None if vals is None else vals.dtype)
        ocp = np.empty(n + 1, row_ptr.dtype)
        orow = np.empty(nnz, col.dtype)
        ov = None if vals is None else np.empty(nnz, vals.dtype)
        rc = self._fn(f"csr_to_csc_{t}")(self._i64(n), self._i64(m), _ptr(row_ptr), _ptr(col),
                                         _ptr(vals), _ptr(ocp), _ptr(orow), _ptr(ov))
        assert rc == 0
        return ocp, orow, ov

    def csr_to_coo(self, n, m, row_ptr, col, vals):
        row_ptr, col, vals = _c(row_ptr), _c(col), _c(vals)
        nnz = int(row_ptr[n])
        t = tag_of(col.dtype, row_ptr.dtype, None if vals is None else vals.dtype)
        orow = np.empty(nnz, col.dtype)
        oc = np.empty(nnz, col.dtype)
        ov = None if vals is None else np.empty(nnz, vals.dtype)
        rc = self._fn(f"csr_to_coo_{t}")(self._i64(n), self._i64(m), _ptr(row_ptr), _ptr(col),
                                         _ptr(vals), _ptr(orow), _ptr(oc), _ptr(ov))
        assert rc == 0
        return orow, oc, ov

    # ---- reorderings ---------------------------------------------------------------
    def degree_reorder(self, n, row_ptr, col, ascending=True, vals=None):
        row_ptr, col = _c(row_ptr), _c(col)
        if n < 32768 and self.prefix == "sbref":
            pass  # harness sets M_MMAP_THRESHOLD=4096 (reference OOB, degree_reorder.cc:41-45)
        t = tag_of(col.dtype, row_ptr.dtype, None if vals is None else vals.dtype)
        inv = np.empty(n, col.dtype)
        rc = self._fn(f"degree_reorder_{t}")(self._i64(n), self._i64(n), _ptr(row_ptr), _ptr(col),
                                             _ptr(vals), ctypes.c_int(1 if ascending else 0),
                                             _ptr(inv))
        assert rc == 0
        return inv

    def rcm_reorder(self, n, row_ptr, col, vals=None):
        row_ptr, col = _c(row_ptr), _c(col)
        t = tag_of(col.dtype, row_ptr.dtype, None if vals is None else vals.dtype)
        inv = np.empty(n, col.dtype)
        rc = self._fn(f"rcm_reorder_{t}")(self._i64(n), self._i64(n), _ptr(row_ptr), _ptr(col),
                                          _ptr(vals), _ptr(inv))
        assert rc == 0
        return inv

    # ---- permutation ---------------------------------------------------------------
    def permute2d(self, n, m, row_ptr, col, vals, row_order, col_order):
        row_ptr, col, vals = _c(row_ptr), _c(col), _c(vals)
        same = row_order is col_order
        row_order = _c(row_order)
        col_order = row_order if same else _c(col_order)
        nnz = int(row_ptr[n])
        t = tag_of(col.dtype, row_ptr.dtype, None if vals is None else vals.dtype)
        orp = np.empty(n + 1, row_ptr.dtype)
        oc = np.empty(nnz, col.dtype)
        ov = None if vals is None else np.empty(nnz, vals.dtype)
        rc = self._fn(f"permute2d_{t}")(self._i64(n), self._i64(m), _ptr(row_ptr), _ptr(col),
                                        _ptr(vals), _ptr(row_order), _ptr(col_order), _ptr(orp),
                                        _ptr(oc), _ptr(ov))
        assert rc == 0
        return orp, oc, ov

    def permute1d(self, vals, order):
        vals, order = _c(vals), _c(order)
        t = f"{_TAG[order.dtype]}_{_TAG[vals.dtype]}"
        out = np.empty_like(vals)
        rc = self._fn(f"permute1d_{t}")(self._i64(len(vals)), _ptr(vals), _ptr(order), _ptr(out))
        assert rc == 0
        return out

    def inverse_permutation(self, perm):
        perm = _c(perm)
        out = np.empty_like(perm)
        rc = self._fn(f"inverse_permutation_{_TAG[perm.dtype]}")(self._i64(len(perm)), _ptr(perm),
                                                                 _ptr(out))
        assert rc == 0
        return out

    # ---- features ------------------------------------------------------------------
    def degrees(self, n, row_ptr, col, vals=None):
        row_ptr, col = _c(row_ptr), _c(col)
        t = tag_of(col.dtype, row_ptr.dtype, None if vals is None else vals.dtype)
        out = np.empty(n, col.dtype)
        rc = self._fn(f"degrees_{t}")(self._i64(n), self._i64(n), _ptr(row_ptr), _ptr(col),
                                      _ptr(vals), _ptr(out))
        assert rc == 0
        return out

    def degree_distribution(self, n, row_ptr, col, vals=None):
        row_ptr, col = _c(row_ptr), _c(col)
        t = tag_of(col.dtype, row_ptr.dtype, None if vals is None else vals.dtype)
        out = np.empty(n, _FEAT[t])
        rc = self._fn(f"degree_distribution_{t}")(self._i64(n), self._i64(n), _ptr(row_ptr),
                                                  _ptr(col), _ptr(vals), _ptr(out))
        assert rc == 0
        return out

    def degree_features(self, n, row_ptr, col, vals_dtype=np.float32):
        """(degrees, dist, {min_degree, max_degree, bandwidth, profile}, avg_degree)."""
        row_ptr, col = _c(row_ptr), _c(col)
        t = tag_of(col.dtype, row_ptr.dtype, vals_dtype)
        deg = np.empty(n, col.dtype)
        dist = np.empty(n, _FEAT[t])
        sc = (ctypes.c_int64 * 4)()
        avg = np.zeros(1, _FEAT[t])
        rc = self._fn(f"degree_features_{t}")(self._i64(n), self._i64(n), _ptr(row_ptr), _ptr(col),
                                              _ptr(deg), _ptr(dist), sc, _ptr(avg))
        assert rc == 0
        return deg, dist, dict(zip(("min_degree", "max_degree", "bandwidth", "profile"),
                                   (int(x) for x in sc))), avg[0]

    def boba_reorder(self, n, m, row, col, sequential=True, nnz_dtype=np.int32,
                     vals_dtype=np.float32):
        """BOBAReorder on a COO: inv[max(n, m)] (`sequential` only selects the reference's
        variant; the restatement has one form)."""
        row, col = _c(row), _c(col)
        t = tag_of(row.dtype, nnz_dtype, vals_dtype)
        out = np.empty(max(n, m), row.dtype)
        if self.prefix == "sbref":
            rc = self._fn(f"boba_reorder_{t}")(self._i64(n), self._i64(m), self._i64(len(row)),
                                               _ptr(row), _ptr(col), ctypes.c_int(int(sequential)),
                                               _ptr(out))
        else:
            rc = self._fn(f"boba_reorder_{t}")(self._i64(n), self._i64(m), self._i64(len(row)),
                                               _ptr(row), _ptr(col), _ptr(out))
        assert rc == 0
        return out

    def reorder_heatmap(self, n, row_ptr, col, order_r, order_c, num_parts,
                        vals_dtype=np.float32):
        """ReorderHeatmap: the num_parts x num_parts density grid (FeatureType of the type set),
        or None where the reference throws (num_parts larger than a dimension)."""
        row_ptr, col, order_r, order_c = _c(row_ptr), _c(col), _c(order_r), _c(order_c)
        t = tag_of(col.dtype, row_ptr.dtype, vals_dtype)
        out = np.zeros(max(1, num_parts * num_parts), _FEAT[t])
        rc = self._fn(f"reorder_heatmap_{t}")(self._i64(n), self._i64(n), _ptr(row_ptr),
                                              _ptr(col), _ptr(order_r), _ptr(order_c),
                                              ctypes.c_int(num_parts), _ptr(out))
        if rc == 1:
            return None
        assert rc == 0
        return out[:num_parts * num_parts]

    def edges_to_coo(self, u, v, w=None, remove_duplicates=True, remove_self=False,
                     undirected=False, square=False, nnz_dtype=np.int32):
        """EdgeListReader::ReadCOO semantics on arrays: returns (n, m, row, col, vals)."""
        u, v, w = _c(u), _c(v), _c(w)
        t = tag_of(u.dtype, nnz_dtype, None if w is None else w.dtype)
        cap = max(1, len(u) * (2 if undirected else 1))
        orow, ocol = np.empty(cap, u.dtype), np.empty(cap, u.dtype)
        ov = None if w is None else np.empty(cap, w.dtype)
        out3 = (ctypes.c_int64 * 3)()
        rc = self._fn(f"edges_to_coo_{t}")(self._i64(len(u)), _ptr(u), _ptr(v), _ptr(w),
                                           ctypes.c_int(int(remove_duplicates)),
                                           ctypes.c_int(int(remove_self)),
                                           ctypes.c_int(int(undirected)), ctypes.c_int(int(square)),
                                           _ptr(orow), _ptr(ocol), _ptr(ov), out3)
        assert rc == 0
        n, m, nnz = (int(x) for x in out3)
        return n, m, orow[:nnz].copy(), ocol[:nnz].copy(), None if ov is None else ov[:nnz].copy()


def build_oracle():
    """Compile oracle/liboracle.so (and _ref/libsbref.so when /root/reference is mounted)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


_cache = {}


def restated():
    """The plain-C restatement (always available; built on demand with gcc)."""
    if "sbo" not in _cache:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("sb_oracle.c", "sb_oracle_impl.h")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
            subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory", "liboracle.so"],
                           check=True, stdout=subprocess.DEVNULL)
        _cache["sbo"] = Oracle(path, "sbo")
    return _cache["sbo"]


def reference():
    """The compiled unmodified reference, or None when oracle/_ref was not built."""
    if "sbref" not in _cache:
        path = os.path.join(ORACLE_DIR, "_ref", "libsbref.so")
        _cache["sbref"] = Oracle(path, "sbref") if os.path.exists(path) else None
    return _cache["sbref"]
