"""Duplicate column ids inside a row (outside the reference's input contract, but its result is
defined): the CSR/CSC constructor sorts EVERY row by std::less<pair<IDType, ValueType>> -- column,
then value IN ITS REAL TYPE -- as soon as ANY row is unsorted, and leaves all rows untouched
otherwise (format/csr.cc:99-157).  Values here are negative / mixed-sign / signed integers, so an
implementation that compares bit patterns fails; row lengths cover every kernel tier
(<= 8, <= 64, on-chip bitonic, segmented radix).

CPU part: the restated oracle agrees with the compiled reference on these inputs.
GPU part: sb200_compressed_sort / sb200_permute2d / sb200_coo_to_csr against the oracle.
"""
import numpy as np
import pytest
import torch

import graphs
import oracle_lib

TYPES = [(np.int32, np.int32, np.float32), (np.int64, np.int64, np.float64),
         (np.int32, np.int32, np.int32)]
TYPE_IDS = ["i32_i32_f32", "i64_i64_f64", "i32_i32_i32"]
LENGTHS = [5, 40, 200, 3000]


def dup_matrix(maxlen, idt, nt, vt, seed, sorted_rows):
    """n x n CSR with many repeated column ids per row and DISTINCT mixed-sign values."""
    rng = np.random.default_rng(seed)
    n = 2003 if maxlen <= 200 else 4001
    deg = rng.integers(0, maxlen + 1, size=n) if maxlen <= 200 else rng.integers(0, 60, size=n)
    if maxlen > 200:
        deg[rng.integers(0, n, size=40)] = rng.integers(1025, maxlen + 1, size=40)
    deg[[1, n // 2]] = maxlen
    deg[rng.integers(0, n, size=n // 20)] = 0
    cols = []
    for d in deg:
        span = max(2, int(d) // 2)           # about half of the entries collide
        c = rng.integers(0, span, size=d) + rng.integers(0, n - span)
        cols.append(np.sort(c) if sorted_rows else c)
    col = np.concatenate(cols).astype(idt)
    row = np.repeat(np.arange(n), deg).astype(idt)
    nnz = len(col)
    mag = rng.permutation(nnz).astype(np.float64) + 1.0      # distinct magnitudes
    sign = np.where(rng.random(nnz) < 0.5, -1.0, 1.0)
    vals = (mag * sign).astype(vt)
    return n, graphs.csr_of(n, row, col, nt), row, col, vals


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(
        a.view(np.uint8), b.view(np.uint8))


# ------------------------------------------------------------------ CPU: oracle vs reference
needs_ref = pytest.mark.skipif(oracle_lib.reference() is None,
                               reason="oracle/_ref/libsbref.so not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("types", TYPES, ids=TYPE_IDS)
@pytest.mark.parametrize("maxlen", [5, 40, 200])
def test_restated_equals_reference_on_duplicates(types, maxlen):
    idt, nt, vt = types
    a, b = oracle_lib.restated(), oracle_lib.reference()
    rng = np.random.default_rng(maxlen)
    for sorted_rows in (False, True):
        n, rp, row, col, vals = dup_matrix(maxlen, idt, nt, vt, 10 + maxlen, sorted_rows)
        ra, rb = a.csr_ctor_sort(n, n, rp, col, vals), b.csr_ctor_sort(n, n, rp, col, vals)
        assert eq(ra[0], rb[0]) and eq(ra[1], rb[1])
        if sorted_rows:
            ro, co = rng.permutation(n).astype(idt), rng.permutation(n).astype(idt)
            for r_, c_ in ((ro, co), (ro, None), (None, None)):
                pa = a.permute2d(n, n, rp, col, vals, r_, c_)
                pb = b.permute2d(n, n, rp, col, vals, r_, c_)
                for x, y in zip(pa, pb):
                    assert eq(x, y)


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def sb():
    from sparsebase_b200 import lib
    lib.load()
    return lib


@pytest.fixture(scope="module")
def orc():
    return oracle_lib.restated()


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return None if t is None else t.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("types", TYPES, ids=TYPE_IDS)
@pytest.mark.parametrize("maxlen", LENGTHS)
def test_compressed_sort_duplicates_typed_values(sb, orc, types, maxlen):
    """sb200_compressed_sort: unsorted rows with duplicates -> every row in (col, value) order,
    negative values before positive ones; already sorted rows -> untouched."""
    idt, nt, vt = types
    for sorted_rows in (False, True):
        n, rp, row, col, vals = dup_matrix(maxlen, idt, nt, vt, 20 + maxlen, sorted_rows)
        ecol, evals = orc.csr_ctor_sort(n, n, rp, col, vals)
        dcol, dvals = dev(col), dev(vals)
        was_sorted = sb.compressed_sort_(n, n, dev(rp), dcol, dvals)
        assert was_sorted == sorted_rows
        assert eq(host(dcol), ecol), f"cols, maxlen={maxlen} sorted={sorted_rows}"
        assert eq(host(dvals), evals), f"vals, maxlen={maxlen} sorted={sorted_rows}"


@pytest.mark.gpu
@pytest.mark.parametrize("types", TYPES, ids=TYPE_IDS)
@pytest.mark.parametrize("maxlen", LENGTHS)
def test_permute2d_duplicates_typed_values(sb, orc, types, maxlen):
    """sb200_permute2d on rows with duplicate ids: a column renumbering makes some row unsorted
    -> (col, value) order everywhere; a row-only (or identity) permutation of sorted rows leaves
    nothing unsorted -> the reference does not sort and duplicates keep their source order."""
    idt, nt, vt = types
    n, rp, row, col, vals = dup_matrix(maxlen, idt, nt, vt, 30 + maxlen, True)
    rng = np.random.default_rng(99 + maxlen)
    ro, co = rng.permutation(n).astype(idt), rng.permutation(n).astype(idt)
    for r_, c_ in ((ro, co), (ro, ro), (None, co), (ro, None), (None, None)):
        exp = orc.permute2d(n, n, rp, col, vals, r_, c_)
        got = sb.permute2d(n, n, dev(rp), dev(col), dev(vals), dev(r_), dev(c_))
        for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
            assert eq(host(a), b), (f"permute2d {what} maxlen={maxlen} "
                                    f"row_order={'y' if r_ is not None else 'n'} "
                                    f"col_order={'y' if c_ is not None else 'n'}")


@pytest.mark.gpu
@pytest.mark.parametrize("maxlen", LENGTHS)
def test_coo_to_csr_duplicates(sb, orc, maxlen):
    """A (row, col)-sorted COO with repeated pairs: the conversion copies col / vals verbatim
    and the CSR constructor finds nothing to sort (converter_order_two.cc:181-201)."""
    idt, nt, vt = TYPES[0]
    n, rp, row, col, vals = dup_matrix(maxlen, idt, nt, vt, 40 + maxlen, True)
    exp = orc.coo_to_csr(n, n, row, col, vals)
    got = sb.coo_to_csr(n, n, dev(row), dev(col), dev(vals))
    for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
        assert eq(host(a), b), f"coo_to_csr {what} maxlen={maxlen}"
