"""GPU parity tests (-m gpu) aimed at specific kernel paths rather than at graph families:

* long runs of empty segments (head / middle / tail) -> the deferred gap fill of the boundary
  kernels (convert.cu gap_fill_kernel), also through the row-block CSR->CSC entry point;
* operands that are 4-byte but not 16-byte aligned -> the scalar boundary kernel, the
  register-prefetch radix downsweep and the scalar degree-feature path;
* record counts around the 4096-record radix tile -> the padded partial tile of
  rs_downsweep_pipe_kernel;
* duplicate column ids in short rows -> the (col, val) tie rule of the CSR constructor
  (format/csr.cc:147-148) in permute_short_rows_kernel;
* row counts around the 4-rows-per-thread vector path of the degree features.

Everything is compared bit-for-bit with the oracle through the C ABI.
"""
import numpy as np
import pytest
import torch

import graphs
import oracle_lib

pytestmark = pytest.mark.gpu

TT = {np.int32: torch.int32, np.int64: torch.int64, np.float32: torch.float32,
      np.float64: torch.float64}


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


def eq(a, b):
    if a is None or b is None:
        return a is None and b is None
    a, b = np.asarray(a), np.asarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(
        a.view(np.uint8), b.view(np.uint8))


def unaligned(a, off=1):
    """Device copy of `a` that starts `off` elements into a larger allocation."""
    if a is None:
        return None
    buf = torch.empty(len(a) + off + 8, dtype=TT[a.dtype.type], device="cuda")
    v = buf[off:off + len(a)]
    v.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    assert v.data_ptr() % 16 != 0 or len(a) == 0
    return v


def check_all_conversions(sb, orc, n, row, col, vals, nt=np.int32, wrap=dev):
    exp = orc.coo_to_csr(n, n, row, col, vals, nt)
    got = sb.coo_to_csr(n, n, wrap(row), wrap(col), wrap(vals), TT[nt])
    for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
        assert eq(host(a), b), f"coo_to_csr {what}"
    rp, cc, vv = exp
    exp = orc.csr_to_csc(n, n, rp, cc, vv)
    got = sb.csr_to_csc(n, n, wrap(rp), wrap(cc), wrap(vv))
    for a, b, what in zip(got, exp, ("col_ptr", "row", "vals")):
        assert eq(host(a), b), f"csr_to_csc {what}"
    got = sb.coo_to_csc(n, n, wrap(row), wrap(col), wrap(vals), TT[nt])
    for a, b, what in zip(got, exp, ("col_ptr", "row", "vals")):
        assert eq(host(a), b), f"coo_to_csc {what}"
    exp = orc.csr_to_coo(n, n, rp, cc, vv)
    got = sb.csr_to_coo(n, n, wrap(rp), wrap(cc), wrap(vv))
    for a, b, what in zip(got, exp, ("row", "col", "vals")):
        assert eq(host(a), b), f"csr_to_coo {what}"
    return rp, cc, vv


def _gappy(n, seed):
    """Entries only in rows/cols of a few narrow windows: long empty runs at the head, in the
    middle (several lengths around the 1024 / 32768 thresholds) and at the tail."""
    rng = np.random.default_rng(seed)
    windows = [(1000, 1040), (1072, 1073), (1106, 1200), (2223, 2300), (3324, 3330),
               (36100, 36200), (250000, 250003)]  # gaps of 31, 33, 1023, 1024, 32770, 213800
    ids = np.concatenate([np.arange(a, b) for a, b in windows])
    e = rng.choice(ids, size=(4000, 2))
    rr, cc = np.concatenate([e[:, 0], e[:, 1]]), np.concatenate([e[:, 1], e[:, 0]])
    key = np.unique(rr.astype(np.int64) * n + cc)
    return (key // n).astype(np.int32), (key % n).astype(np.int32)


@pytest.mark.parametrize("nt", [np.int32, np.int64])
def test_long_empty_runs(sb, orc, nt):
    n = 300007
    row, col = _gappy(n, 5)
    vals = graphs.vals_for(len(row))
    check_all_conversions(sb, orc, n, row, col, vals, nt)
    # a matrix with a single entry, and one whose only entries sit in the last row
    for r0 in (0, 137, n - 1):
        row1, col1 = np.array([r0], dtype=np.int32), np.array([n - 1 - r0], dtype=np.int32)
        check_all_conversions(sb, orc, n, row1, col1, np.array([2.5], dtype=np.float32), nt)


def test_row_block_csc_trailing_columns(sb, orc):
    """The sharded CSR->CSC hands each rank a row block whose column space is the whole matrix:
    most columns of the block are empty (the 30 ms pathology fixed by the gap fill)."""
    n, rp, col, vals = graphs.poisson(211, 97)
    lo, hi = 3000, 9000
    a, b = int(rp[lo]), int(rp[hi])
    rp_l = (rp[lo:hi + 1] - a).astype(np.int32)
    col_l, val_l = col[a:b], vals[a:b]
    got = sb.csr_to_csc_block(lo, hi - lo, n, dev(rp_l), dev(col_l), dev(val_l))
    # oracle: transpose the block embedded in an n x n matrix (rows outside the block empty)
    rp_full = np.zeros(n + 1, dtype=np.int32)
    rp_full[lo + 1:hi + 1] = rp_l[1:]
    rp_full[hi + 1:] = rp_l[-1]
    exp = orc.csr_to_csc(n, n, rp_full, col_l, val_l)
    for x, y, what in zip(got, exp, ("col_ptr", "row", "vals")):
        assert eq(host(x), y), f"csr_to_csc_block {what}"


@pytest.mark.parametrize("types", [(np.int32, np.int32, np.float32), (np.int64, np.int64, np.float64),
                                   (np.int32, np.int32, None)],
                         ids=["i32_i32_f32", "i64_i64_f64", "i32_i32_void"])
def test_unaligned_operands(sb, orc, types):
    idt, nt, vt = types
    n, row, col = graphs.er(6000, 6, seed=31)
    row, col = row.astype(idt), col.astype(idt)
    vals = None if vt is None else graphs.vals_for(len(row), dtype=vt)
    off = 1 if idt == np.int32 else 1  # 4 or 8 bytes past a 16-byte boundary
    rp, cc, vv = check_all_conversions(sb, orc, n, row, col, vals, nt,
                                       wrap=lambda a: unaligned(a, off))
    # unsorted COO through the constructor sort, unaligned
    rng = np.random.default_rng(3)
    p = rng.permutation(len(row))
    ur, uc = unaligned(row[p]), unaligned(col[p])
    uv = None if vals is None else unaligned(vals[p])
    sb.coo_sort_(n, n, ur, uc, uv)
    er, ec, ev = orc.coo_ctor_sort(n, n, row[p], col[p], None if vals is None else vals[p], nt)
    assert eq(host(ur), er) and eq(host(uc), ec) and eq(host(uv), ev)
    # degree features and DegreeReorder on an unaligned row_ptr
    urp = unaligned(rp)
    z = np.zeros(len(cc), dtype=vt or np.float32)
    assert eq(host(sb.degrees(n, urp, TT[idt])), orc.degrees(n, rp, cc, z))
    ft = torch.float64 if vt == np.float64 else torch.float32
    assert eq(host(sb.degree_distribution(n, len(cc), urp, ft)),
              orc.degree_distribution(n, rp, cc, z if vt is not None else None))
    assert eq(host(sb.degree_reorder(n, urp, True, TT[idt])),
              orc.degree_reorder(n, rp, cc, True, None if vt is None else z))


@pytest.mark.parametrize("nnz_target", [1, 2, 31, 4095, 4096, 4097, 8191, 8193, 12289])
def test_radix_tile_edges(sb, orc, nnz_target):
    """Exactly nnz_target nonzeros: the last radix tile is empty, full, or holds 1 record."""
    rng = np.random.default_rng(nnz_target)
    n = 700
    key = rng.choice(n * n, size=nnz_target, replace=False).astype(np.int64)
    key.sort()
    row, col = (key // n).astype(np.int32), (key % n).astype(np.int32)
    vals = graphs.vals_for(nnz_target)
    check_all_conversions(sb, orc, n, row, col, vals)
    p = rng.permutation(nnz_target)
    ur, uc, uv = dev(row[p]), dev(col[p]), dev(vals[p])
    sb.coo_sort_(n, n, ur, uc, uv)
    er, ec, ev = orc.coo_ctor_sort(n, n, row[p], col[p], vals[p])
    assert eq(host(ur), er) and eq(host(uc), ec) and eq(host(uv), ev)


def test_short_rows_duplicate_columns(sb, orc):
    """Rows of <= 8 entries with repeated column ids and distinct values: the CSR constructor
    orders pairs (col, val) (csr.cc:147-148); exercised through Permute2D's short-row kernel."""
    rng = np.random.default_rng(123)
    n = 4099
    deg = rng.integers(0, 9, size=n)
    row = np.repeat(np.arange(n), deg).astype(np.int32)
    col = np.concatenate([np.sort(rng.integers(0, 12, size=d) + rng.integers(0, n - 12))
                          for d in deg]).astype(np.int32)
    vals = rng.permutation(len(col)).astype(np.float32)  # all distinct
    rp = graphs.csr_of(n, row, col)
    cc, vv = orc.csr_ctor_sort(n, n, rp, col, vals)
    order = rng.permutation(n).astype(np.int32)
    exp = orc.permute2d(n, n, rp, cc, vv, order, order)
    got = sb.permute2d(n, n, dev(rp), dev(cc), dev(vv), dev(order), dev(order))
    for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
        assert eq(host(a), b), f"permute2d with duplicate columns: {what}"


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 127, 128, 129, 1023, 1024, 1025, 4099])
def test_degree_features_vector_edges(sb, orc, n):
    rng = np.random.default_rng(n)
    deg = rng.integers(0, 6, size=n)
    row = np.repeat(np.arange(n), deg).astype(np.int32)
    col = rng.integers(0, n, size=len(row)).astype(np.int32)
    for nt in (np.int32, np.int64):
        rp = graphs.csr_of(n, row, col, nt)
        z = np.zeros(len(col), dtype=np.float32)
        assert eq(host(sb.degrees(n, dev(rp), torch.int32)), orc.degrees(n, rp, col, z))
        if len(col):
            assert eq(host(sb.degree_distribution(n, len(col), dev(rp), torch.float32)),
                      orc.degree_distribution(n, rp, col, z))


@pytest.mark.parametrize("types", [(np.int32, np.int32, np.float32), (np.int64, np.int64, np.float64),
                                   (np.int32, np.int32, None)],
                         ids=["i32_i32_f32", "i64_i64_f64", "i32_i32_void"])
def test_long_rows_segmented_sort(sb, orc, types):
    """Rows longer than the on-chip limit go through the segmented radix sort; rows longer than
    its 8192-record tile span several chunks (lengths around the chunk size on purpose)."""
    idt, nt, vt = types
    rng = np.random.default_rng(41)
    n = 30011
    # (3072 | 9216, 1536 | 4608, 5632 | 16896 = what one CTA of the two on-chip shapes sorts in
    # shared memory for the three type sets)
    lens = {0: 20000, 5: 9000, 6: 8192, 7: 8193, 11: 1025, 12: 1024, 29999: 16385, 40: 3072,
            41: 3073, 42: 9216, 43: 9217, 44: 1536, 45: 1537, 46: 4608, 47: 4609, 48: 5632,
            49: 5633, 50: 16896, 51: 16897, 52: 2047, 53: 3001}
    rows = [np.full(c, r) for r, c in lens.items()] + [rng.integers(13, n - 20, 50000)]
    cols = [rng.choice(n, c, replace=False) for c in lens.values()] + [rng.integers(0, n, 50000)]
    key = np.unique(np.concatenate(rows).astype(np.int64) * n + np.concatenate(cols))
    row, col = (key // n).astype(idt), (key % n).astype(idt)
    rp = graphs.csr_of(n, row, col, nt)
    vals = None if vt is None else graphs.vals_for(len(col), dtype=vt)
    order = rng.permutation(n).astype(idt)
    exp = orc.permute2d(n, n, rp, col, vals, order, order)
    got = sb.permute2d(n, n, dev(rp), dev(col), dev(vals), dev(order), dev(order))
    for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
        assert eq(host(a), b), f"permute2d long rows {what}"
    # CSR constructor sort of the same rows after shuffling every row
    sh = col.copy()
    for r in lens:
        a, b = int(rp[r]), int(rp[r + 1])
        sh[a:b] = rng.permutation(sh[a:b])
    ecol, evals = orc.csr_ctor_sort(n, n, rp, sh, vals)
    dcol, dvals = dev(sh), dev(vals)
    sb.compressed_sort_(n, n, dev(rp), dcol, dvals)
    assert eq(host(dcol), ecol) and eq(host(dvals), evals)


@pytest.mark.parametrize("maxdeg", [9, 32, 33, 64])
@pytest.mark.parametrize("types", [(np.int32, np.int32, np.float32), (np.int32, np.int64, np.float32),
                                   (np.int64, np.int64, np.float64), (np.int32, np.int32, None)],
                         ids=["i32_i32_f32", "i32_i64_f32", "i64_i64_f64", "i32_i32_void"])
def test_permute2d_mid_rows_kernel(sb, orc, types, maxdeg):
    """Longest row in 9..64: the barrier-free warp-batch kernel (16 rows per warp up to 32
    entries, 8 rows per warp up to 64), with empty rows, a ragged last batch, rows of exactly
    the maximum length, and row / column orders that differ."""
    idt, nt, vt = types
    rng = np.random.default_rng(1000 + maxdeg)
    n = 20011
    deg = rng.integers(0, maxdeg + 1, size=n)
    deg[rng.integers(0, n, size=300)] = 0
    deg[[0, 17, n - 1]] = maxdeg
    row = np.repeat(np.arange(n), deg)
    col = np.concatenate([rng.choice(n, size=d, replace=False) for d in deg]).astype(np.int64)
    rp = graphs.csr_of(n, row.astype(idt), col.astype(idt), nt)
    cc = col.astype(idt)
    vv = None if vt is None else graphs.vals_for(len(cc), dtype=vt)
    cc2, vv2 = orc.csr_ctor_sort(n, n, rp, cc, vv)
    ro, co = rng.permutation(n).astype(idt), rng.permutation(n).astype(idt)
    for r_, c_ in ((ro, ro), (ro, co), (None, co), (ro, None)):
        exp = orc.permute2d(n, n, rp, cc2, vv2, r_, c_)
        got = sb.permute2d(n, n, dev(rp), dev(cc2), dev(vv2), dev(r_), dev(c_))
        for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
            assert eq(host(a), b), f"mid-row permute2d {what} {idt} {nt} {vt} max={maxdeg}"


def test_mid_rows_duplicate_columns(sb, orc):
    rng = np.random.default_rng(321)
    n = 3001
    deg = rng.integers(0, 41, size=n)
    row = np.repeat(np.arange(n), deg).astype(np.int32)
    col = np.concatenate([np.sort(rng.integers(0, 25, size=d) + rng.integers(0, n - 25))
                          for d in deg]).astype(np.int32)
    vals = rng.permutation(len(col)).astype(np.float32)
    rp = graphs.csr_of(n, row, col)
    cc, vv = orc.csr_ctor_sort(n, n, rp, col, vals)
    order = rng.permutation(n).astype(np.int32)
    exp = orc.permute2d(n, n, rp, cc, vv, order, order)
    got = sb.permute2d(n, n, dev(rp), dev(cc), dev(vv), dev(order), dev(order))
    for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
        assert eq(host(a), b), f"mid rows with duplicate columns: {what}"
