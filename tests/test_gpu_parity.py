"""GPU parity tests (-m gpu): every C-ABI entry point against the oracle, bit-exact.

Inputs are seeded; sizes are chosen so that the single-threaded oracle finishes in seconds.
All calls go through libsb200.so (ctypes, sparsebase_b200.lib); results are compared with
np.array_equal on the raw arrays (row_ptr / col / vals / permutations / float features).
"""
import json
import os

import numpy as np
import pytest
import torch

import graphs
import oracle_lib

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
    GOLD = json.load(f)


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


def _graph_cases():
    cases = []
    n, r, c = graphs.rmat(12, 16, seed=3)
    cases.append(("rmat12", n, r, c))
    n, r, c = graphs.er(20000, 8, seed=5)
    cases.append(("er20k", n, r, c))
    n, r, c = graphs.band(30000, 31, 0.5, seed=9, shuffle_seed=10)
    cases.append(("band30k", n, r, c))
    n, rp, col, _ = graphs.poisson(173, 97)
    cases.append(("poisson173x97", n, np.repeat(np.arange(n, dtype=np.int32), np.diff(rp)), col))
    n, r, c = graphs.multi_component()
    cases.append(("multi", n, r, c))
    # hub rows: exercises the 33..2048 bitonic tier and the >2048 global tier
    rng = np.random.default_rng(17)
    n = 9000
    rows = [np.full(7000, 0), np.full(2500, 1), np.full(1000, 2), np.full(40, 3),
            rng.integers(4, n, 30000)]
    cols = [rng.choice(n, 7000, replace=False), rng.choice(n, 2500, replace=False),
            rng.choice(n, 1000, replace=False), rng.choice(n, 40, replace=False),
            rng.integers(0, n, 30000)]
    r, c = np.concatenate(rows), np.concatenate(cols)
    rr, cc = np.concatenate([r, c]), np.concatenate([c, r])
    key = np.unique(rr.astype(np.int64) * n + cc)
    cases.append(("hubs", n, (key // n).astype(np.int32), (key % n).astype(np.int32)))
    return cases


CASES = _graph_cases()
IDS = [c[0] for c in CASES]
TYPES = [(np.int32, np.int32, np.float32), (np.int32, np.int64, np.float32),
         (np.int64, np.int64, np.float64), (np.int32, np.int32, None)]
TYPE_IDS = ["i32_i32_f32", "i32_i64_f32", "i64_i64_f64", "i32_i32_void"]
TT = {np.int32: torch.int32, np.int64: torch.int64, np.float32: torch.float32,
      np.float64: torch.float64}


# ------------------------------------------------------------------ golden vectors
def test_golden_conversions(sb):
    g = GOLD["converter_common"]
    n, m = g["n"], g["m"]
    i32 = lambda x: dev(np.asarray(x, dtype=np.int32))  # noqa: E731
    rp, col, vals = sb.coo_to_csr(n, m, i32(g["coo_row"]), i32(g["coo_col"]), i32(g["coo_vals"]))
    assert host(rp).tolist() == g["csr_row_ptr"] and host(col).tolist() == g["csr_col"]
    assert host(vals).tolist() == g["csr_vals"]
    cp, row, vals = sb.coo_to_csc(n, m, i32(g["coo_row"]), i32(g["coo_col"]), i32(g["coo_vals"]))
    assert host(cp).tolist() == g["csc_col_ptr"] and host(row).tolist() == g["csc_row"]
    assert host(vals).tolist() == g["csc_vals"]
    cp, row, vals = sb.csr_to_csc(n, m, i32(g["csr_row_ptr"]), i32(g["csr_col"]),
                                  i32(g["csr_vals"]))
    assert host(cp).tolist() == g["csc_col_ptr"] and host(row).tolist() == g["csc_row"]
    assert host(vals).tolist() == g["csc_vals"]
    row, col, vals = sb.csr_to_coo(n, m, i32(g["csr_row_ptr"]), i32(g["csr_col"]),
                                   i32(g["csr_vals"]))
    assert host(row).tolist() == g["coo_row"] and host(col).tolist() == g["coo_col"]
    assert host(vals).tolist() == g["coo_vals"]


def test_golden_ctor_sorts(sb):
    g = GOLD["format_common"]
    i32 = lambda x: dev(np.asarray(x, dtype=np.int32))  # noqa: E731
    col, vals = i32(g["csr_col_shuffled"]), i32(g["csr_vals_shuffled"])
    assert sb.compressed_sort_(4, 4, i32(g["csr_row_ptr"]), col, vals) is False
    assert host(col).tolist() == g["csr_col"] and host(vals).tolist() == g["csr_vals"]
    col = i32(g["csr_col_shuffled"])
    sb.compressed_sort_(4, 4, i32(g["csr_row_ptr"]), col, None)
    assert host(col).tolist() == g["csr_col"]
    row, col, vals = i32(g["coo_row_shuffled"]), i32(g["coo_col_shuffled"]), i32(
        g["coo_vals_shuffled"])
    assert sb.coo_sort_(4, 4, row, col, vals) is False
    assert host(row).tolist() == g["coo_row"] and host(col).tolist() == g["coo_col"]
    assert host(vals).tolist() == g["coo_vals"]
    row, col, vals = i32(g["coo_row"]), i32(g["coo_col"]), i32(g["coo_vals"])
    assert sb.coo_sort_(4, 4, row, col, vals) is True


def test_golden_permute_and_features(sb):
    g = GOLD["functionality_common"]
    n = g["n"]
    i32 = lambda x: dev(np.asarray(x, dtype=np.int32))  # noqa: E731
    rp, cols, vals = i32(g["row_ptr"]), i32(g["cols"]), i32(g["vals"])
    r, c = i32(g["r_reorder_vector"]), i32(g["c_reorder_vector"])
    out = sb.permute2d(n, n, rp, cols, vals, r, None)
    assert [host(x).tolist() for x in out] == [g["r_row_ptr"], g["r_cols"], g["r_vals"]]
    out = sb.permute2d(n, n, rp, cols, vals, None, c)
    assert [host(x).tolist() for x in out] == [g["c_row_ptr"], g["c_cols"], g["c_vals"]]
    out = sb.permute2d(n, n, rp, cols, vals, r, c)
    assert [host(x).tolist() for x in out] == [g["rc_row_ptr"], g["rc_cols"], g["rc_vals"]]
    out = sb.permute2d(n, n, i32(g["rc_row_ptr"]), i32(g["rc_cols"]), i32(g["rc_vals"]),
                       sb.inverse_permutation(r), sb.inverse_permutation(c))
    assert [host(x).tolist() for x in out] == [g["row_ptr"], g["cols"], g["vals"]]
    assert host(sb.inverse_permutation(i32(g["perm_array"]))).tolist() == g["inverse_perm_array"]
    arr = dev(np.asarray(g["original_array"], dtype=np.float32))
    got = host(sb.permute1d(arr, i32(g["inverse_perm_array"])))
    assert got.tolist() == np.asarray(g["reordered_array"], dtype=np.float32).tolist()
    assert host(sb.degrees(n, rp)).tolist() == g["degrees"]
    assert host(sb.degree_distribution(n, g["nnz"], rp)).tolist() == g["distribution"]


# ------------------------------------------------------------------ conversions
@pytest.mark.parametrize("types", TYPES, ids=TYPE_IDS)
@pytest.mark.parametrize("name,n,row,col", CASES, ids=IDS)
def test_conversions(sb, orc, name, n, row, col, types):
    idt, nt, vt = types
    row, col = row.astype(idt), col.astype(idt)
    nnz = len(row)
    vals = None if vt is None else graphs.vals_for(nnz, dtype=vt)
    # COO -> CSR on the sorted COO
    exp = orc.coo_to_csr(n, n, row, col, vals, nt)
    got = sb.coo_to_csr(n, n, dev(row), dev(col), dev(vals), TT[nt])
    for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
        assert eq(host(a), b), f"coo_to_csr {what}"
    rp, cc, vv = exp
    # CSR -> CSC, COO -> CSC, CSR -> COO
    exp = orc.csr_to_csc(n, n, rp, cc, vv)
    got = sb.csr_to_csc(n, n, dev(rp), dev(cc), dev(vv))
    for a, b, what in zip(got, exp, ("col_ptr", "row", "vals")):
        assert eq(host(a), b), f"csr_to_csc {what}"
    got = sb.coo_to_csc(n, n, dev(row), dev(col), dev(vals), TT[nt])
    for a, b, what in zip(got, exp, ("col_ptr", "row", "vals")):
        assert eq(host(a), b), f"coo_to_csc {what}"
    exp = orc.csr_to_coo(n, n, rp, cc, vv)
    got = sb.csr_to_coo(n, n, dev(rp), dev(cc), dev(vv))
    for a, b, what in zip(got, exp, ("row", "col", "vals")):
        assert eq(host(a), b), f"csr_to_coo {what}"


@pytest.mark.parametrize("types", TYPES, ids=TYPE_IDS)
@pytest.mark.parametrize("name,n,row,col", CASES[:3] + CASES[5:], ids=IDS[:3] + IDS[5:])
def test_unsorted_coo(sb, orc, name, n, row, col, types):
    """COO constructor sort (coo.cc:110-157) and COO->CSR from a shuffled edge list."""
    idt, nt, vt = types
    rng = np.random.default_rng(23)
    shuf = rng.permutation(len(row))
    r, c = row.astype(idt)[shuf], col.astype(idt)[shuf]
    v = None if vt is None else graphs.vals_for(len(row), dtype=vt)[shuf]
    exp = orc.coo_ctor_sort(n, n, r, c, v, nt)
    dr, dc, dv = dev(r), dev(c), dev(v)
    assert sb.coo_sort_(n, n, dr, dc, dv) is False
    for a, b, what in zip((dr, dc, dv), exp, ("row", "col", "vals")):
        assert eq(host(a), b), f"coo_sort {what}"
    # ignore_sort=true COO handed to COO->CSR: histogram path + CSR-constructor sort
    exp = orc.coo_to_csr(n, n, r, c, v, nt)
    # the oracle's coo_to_csr runs the COO ctor first; the library's contract is the same
    # once the caller has constructed the COO (coo_sort_), which is what dr/dc/dv now hold
    got = sb.coo_to_csr(n, n, dr, dc, dv, TT[nt])
    for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
        assert eq(host(a), b), f"coo_to_csr(after sort) {what}"


def test_coo_to_csr_ignore_sort_semantics(sb):
    """COO built with ignore_sort=true and unsorted rows: the reference histograms the rows,
    copies col/vals verbatim (converter_order_two.cc:180-207) and the CSR ctor then sorts each
    row segment by (col, val) (csr.cc:99-157)."""
    n = 6
    row = np.array([3, 0, 3, 1, 0, 5, 3], dtype=np.int32)
    col = np.array([2, 4, 0, 1, 1, 5, 4], dtype=np.int32)
    vals = np.arange(7, dtype=np.float32) + 1
    rp = np.zeros(n + 1, dtype=np.int32)
    np.add.at(rp, row + 1, 1)
    rp = np.cumsum(rp).astype(np.int32)
    c2, v2 = col.copy(), vals.copy()
    for i in range(n):  # reference semantics, restated inline for this 7-entry case
        s, e = rp[i], rp[i + 1]
        o = np.lexsort((v2[s:e], c2[s:e]))
        c2[s:e], v2[s:e] = c2[s:e][o], v2[s:e][o]
    got = sb.coo_to_csr(n, n, dev(row), dev(col), dev(vals))
    assert eq(host(got[0]), rp) and eq(host(got[1]), c2) and eq(host(got[2]), v2)


@pytest.mark.parametrize("types", TYPES, ids=TYPE_IDS)
def test_compressed_sort(sb, orc, types):
    """CSR constructor: rows shuffled internally must come back sorted (csr.cc:123-157)."""
    idt, nt, vt = types
    name, n, row, col = CASES[5]
    rp = graphs.csr_of(n, row, col, nt)
    rng = np.random.default_rng(29)
    c = col.astype(idt).copy()
    v = None if vt is None else graphs.vals_for(len(row), dtype=vt)
    for i in range(n):
        s, e = rp[i], rp[i + 1]
        p = rng.permutation(e - s)
        c[s:e] = c[s:e][p]
        if v is not None:
            v[s:e] = v[s:e][p]
    exp = orc.csr_ctor_sort(n, n, rp, c, v)
    dc, dv = dev(c), dev(v)
    assert sb.compressed_sort_(n, n, dev(rp), dc, dv) is False
    assert eq(host(dc), exp[0]) and eq(host(dv), exp[1])
    assert sb.compressed_sort_(n, n, dev(rp), dc, dv) is True


def test_rectangular(sb, orc):
    n, m = 5000, 1200
    row, col = graphs.random_rect(n, m, 60000, seed=31)
    vals = graphs.vals_for(len(row))
    exp = orc.coo_to_csr(n, m, row, col, vals)
    got = sb.coo_to_csr(n, m, dev(row), dev(col), dev(vals))
    for a, b in zip(got, exp):
        assert eq(host(a), b)
    rp, cc, vv = exp
    exp = orc.csr_to_csc(n, m, rp, cc, vv)
    got = sb.csr_to_csc(n, m, dev(rp), dev(cc), dev(vv))
    for a, b in zip(got, exp):
        assert eq(host(a), b)
    rng = np.random.default_rng(37)
    ro, co = rng.permutation(n).astype(np.int32), rng.permutation(m).astype(np.int32)
    exp = orc.permute2d(n, m, rp, cc, vv, ro, co)
    got = sb.permute2d(n, m, dev(rp), dev(cc), dev(vv), dev(ro), dev(co))
    for a, b in zip(got, exp):
        assert eq(host(a), b)


def test_empty_and_tiny(sb, orc):
    e32 = np.zeros(0, dtype=np.int32)
    ef = np.zeros(0, dtype=np.float32)
    for n in (0, 1, 5):
        got = sb.coo_to_csr(n, n, dev(e32), dev(e32), dev(ef))
        assert host(got[0]).tolist() == [0] * (n + 1) and got[1].numel() == 0
        rp = np.zeros(n + 1, dtype=np.int32)
        got = sb.csr_to_csc(n, n, dev(rp), dev(e32), dev(ef))
        assert host(got[0]).tolist() == [0] * (n + 1)
        if n:
            order = np.arange(n, dtype=np.int32)[::-1].copy()
            got = sb.permute2d(n, n, dev(rp), dev(e32), dev(ef), dev(order), dev(order))
            assert host(got[0]).tolist() == [0] * (n + 1)
            assert eq(host(sb.degree_reorder(n, dev(rp))), orc.degree_reorder(n, rp, e32))
    # single entry
    row, col, vals = np.array([2], np.int32), np.array([1], np.int32), np.array([7.5], np.float32)
    exp = orc.coo_to_csc(4, 4, row, col, vals)
    got = sb.coo_to_csc(4, 4, dev(row), dev(col), dev(vals))
    for a, b in zip(got, exp):
        assert eq(host(a), b)


# ------------------------------------------------------------------ reorderings + permutation
@pytest.mark.parametrize("types", TYPES[:3], ids=TYPE_IDS[:3])
@pytest.mark.parametrize("name,n,row,col", CASES, ids=IDS)
def test_degree_reorder_and_features(sb, orc, name, n, row, col, types):
    idt, nt, vt = types
    rp = graphs.csr_of(n, row, col, nt)
    cc = col.astype(idt)
    for asc in (True, False):
        exp = orc.degree_reorder(n, rp, cc, asc, None if vt is None else
                                 np.zeros(len(cc), dtype=vt))
        got = sb.degree_reorder(n, dev(rp), asc, TT[idt])
        assert eq(host(got), exp), f"degree_reorder asc={asc}"
    v = np.zeros(len(cc), dtype=vt)
    assert eq(host(sb.degrees(n, dev(rp), TT[idt])), orc.degrees(n, rp, cc, v))
    ft = torch.float64 if vt == np.float64 else torch.float32
    assert eq(host(sb.degree_distribution(n, len(cc), dev(rp), ft)),
              orc.degree_distribution(n, rp, cc, v))


@pytest.mark.parametrize("types", TYPES, ids=TYPE_IDS)
@pytest.mark.parametrize("name,n,row,col", CASES, ids=IDS)
def test_permute2d(sb, orc, name, n, row, col, types):
    idt, nt, vt = types
    rp = graphs.csr_of(n, row, col, nt)
    cc = col.astype(idt)
    vv = None if vt is None else graphs.vals_for(len(cc), dtype=vt)
    rng = np.random.default_rng(41)
    orders = [rng.permutation(n).astype(idt),
              orc.degree_reorder(n, rp, cc, True, vv)]
    for order in orders:
        exp = orc.permute2d(n, n, rp, cc, vv, order, order)
        d_order = dev(order)
        got = sb.permute2d(n, n, dev(rp), dev(cc), dev(vv), d_order, d_order)
        for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
            assert eq(host(a), b), f"permute2d {what}"
    # row-only / col-only
    exp = orc.permute2d(n, n, rp, cc, vv, orders[0], None)
    got = sb.permute2d(n, n, dev(rp), dev(cc), dev(vv), dev(orders[0]), None)
    for a, b in zip(got, exp):
        assert eq(host(a), b)
    exp = orc.permute2d(n, n, rp, cc, vv, None, orders[0])
    got = sb.permute2d(n, n, dev(rp), dev(cc), dev(vv), None, dev(orders[0]))
    for a, b in zip(got, exp):
        assert eq(host(a), b)


def test_permute1d_inverse(sb, orc):
    rng = np.random.default_rng(43)
    for idt, vt in ((np.int32, np.float32), (np.int64, np.float64)):
        n = 100003
        order = rng.permutation(n).astype(idt)
        vals = rng.standard_normal(n).astype(vt)
        assert eq(host(sb.permute1d(dev(vals), dev(order))), orc.permute1d(vals, order))
        assert eq(host(sb.inverse_permutation(dev(order))), orc.inverse_permutation(order))


@pytest.mark.parametrize("name,n,row,col", CASES, ids=IDS)
def test_rcm(sb, orc, name, n, row, col):
    rp = graphs.csr_of(n, row, col)
    exp = orc.rcm_reorder(n, rp, col)
    got = host(sb.rcm_reorder(n, dev(rp), dev(col)))
    assert eq(got, exp), f"rcm mismatch at {np.flatnonzero(got != exp)[:10]}"


def test_rcm_other_types(sb, orc):
    name, n, row, col = CASES[0]
    for idt, nt in ((np.int32, np.int64), (np.int64, np.int64)):
        rp = graphs.csr_of(n, row, col, nt)
        cc = col.astype(idt)
        v = np.zeros(len(cc), dtype=np.float32 if idt == np.int32 else np.float64)
        exp = orc.rcm_reorder(n, rp, cc, v)
        assert eq(host(sb.rcm_reorder(n, dev(rp), dev(cc))), exp)


def test_partition_rows(sb):
    name, n, row, col = CASES[0]
    rp = graphs.csr_of(n, row, col)
    nnz = len(row)
    for parts in (1, 2, 4, 8):
        b = sb.partition_rows(n, nnz, dev(rp), parts)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        for k in range(1, parts):
            tgt = nnz * k // parts
            assert b[k] == int(np.searchsorted(rp[:n], tgt, side="left"))


def test_no_cpu_fallback_error_path(sb):
    """Bad dtype combinations are rejected with a message, not silently mis-computed."""
    import ctypes
    lib = sb.load()
    rc = lib.sb200_degrees(0, ctypes.c_int64(4), ctypes.c_void_p(8), ctypes.c_void_p(8),
                           sb.F32, sb.I32, None)
    assert rc == 2 and b"id_type" in lib.sb200_last_error()


@pytest.mark.parametrize("shape", [(1500, 1111), (64, 20000), (3000, 7)])
def test_rcm_wide_grids_cluster_regime(sb, orc, shape):
    """Frontiers of hundreds to thousands of vertices: the cluster kernel keeps every CTA's
    share resident across levels and re-splits when the shares drift (rcm.cu cl_reload)."""
    n, rp, col, _ = graphs.poisson(*shape)
    exp = orc.rcm_reorder(n, rp, col)
    got = host(sb.rcm_reorder(n, dev(rp), dev(col)))
    assert eq(got, exp), f"rcm mismatch at {np.flatnonzero(got != exp)[:10]}"
    st = sb.rcm_last_stats()
    assert st["levels_narrow"] > 0 and st["levels_wide"] == 0


def test_permute2d_short_rows_kernel(sb, orc):
    """Matrices whose longest row has <= 8 entries take the warp-batch kernel
    (reorder.cu permute_short_rows_kernel); empty rows, every type combination."""
    rng = np.random.default_rng(77)
    n = 50021
    deg = rng.integers(0, 9, size=n)
    deg[rng.integers(0, n, size=500)] = 0
    row = np.repeat(np.arange(n), deg)
    col = np.concatenate([rng.choice(n, size=d, replace=False) for d in deg]).astype(np.int64)
    for idt, nt, vt in TYPES:
        rp = graphs.csr_of(n, row.astype(idt), col.astype(idt), nt)
        cc = col.astype(idt)
        vv = None if vt is None else graphs.vals_for(len(cc), dtype=vt)
        # CSR rows must be sorted for the oracle's input contract
        cc2, vv2 = orc.csr_ctor_sort(n, n, rp, cc, vv)
        order = rng.permutation(n).astype(idt)
        exp = orc.permute2d(n, n, rp, cc2, vv2, order, order)
        got = sb.permute2d(n, n, dev(rp), dev(cc2), dev(vv2), dev(order), dev(order))
        for a, b, what in zip(got, exp, ("row_ptr", "col", "vals")):
            assert eq(host(a), b), f"short-row permute2d {what} {idt} {nt} {vt}"


def _sparse_forest(n, seed, ppv=0.6):
    rng = np.random.default_rng(seed)
    e = rng.integers(0, n, size=(int(n * ppv), 2))
    rr, cc = np.concatenate([e[:, 0], e[:, 1]]), np.concatenate([e[:, 1], e[:, 0]])
    key = np.unique(rr.astype(np.int64) * n + cc)
    return n, (key // n).astype(np.int32), (key % n).astype(np.int32)


def test_rcm_speculative_peripheral(sb, orc, monkeypatch):
    """peripheral()'s BFS traversals after the first run as the Cuthill-McKee traversal itself
    (rcm.cu PH_PBFS_END / PH_CM_END): confirmed, continued from a unique new root, or replayed
    literally -- always the reference's permutation, with and without the speculation."""
    seen = {"spec_confirmed": 0, "spec_continued": 0, "spec_replayed": 0}
    graphs_ = [_sparse_forest(3000, 11), _sparse_forest(20000, 12, 0.8), graphs.er(2000, 2, seed=5),
               graphs.multi_component(), graphs.rmat(10, 4, seed=3)]
    for n, row, col in graphs_:
        rp = graphs.csr_of(n, row, col)
        exp = orc.rcm_reorder(n, rp, col)
        for force_wide in ("0", "1"):
            monkeypatch.setenv("SB200_RCM_FORCE_WIDE", force_wide)
            monkeypatch.setenv("SB200_RCM_NO_SPEC", "0")
            got = host(sb.rcm_reorder(n, dev(rp), dev(col)))
            assert eq(got, exp), f"rcm (speculative) mismatch at {np.flatnonzero(got != exp)[:10]}"
            st = sb.rcm_last_stats()
            for k in seen:
                seen[k] += st[k]
            monkeypatch.setenv("SB200_RCM_NO_SPEC", "1")
            got = host(sb.rcm_reorder(n, dev(rp), dev(col)))
            assert eq(got, exp)
            st = sb.rcm_last_stats()
            assert st["spec_confirmed"] == st["spec_continued"] == st["spec_replayed"] == 0
    assert all(v > 0 for v in seen.values()), seen


@pytest.mark.parametrize("cluster", ["1", "2", "4", "8", "16"])
def test_rcm_pinned_cluster_sizes(sb, orc, monkeypatch, cluster):
    """Every cluster size of the narrow regime (single CTA with block barriers up to 16 CTAs with
    one cluster barrier per level) yields the reference's permutation: grids (frontier grows and
    shrinks), a shuffled band (constant narrow frontier, sibling groups of up to 31), several
    components."""
    monkeypatch.setenv("SB200_RCM_CLUSTER", cluster)
    cases = []
    n, rp, col, _ = graphs.poisson(301, 157)
    cases.append((n, rp, col))
    n, row, col = graphs.band(40000, 31, 0.5)
    cases.append((n, graphs.csr_of(n, row, col), col))
    n, row, col = graphs.multi_component()
    cases.append((n, graphs.csr_of(n, row, col), col))
    for n, rp, col in cases:
        exp = orc.rcm_reorder(n, rp, col)
        got = host(sb.rcm_reorder(n, dev(rp), dev(col)))
        assert eq(got, exp), f"rcm mismatch at {np.flatnonzero(got != exp)[:10]}"
        assert sb.rcm_last_stats()["resizes"] == 0


def test_rcm_adaptive_cluster_and_resplit(sb, orc):
    """Adaptive sizing: a wide grid makes the kernel ask for larger clusters as the frontier
    grows and smaller ones as it shrinks, and the shares are re-split when they drift."""
    n, rp, col, _ = graphs.poisson(2500, 900)
    exp = orc.rcm_reorder(n, rp, col)
    got = host(sb.rcm_reorder(n, dev(rp), dev(col)))
    assert eq(got, exp), f"rcm mismatch at {np.flatnonzero(got != exp)[:10]}"
    st = sb.rcm_last_stats()
    assert st["levels_wide"] == 0 and st["resizes"] > 0 and st["resplits"] > 0, st


def test_rcm_share_overflow_takes_claims_back(sb, orc, monkeypatch):
    """A hub inside a narrow graph: the share holding it does not fit, the other CTAs take their
    claims of that level back, the frontier is re-split and -- when even that does not fit --
    the level is done by the wide regime; then the cluster resumes (rcm.cu cl_level abort)."""
    nx, ny = 400, 60
    n, rp, col, _ = graphs.poisson(nx, ny)
    row = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    rng = np.random.default_rng(5)
    hubs = rng.choice(n, size=3, replace=False)
    extra_r, extra_c = [], []
    for h, deg in zip(hubs, (5000, 900, 70)):   # beyond one CTA, beyond a re-split, just wide rows
        nb = rng.choice(n, size=deg, replace=False)
        extra_r += [np.full(deg, h), nb]
        extra_c += [nb, np.full(deg, h)]
    rr = np.concatenate([row] + extra_r)
    cc = np.concatenate([col.astype(np.int64)] + extra_c)
    key = np.unique(rr * n + cc)
    row2, col2 = (key // n).astype(np.int32), (key % n).astype(np.int32)
    rp2 = graphs.csr_of(n, row2, col2)
    exp = orc.rcm_reorder(n, rp2, col2)
    for cluster in (None, "2", "16"):
        if cluster is None:
            monkeypatch.delenv("SB200_RCM_CLUSTER", raising=False)
        else:
            monkeypatch.setenv("SB200_RCM_CLUSTER", cluster)
        got = host(sb.rcm_reorder(n, dev(rp2), dev(col2)))
        assert eq(got, exp), f"rcm mismatch (cluster={cluster}) at {np.flatnonzero(got != exp)[:10]}"
        st = sb.rcm_last_stats()
        assert st["levels_narrow"] > 0 and st["levels_wide"] > 0, st
