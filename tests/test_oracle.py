"""CPU tests (-m "not gpu"): pin the parity oracle.

1. the plain-C restatement (oracle/liboracle.so) against the golden vectors of the
   reference's own tests (tests/golden/reference_vectors.json);
2. the restatement against the compiled, unmodified reference (oracle/_ref/libsbref.so)
   on seeded random graphs -- skipped where oracle/_ref was not built;
3. the restatement against the committed fixtures generated from the reference
   (tests/golden/*.npz, made by tests/golden/make_golden.py).
"""
import glob
import json
import os

import numpy as np
import pytest

import graphs
import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
    GOLD = json.load(f)


def impls():
    out = [pytest.param(oracle_lib.restated(), id="restated")]
    if oracle_lib.reference() is not None:
        out.append(pytest.param(oracle_lib.reference(), id="reference"))
    return out


def i32(x):
    return np.asarray(x, dtype=np.int32)


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("o", impls())
def test_golden_conversions(o):
    g = GOLD["converter_common"]
    n, m = g["n"], g["m"]
    rp, col, vals = o.coo_to_csr(n, m, i32(g["coo_row"]), i32(g["coo_col"]), i32(g["coo_vals"]))
    assert rp.tolist() == g["csr_row_ptr"] and col.tolist() == g["csr_col"]
    assert vals.tolist() == g["csr_vals"]
    cp, row, vals = o.coo_to_csc(n, m, i32(g["coo_row"]), i32(g["coo_col"]), i32(g["coo_vals"]))
    assert cp.tolist() == g["csc_col_ptr"] and row.tolist() == g["csc_row"]
    assert vals.tolist() == g["csc_vals"]
    cp, row, vals = o.csr_to_csc(n, m, i32(g["csr_row_ptr"]), i32(g["csr_col"]), i32(g["csr_vals"]))
    assert cp.tolist() == g["csc_col_ptr"] and row.tolist() == g["csc_row"]
    assert vals.tolist() == g["csc_vals"]
    row, col, vals = o.csr_to_coo(n, m, i32(g["csr_row_ptr"]), i32(g["csr_col"]), i32(g["csr_vals"]))
    assert row.tolist() == g["coo_row"] and col.tolist() == g["coo_col"]
    assert vals.tolist() == g["coo_vals"]


@pytest.mark.parametrize("o", impls())
def test_golden_ctor_sorts(o):
    g = GOLD["format_common"]
    col, vals = o.csr_ctor_sort(4, 4, i32(g["csr_row_ptr"]), i32(g["csr_col_shuffled"]),
                                i32(g["csr_vals_shuffled"]))
    assert col.tolist() == g["csr_col"] and vals.tolist() == g["csr_vals"]
    col, vals = o.csr_ctor_sort(4, 4, i32(g["csr_row_ptr"]), i32(g["csr_col_shuffled"]), None)
    assert col.tolist() == g["csr_col"] and vals is None
    row, col, vals = o.coo_ctor_sort(4, 4, i32(g["coo_row_shuffled"]), i32(g["coo_col_shuffled"]),
                                     i32(g["coo_vals_shuffled"]))
    assert row.tolist() == g["coo_row"] and col.tolist() == g["coo_col"]
    assert vals.tolist() == g["coo_vals"]
    row, col, vals = o.coo_ctor_sort(4, 4, i32(g["coo_row_shuffled"]), i32(g["coo_col_shuffled"]),
                                     None)
    assert row.tolist() == g["coo_row"] and col.tolist() == g["coo_col"]


@pytest.mark.parametrize("o", impls())
def test_golden_permute_and_features(o):
    g = GOLD["functionality_common"]
    n = g["n"]
    rp, cols, vals = i32(g["row_ptr"]), i32(g["cols"]), i32(g["vals"])
    r, c = i32(g["r_reorder_vector"]), i32(g["c_reorder_vector"])
    out = o.permute2d(n, n, rp, cols, vals, r, None)
    assert [x.tolist() for x in out] == [g["r_row_ptr"], g["r_cols"], g["r_vals"]]
    out = o.permute2d(n, n, rp, cols, vals, None, c)
    assert [x.tolist() for x in out] == [g["c_row_ptr"], g["c_cols"], g["c_vals"]]
    out = o.permute2d(n, n, rp, cols, vals, r, c)
    assert [x.tolist() for x in out] == [g["rc_row_ptr"], g["rc_cols"], g["rc_vals"]]
    # InversePermuteTest.RowColWise (permute_order_two_tests.cc:44-62)
    out = o.permute2d(n, n, i32(g["rc_row_ptr"]), i32(g["rc_cols"]), i32(g["rc_vals"]),
                      o.inverse_permutation(r), o.inverse_permutation(c))
    assert [x.tolist() for x in out] == [g["row_ptr"], g["cols"], g["vals"]]
    assert o.inverse_permutation(i32(g["perm_array"])).tolist() == g["inverse_perm_array"]
    arr = np.asarray(g["original_array"], dtype=np.float32)
    got = o.permute1d(arr, i32(g["inverse_perm_array"]))
    assert got.tolist() == np.asarray(g["reordered_array"], dtype=np.float32).tolist()
    assert o.degrees(n, rp, cols, vals).tolist() == g["degrees"]
    assert o.degree_distribution(n, rp, cols, vals).tolist() == g["distribution"]


# ------------------------------------------------------------------ restatement == reference
def _cases():
    cases = []
    n, r, c = graphs.rmat(10, 8, seed=3)
    cases.append(("rmat10", n, r, c))
    n, r, c = graphs.er(3000, 4, seed=5)
    cases.append(("er3000", n, r, c))
    n, r, c = graphs.band(2000, 7, 0.5, seed=9, shuffle_seed=10)
    cases.append(("band2000", n, r, c))
    n, rp, col, _ = graphs.poisson(37, 23)
    row = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp))
    cases.append(("poisson37x23", n, row, col))
    n, r, c = graphs.multi_component()
    cases.append(("multi", n, r, c))
    return cases


CASES = _cases()
needs_ref = pytest.mark.skipif(oracle_lib.reference() is None,
                               reason="oracle/_ref/libsbref.so not built (reference tree absent)")


@needs_ref
@pytest.mark.parametrize("name,n,row,col", CASES, ids=[c[0] for c in CASES])
def test_restated_equals_reference(name, n, row, col):
    a, b = oracle_lib.restated(), oracle_lib.reference()
    nnz = len(row)
    vals = graphs.vals_for(nnz)
    rng = np.random.default_rng(11)
    shuf = rng.permutation(nnz)
    for x, y in zip(a.coo_ctor_sort(n, n, row[shuf], col[shuf], vals[shuf]),
                    b.coo_ctor_sort(n, n, row[shuf], col[shuf], vals[shuf])):
        assert np.array_equal(x, y)
    for x, y in zip(a.coo_to_csr(n, n, row[shuf], col[shuf], vals[shuf]),
                    b.coo_to_csr(n, n, row[shuf], col[shuf], vals[shuf])):
        assert np.array_equal(x, y)
    rp, cc, vv = a.coo_to_csr(n, n, row, col, vals)
    assert np.array_equal(rp, graphs.csr_of(n, row, col))
    for x, y in zip(a.csr_to_csc(n, n, rp, cc, vv), b.csr_to_csc(n, n, rp, cc, vv)):
        assert np.array_equal(x, y)
    for x, y in zip(a.coo_to_csc(n, n, row, col, vals), b.coo_to_csc(n, n, row, col, vals)):
        assert np.array_equal(x, y)
    for x, y in zip(a.csr_to_coo(n, n, rp, cc, vv), b.csr_to_coo(n, n, rp, cc, vv)):
        assert np.array_equal(x, y)
    for asc in (True, False):
        assert np.array_equal(a.degree_reorder(n, rp, cc, asc), b.degree_reorder(n, rp, cc, asc))
    ia, ib = a.rcm_reorder(n, rp, cc), b.rcm_reorder(n, rp, cc)
    assert np.array_equal(ia, ib)
    assert np.array_equal(np.sort(ia), np.arange(n))
    for order in (ia, a.degree_reorder(n, rp, cc, True), rng.permutation(n).astype(np.int32)):
        for x, y in zip(a.permute2d(n, n, rp, cc, vv, order, order),
                        b.permute2d(n, n, rp, cc, vv, order, order)):
            assert np.array_equal(x, y)
    assert np.array_equal(a.degrees(n, rp, cc), b.degrees(n, rp, cc))
    assert np.array_equal(a.degree_distribution(n, rp, cc), b.degree_distribution(n, rp, cc))


@needs_ref
def test_restated_equals_reference_other_types():
    a, b = oracle_lib.restated(), oracle_lib.reference()
    n, row, col = graphs.rmat(9, 8, seed=21)
    nnz = len(row)
    for idt, nt, vt in ((np.int32, np.int64, np.float32), (np.int64, np.int64, np.float64),
                        (np.int32, np.int32, None)):
        r, c = row.astype(idt), col.astype(idt)
        v = None if vt is None else graphs.vals_for(nnz, dtype=vt)
        for x, y in zip(a.coo_to_csr(n, n, r, c, v, nt), b.coo_to_csr(n, n, r, c, v, nt)):
            assert (x is None and y is None) or np.array_equal(x, y)
        rp, cc, vv = a.coo_to_csr(n, n, r, c, v, nt)
        for x, y in zip(a.csr_to_csc(n, n, rp, cc, vv), b.csr_to_csc(n, n, rp, cc, vv)):
            assert (x is None and y is None) or np.array_equal(x, y)
        ia, ib = a.rcm_reorder(n, rp, cc, vv), b.rcm_reorder(n, rp, cc, vv)
        assert np.array_equal(ia, ib)
        da, db = a.degree_reorder(n, rp, cc, True, vv), b.degree_reorder(n, rp, cc, True, vv)
        assert np.array_equal(da, db)
        for x, y in zip(a.permute2d(n, n, rp, cc, vv, ia, ia), b.permute2d(n, n, rp, cc, vv, ib, ib)):
            assert (x is None and y is None) or np.array_equal(x, y)
        assert np.array_equal(a.degree_distribution(n, rp, cc, vv),
                              b.degree_distribution(n, rp, cc, vv))


def test_rect_matrix_conversions_property():
    """12x9-style rectangular case (m < n): round trip through the restatement."""
    o = oracle_lib.restated()
    n, m = 57, 23
    row, col = graphs.random_rect(n, m, 400, seed=2)
    vals = graphs.vals_for(len(row))
    rp, cc, vv = o.coo_to_csr(n, m, row, col, vals)
    r2, c2, v2 = o.csr_to_coo(n, m, rp, cc, vv)
    assert np.array_equal(r2, row) and np.array_equal(c2, col) and np.array_equal(v2, vals)
    cp, rr, cv = o.csr_to_csc(n, m, rp, cc, vv)
    # transpose twice == identity (as a set of triples, sorted by (row, col))
    cols = np.repeat(np.arange(n, dtype=np.int32), np.diff(cp))
    order = np.lexsort((cols, rr))
    assert np.array_equal(rr[order], row) and np.array_equal(cols[order], col)
    assert np.array_equal(cv[order], vals)


# ------------------------------------------------------------------ committed fixtures
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "ref_*.npz"))),
                         ids=os.path.basename)
def test_restated_equals_committed_reference_fixtures(path):
    z = np.load(path)
    o = oracle_lib.restated()
    n = int(z["n"])
    rp, col, vals = z["row_ptr"], z["col"], z["vals"]
    assert np.array_equal(o.degree_reorder(n, rp, col, True), z["degree_asc"])
    assert np.array_equal(o.degree_reorder(n, rp, col, False), z["degree_desc"])
    assert np.array_equal(o.rcm_reorder(n, rp, col), z["rcm"])
    out = o.permute2d(n, n, rp, col, vals, z["rcm"], z["rcm"])
    assert np.array_equal(out[0], z["p2d_row_ptr"]) and np.array_equal(out[1], z["p2d_col"])
    assert np.array_equal(out[2], z["p2d_vals"])
    out = o.csr_to_csc(n, n, rp, col, vals)
    assert np.array_equal(out[0], z["csc_col_ptr"]) and np.array_equal(out[1], z["csc_row"])
    assert np.array_equal(out[2], z["csc_vals"])
    assert np.array_equal(o.degree_distribution(n, rp, col), z["degree_distribution"])
