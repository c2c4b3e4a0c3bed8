"""The callers on either side of the path:

* edge list -> sorted, de-duplicated COO   io/edge_list_reader.cc:28-151
* fused degree features + Bandwidth + Profile   feature/degrees_degree_distribution.cc:147-166,
  feature/min_max_avg_degree.cc:168-191, feature/bandwidth.cc:92-111, feature/profile.cc:92-106
* ReorderHeatmap (rank 3)   reorder/reorder_heatmap.cc:43-120
* BOBAReorder (rank 4)   reorder/boba_reorder.cc:35-137

CPU part: the restated oracle against the compiled reference (the reference's EdgeListReader
reads a text file the harness writes) and against the reference's own golden vectors
(tests/suites/sparsebase/feature/*_tests.cc via functionality_common.inc:6-15).
GPU part: sb200_edges_to_coo / sb200_degree_features against the oracle, bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

import graphs
import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
    GOLD = json.load(f)

needs_ref = pytest.mark.skipif(oracle_lib.reference() is None,
                               reason="oracle/_ref/libsbref.so not built (no /root/reference)")


def eq(a, b):
    if a is None or b is None:
        return a is None and b is None
    a, b = np.asarray(a), np.asarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(
        a.view(np.uint8), b.view(np.uint8))


def edge_lists():
    """(name, u, v, w) -- weights are a function of the unordered pair, so that duplicate edges
    carry equal weights (the survivor of the reference's unstable sort + unique is then defined)."""
    rng = np.random.default_rng(7)
    out = []
    for name, n, e in (("small", 50, 400), ("mid", 5000, 60000), ("sparse", 200000, 300000)):
        u = rng.integers(0, n, size=e).astype(np.int32)
        v = rng.integers(0, n, size=e).astype(np.int32)
        u[: e // 20] = v[: e // 20]                      # self loops
        u[e // 2: e // 2 + e // 10] = u[: e // 10]        # duplicates
        v[e // 2: e // 2 + e // 10] = v[: e // 10]
        lo, hi = np.minimum(u, v).astype(np.int64), np.maximum(u, v).astype(np.int64)
        w = (((lo * 2654435761 + hi * 40503) % 65536).astype(np.float32) - 32768.0) * 0.5
        out.append((name, u, v, w))
    # rectangular: row ids up to 300, column ids up to 17
    u = rng.integers(0, 300, size=2000).astype(np.int32)
    v = rng.integers(0, 17, size=2000).astype(np.int32)
    lo, hi = np.minimum(u, v).astype(np.int64), np.maximum(u, v).astype(np.int64)
    out.append(("rect", u, v, ((lo * 31 + hi) % 97).astype(np.float32) - 40.0))
    return out


FLAGS = [  # remove_duplicates, remove_self, undirected, square
    (True, False, False, False), (True, True, True, False), (False, True, False, True),
    (True, True, False, True), (False, False, True, False)]


def feature_graphs():
    cases = []
    n, r, c = graphs.rmat(10, 8, seed=3)
    cases.append(("rmat10", n, graphs.csr_of(n, r, c), c))
    n, r, c = graphs.band(5000, 31, 0.5, seed=9, shuffle_seed=10)
    cases.append(("band5k_shuffled", n, graphs.csr_of(n, r, c), c))
    n, r, c = graphs.band(5000, 31, 0.5, seed=9, shuffle_seed=None)
    cases.append(("band5k", n, graphs.csr_of(n, r, c), c))
    n, rp, col, _ = graphs.poisson(57, 33)
    cases.append(("poisson57x33", n, rp, col))
    n, r, c = graphs.multi_component()
    cases.append(("multi", n, graphs.csr_of(n, r, c), c))
    return cases


# ------------------------------------------------------------------ CPU: oracle pinned
@needs_ref
@pytest.mark.parametrize("flags", FLAGS, ids=[str(f) for f in FLAGS])
def test_edges_to_coo_restated_equals_reference(flags):
    a, b = oracle_lib.restated(), oracle_lib.reference()
    for name, u, v, w in edge_lists():
        if name == "sparse":
            continue  # (the reference parses a text file: keep it short)
        for ww in (w, None):
            ra = a.edges_to_coo(u, v, ww, *flags)
            rb = b.edges_to_coo(u, v, ww, *flags)
            assert ra[0] == rb[0] and ra[1] == rb[1], (name, ra[:2], rb[:2])
            for x, y, what in zip(ra[2:], rb[2:], ("row", "col", "vals")):
                assert eq(x, y), f"{name} {flags} {what}"


@needs_ref
def test_degree_features_restated_equals_reference():
    a, b = oracle_lib.restated(), oracle_lib.reference()
    for name, n, rp, col in feature_graphs():
        ra, rb = a.degree_features(n, rp, col), b.degree_features(n, rp, col)
        assert eq(ra[0], rb[0]) and eq(ra[1], rb[1]), name
        assert ra[2] == rb[2], (name, ra[2], rb[2])
        assert np.float32(ra[3]).tobytes() == np.float32(rb[3]).tobytes(), name


def test_degree_features_reference_goldens():
    """functionality_common.inc:6-15: the 4-vertex graph of the reference's feature tests --
    degrees {2, 1, 1, 0}... as transcribed in tests/golden/reference_vectors.json."""
    g = GOLD["functionality_common"]
    rp = np.asarray(g["row_ptr"], np.int32)
    col = np.asarray(g["cols"], np.int32)
    n = len(rp) - 1
    deg, dist, sc, avg = oracle_lib.restated().degree_features(n, rp, col)
    assert list(deg) == g["degrees"]
    assert np.array_equal(dist, np.asarray(g["distribution"], np.float32))
    assert sc["min_degree"] == min(g["degrees"]) and sc["max_degree"] == max(g["degrees"])


# ------------------------------------------------------------------ GPU
def boba_cases():
    """(name, n, m, row, col): (row, col)-sorted, unique COO lists -- symmetric graphs, and an
    asymmetric one with vertices that only appear as columns and vertices without entries."""
    out = []
    n, r, c = graphs.rmat(10, 8, seed=3)
    out.append(("rmat10", n, n, r, c))
    n, r, c = graphs.band(3000, 15, 0.5, seed=9, shuffle_seed=10)
    out.append(("band3k_shuffled", n, n, r, c))
    n, r, c = graphs.multi_component()
    out.append(("multi", n, n, r, c))
    rng = np.random.default_rng(5)
    key = np.unique(rng.integers(0, 200, 2000).astype(np.int64) * 500 + rng.integers(150, 450, 2000))
    out.append(("asym", 500, 500, key // 500, key % 500))
    return out


@needs_ref
def test_boba_restated_equals_reference():
    """Both variants of the reference (the parallel one run with one thread: its OpenMP loop
    updates the minima without atomics)."""
    orc, ref = oracle_lib.restated(), oracle_lib.reference()
    for name, n, m, row, col in boba_cases():
        for idt, nt, vt in ((np.int32, np.int32, np.float32), (np.int64, np.int64, np.float64)):
            r2, c2 = row.astype(idt), col.astype(idt)
            a = orc.boba_reorder(n, m, r2, c2, nnz_dtype=nt, vals_dtype=vt)
            assert sorted(a.tolist()) == list(range(max(n, m))), name
            for seq in (True, False):
                e = ref.boba_reorder(n, m, r2, c2, sequential=seq, nnz_dtype=nt, vals_dtype=vt)
                assert eq(a, e), (name, idt, seq)


HEAT_PARTS = (1, 2, 3, 7, 64, 100)


def heat_orders(n, idt, seed):
    rng = np.random.default_rng(seed)
    return rng.permutation(n).astype(idt), rng.permutation(n).astype(idt)


def test_reorder_heatmap_reference_goldens():
    """tests/suites/sparsebase/reorder/reorder_heatmap_tests.cc:25-82 (Instance,
    InstanceTwoReorders, InstanceDefaultConstructor: num_parts = 3) with the vectors of
    functionality_common.inc:6-52."""
    orc = oracle_lib.restated()
    rp = np.array([0, 2, 3, 4], np.int32)
    col = np.array([1, 2, 0, 0], np.int32)
    ident = np.arange(3, dtype=np.int32)
    no_order = np.array([0, 0.25, 0.25, 0.25, 0, 0, 0.25, 0, 0], np.float32)
    rc_order = np.array([0, 0, 0.25, 0.25, 0.25, 0, 0, 0, 0.25], np.float32)
    assert eq(orc.reorder_heatmap(3, rp, col, ident, ident, 3), no_order)
    r_vec, c_vec = np.array([1, 2, 0], np.int32), np.array([2, 0, 1], np.int32)
    assert eq(orc.reorder_heatmap(3, rp, col, r_vec, c_vec, 3), rc_order)
    assert orc.reorder_heatmap(3, rp, col, ident, ident, 4) is None  # the reference throws


@needs_ref
def test_reorder_heatmap_restated_equals_reference():
    orc, ref = oracle_lib.restated(), oracle_lib.reference()
    for name, n, rp, col in feature_graphs():
        for idt, nt, vt in ((np.int32, np.int32, np.float32), (np.int64, np.int64, np.float64),
                            (np.int32, np.int64, np.float32)):
            rp2, col2 = rp.astype(nt), col.astype(idt)
            pr, pc = heat_orders(n, idt, 11)
            for b in HEAT_PARTS + (n, n + 1):
                a = orc.reorder_heatmap(n, rp2, col2, pr, pc, b, vt)
                e = ref.reorder_heatmap(n, rp2, col2, pr, pc, b, vt)
                assert eq(a, e), (name, idt, b)


FEATURE_FIXTURES = ("rmat9", "poisson24x17", "band600", "multi")


@pytest.mark.parametrize("name", FEATURE_FIXTURES)
def test_restated_equals_committed_reference_fixtures(name):
    """tests/golden/ref_features_*.npz were written by the compiled reference
    (tests/golden/make_golden.py): the restatement must reproduce them where the reference tree
    is not mounted, too."""
    orc = oracle_lib.restated()
    f = np.load(os.path.join(HERE, "golden", f"ref_features_{name}.npz"))
    n, rp, col = int(f["n"]), f["row_ptr"], f["col"]
    deg, dist, sc, avg = orc.degree_features(n, rp, col)
    assert eq(deg, f["degrees"]) and eq(dist, f["dist"])
    assert [sc[k] for k in ("min_degree", "max_degree", "bandwidth", "profile")] == list(f["scalars"])
    assert np.float32(avg).tobytes() == f["avg"].astype(np.float32).tobytes()
    ident = np.arange(n, dtype=np.int32)
    assert eq(orc.reorder_heatmap(n, rp, col, f["rcm"], f["rcm"], 5), f["heat_rcm_5"])
    assert eq(orc.reorder_heatmap(n, rp, col, f["degree_asc"], f["degree_asc"], 3), f["heat_degree_3"])
    assert eq(orc.reorder_heatmap(n, rp, col, ident, ident, 16), f["heat_identity_16"])
    assert eq(orc.boba_reorder(n, n, f["coo_row"], f["coo_col"]), f["boba"])


def test_restated_edge_list_equals_committed_reference_fixture():
    orc = oracle_lib.restated()
    f = np.load(os.path.join(HERE, "golden", "ref_edge_list.npz"))
    for tag, flags in (("dedup", (True, False, False, False)), ("sym", (True, True, True, False)),
                       ("square", (False, True, False, True))):
        n, m, r, c, v = orc.edges_to_coo(f["u"], f["v"], f["w"], *flags)
        assert [n, m] == list(f[f"{tag}_dims"]), tag
        assert eq(r, f[f"{tag}_row"]) and eq(c, f[f"{tag}_col"]) and eq(v, f[f"{tag}_vals"]), tag


@pytest.fixture(scope="module")
def sb():
    from sparsebase_b200 import lib
    lib.load()
    return lib


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return None if t is None else t.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("flags", FLAGS, ids=[str(f) for f in FLAGS])
def test_edges_to_coo_gpu(sb, flags):
    orc = oracle_lib.restated()
    for name, u, v, w in edge_lists():
        for ww in (w, None):
            en, em, er, ec, ev = orc.edges_to_coo(u, v, ww, *flags)
            gn, gm, gr, gc, gv = sb.edges_to_coo(dev(u), dev(v), dev(ww), *flags)
            assert (gn, gm) == (en, em), (name, flags)
            assert eq(host(gr), er) and eq(host(gc), ec), f"{name} {flags} row/col"
            assert eq(host(gv), ev), f"{name} {flags} vals"
    # 64-bit ids, weights float64
    name, u, v, w = edge_lists()[1]
    u64, v64, w64 = u.astype(np.int64) * 70000, v.astype(np.int64) * 70000, w.astype(np.float64)
    en, em, er, ec, ev = orc.edges_to_coo(u64, v64, w64, *flags, nnz_dtype=np.int64)
    gn, gm, gr, gc, gv = sb.edges_to_coo(dev(u64), dev(v64), dev(w64), *flags)
    assert (gn, gm) == (en, em) and eq(host(gr), er) and eq(host(gc), ec) and eq(host(gv), ev)


@pytest.mark.gpu
def test_edges_to_coo_feeds_the_coo_constructor(sb):
    """The result is what format::COO's constructor wants: already (row, col)-sorted."""
    name, u, v, w = edge_lists()[2]
    n, m, r, c, vals = sb.edges_to_coo(dev(u), dev(v), dev(w), True, True, True, False)
    assert sb.coo_sort_(n, m, r.clone(), c.clone(), vals.clone()) is True
    key = host(r).astype(np.int64) * n + host(c)
    assert np.all(np.diff(key) > 0)


@pytest.mark.gpu
def test_degree_features_gpu(sb):
    orc = oracle_lib.restated()
    for name, n, rp, col in feature_graphs():
        for idt, nt, ft, tft in ((np.int32, np.int32, np.float32, torch.float32),
                                 (np.int64, np.int64, np.float64, torch.float64)):
            rp2, col2 = rp.astype(nt), col.astype(idt)
            edeg, edist, esc, eavg = orc.degree_features(
                n, rp2, col2, vals_dtype=np.float32 if idt == np.int32 else np.float64)
            tid = torch.int32 if idt == np.int32 else torch.int64
            gdeg, gdist, gsc = sb.degree_features(n, len(col2), dev(rp2), dev(col2), id_dtype=tid,
                                                  feature_dtype=tft)
            assert eq(host(gdeg), edeg) and eq(host(gdist), edist), name
            for k in ("min_degree", "max_degree", "bandwidth", "profile"):
                assert gsc[k] == esc[k], (name, k, gsc[k], esc[k])
            assert ft(gsc["avg_degree"]).tobytes() == ft(eavg).tobytes(), name
    # scalars only
    name, n, rp, col = feature_graphs()[0]
    _, _, sc = sb.degree_features(n, len(col), dev(rp), dev(col), want_arrays=False)
    assert sc["bandwidth"] == orc.degree_features(n, rp, col)[2]["bandwidth"]


@pytest.mark.gpu
def test_reorder_heatmap_gpu(sb):
    """sb200_reorder_heatmap against the oracle: shared-memory grid (num_parts <= 64), global
    grid (100), one cell, one row per block (num_parts = n), identity orders, 64-bit types, and
    the reference's exception."""
    from sparsebase_b200.lib import Sb200Error
    orc = oracle_lib.restated()
    for name, n, rp, col in feature_graphs():
        for idt, nt, vt, tft in ((np.int32, np.int32, np.float32, torch.float32),
                                 (np.int64, np.int64, np.float64, torch.float64),
                                 (np.int32, np.int64, np.float32, torch.float32)):
            rp2, col2 = rp.astype(nt), col.astype(idt)
            pr, pc = heat_orders(n, idt, 11)
            for b in HEAT_PARTS + ((n,) if n <= 2000 else ()):
                exp = orc.reorder_heatmap(n, rp2, col2, pr, pc, b, vt)
                got = sb.reorder_heatmap(n, n, dev(rp2), dev(col2), dev(pr), dev(pc), b, tft)
                assert eq(host(got), exp), (name, idt, b)
            ident = np.arange(n, dtype=idt)
            exp = orc.reorder_heatmap(n, rp2, col2, ident, ident, 5, vt)
            got = sb.reorder_heatmap(n, n, dev(rp2), dev(col2), None, None, 5, tft)
            assert eq(host(got), exp), (name, "identity")
    name, n, rp, col = feature_graphs()[0]
    with pytest.raises(Sb200Error):
        sb.reorder_heatmap(n, n, dev(rp), dev(col), None, None, n + 1)
    # a heavier matrix: hub rows spanning many 4096-entry chunks
    n, r, c = graphs.rmat(15, 8, seed=21)
    rp, pr = graphs.csr_of(n, r, c), sb.degree_reorder(n, dev(graphs.csr_of(n, r, c)), True)
    exp = orc.reorder_heatmap(n, rp, c, host(pr), host(pr), 16)
    assert eq(host(sb.reorder_heatmap(n, n, dev(rp), dev(c), pr, pr, 16)), exp)


@pytest.mark.gpu
def test_boba_reorder_gpu(sb):
    orc = oracle_lib.restated()
    cases = boba_cases()
    n, r, c = graphs.rmat(15, 8, seed=21)
    cases.append(("rmat15", n, n, r, c))
    cases.append(("empty", 7, 7, np.zeros(0, np.int64), np.zeros(0, np.int64)))
    for name, n, m, row, col in cases:
        for idt, nt, vt in ((np.int32, np.int32, np.float32), (np.int64, np.int64, np.float64)):
            r2, c2 = row.astype(idt), col.astype(idt)
            exp = orc.boba_reorder(n, m, r2, c2, nnz_dtype=nt, vals_dtype=vt)
            got = sb.boba_reorder(n, m, dev(r2), dev(c2))
            assert eq(host(got), exp), (name, idt)
