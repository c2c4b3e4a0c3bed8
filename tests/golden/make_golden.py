"""Generate the committed reference fixtures tests/golden/ref_*.npz.

Runs the UNMODIFIED reference (oracle/_ref/libsbref.so, built by `make -C oracle ref` from
/root/reference/src) on small seeded graphs and stores inputs + outputs.  Needs the reference
tree, so it runs in the development container only; the fixtures travel with the repo.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import graphs  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    ref = oracle_lib.reference()
    assert ref is not None, "build oracle/_ref first: make -C oracle ref"
    cases = {}
    n, r, c = graphs.rmat(9, 8, seed=101)
    cases["rmat9"] = (n, r, c)
    n, rp, col, _ = graphs.poisson(24, 17)
    cases["poisson24x17"] = (n, np.repeat(np.arange(n, dtype=np.int32), np.diff(rp)), col)
    n, r, c = graphs.band(600, 5, 0.5, seed=103, shuffle_seed=104)
    cases["band600"] = (n, r, c)
    n, r, c = graphs.multi_component(seed=105)
    cases["multi"] = (n, r, c)
    for name, (n, row, col) in cases.items():
        vals = graphs.vals_for(len(row), seed=7)
        rp, cc, vv = ref.coo_to_csr(n, n, row, col, vals)
        rcm = ref.rcm_reorder(n, rp, cc)
        p2d = ref.permute2d(n, n, rp, cc, vv, rcm, rcm)
        csc = ref.csr_to_csc(n, n, rp, cc, vv)
        np.savez_compressed(
            os.path.join(HERE, f"ref_{name}.npz"), n=n, row_ptr=rp, col=cc, vals=vv,
            degree_asc=ref.degree_reorder(n, rp, cc, True),
            degree_desc=ref.degree_reorder(n, rp, cc, False), rcm=rcm, p2d_row_ptr=p2d[0],
            p2d_col=p2d[1], p2d_vals=p2d[2], csc_col_ptr=csc[0], csc_row=csc[1], csc_vals=csc[2],
            degree_distribution=ref.degree_distribution(n, rp, cc))
        # the callers either side of the path: fused features, heatmaps under two
        # orders, BOBA
        fdeg, fdist, fsc, favg = ref.degree_features(n, rp, cc)
        deg_asc = ref.degree_reorder(n, rp, cc, True)
        np.savez_compressed(
            os.path.join(HERE, f"ref_features_{name}.npz"), n=n, row_ptr=rp, col=cc,
            coo_row=row, coo_col=col, degrees=fdeg, dist=fdist,
            scalars=np.array([fsc[k] for k in ("min_degree", "max_degree", "bandwidth", "profile")],
                             np.int64),
            avg=np.array([favg]), rcm=rcm, degree_asc=deg_asc,
            heat_rcm_5=ref.reorder_heatmap(n, rp, cc, rcm, rcm, 5),
            heat_degree_3=ref.reorder_heatmap(n, rp, cc, deg_asc, deg_asc, 3),
            heat_identity_16=ref.reorder_heatmap(n, rp, cc, np.arange(n, dtype=np.int32),
                                                 np.arange(n, dtype=np.int32), 16),
            boba=ref.boba_reorder(n, n, row, col, sequential=True))
        print(name, n, len(row))
    # an edge list through EdgeListReader::ReadCOO (weights a function of the unordered pair)
    rng = np.random.default_rng(107)
    u = rng.integers(0, 300, 4000).astype(np.int32)
    v = rng.integers(0, 300, 4000).astype(np.int32)
    lo, hi = np.minimum(u, v).astype(np.int64), np.maximum(u, v).astype(np.int64)
    w = (((lo * 2654435761 + hi * 40503) % 4096).astype(np.float32) - 2048.0) * 0.5
    out = {"u": u, "v": v, "w": w}
    for tag, flags in (("dedup", (True, False, False, False)), ("sym", (True, True, True, False)),
                       ("square", (False, True, False, True))):
        en, em, er, ec, evv = ref.edges_to_coo(u, v, w, *flags)
        out.update({f"{tag}_dims": np.array([en, em], np.int64), f"{tag}_row": er, f"{tag}_col": ec,
                    f"{tag}_vals": evv})
    np.savez_compressed(os.path.join(HERE, "ref_edge_list.npz"), **out)


if __name__ == "__main__":
    main()
