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
        print(name, n, len(row))


if __name__ == "__main__":
    main()
