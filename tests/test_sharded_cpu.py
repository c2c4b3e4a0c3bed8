"""world_size-2 (and 4) gloo tests of the multi-GPU host logic (sparsebase_b200/sharded.py) on CPU.

Each rank runs the sharded operator on its row block with tests/cpu_ops.py standing in for the
CUDA kernels; rank results are concatenated and compared bit-for-bit with the oracle run on
the whole matrix.  (The same comparison with the real kernels over NCCL is
tests/test_sharded_gpu.py.)"""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _case(name):
    import graphs
    if name == "rmat":
        n, r, c = graphs.rmat(10, 8, seed=11)
    elif name == "er":
        n, r, c = graphs.er(3000, 6, seed=12)
    else:
        n, rp, col, _ = graphs.poisson(37, 29)
        r, c = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp)), col
    vals = graphs.vals_for(len(r), seed=3)
    return n, r, c, vals


def _worker(rank, world, name, initfile, outdir):
    import cpu_ops
    import graphs
    import oracle_lib
    from sparsebase_b200 import sharded
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    try:
        orc = oracle_lib.restated()
        n, r, c, vals = _case(name)
        nnz = len(r)
        rp = graphs.csr_of(n, r, c)
        # rank-local slice of the COO, delivered shuffled (exercises the local ctor sort)
        bounds = cpu_ops.partition_rows(n, nnz, torch.from_numpy(rp), world)
        lo, hi = bounds[rank], bounds[rank + 1]
        sl = slice(int(rp[lo]), int(rp[hi]))
        perm = np.random.default_rng(5 + rank).permutation(sl.stop - sl.start)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
        s = sharded.coo_to_csr(cpu_ops, n, n, bounds, t(r[sl][perm]), t(c[sl][perm]),
                               t(vals[sl][perm]))
        # ---- COO -> CSR
        g_rp, g_col, g_val = orc.coo_to_csr(n, n, r, c, vals)
        assert s.nnz == nnz and s.nnz_base == int(rp[lo])
        assert np.array_equal(s.global_row_ptr().numpy(), g_rp)
        assert np.array_equal(s.col.numpy(), g_col[sl]) and np.array_equal(s.vals.numpy(), g_val[sl])
        # ---- features
        dd = sharded.degree_distribution(cpu_ops, s).numpy()
        assert np.array_equal(dd, orc.degree_distribution(n, g_rp, g_col)[lo:hi])
        # ---- DegreeReorder, both directions: full permutation on every rank
        for asc in (True, False):
            inv = sharded.degree_reorder(cpu_ops, s, asc).numpy()
            exp = orc.degree_reorder(n, g_rp, g_col, asc)
            assert np.array_equal(inv, exp), ("degree", asc, np.nonzero(inv != exp)[0][:8],
                                              inv[inv != exp][:8], exp[inv != exp][:8])
        inv = sharded.degree_reorder(cpu_ops, s, True)
        # ---- Permute2D with the degree order
        p = sharded.permute2d(cpu_ops, s, inv, inv)
        e_rp, e_col, e_val = orc.permute2d(n, n, g_rp, g_col, g_val, inv.numpy(), inv.numpy())
        nlo, nhi = p.bounds[rank], p.bounds[rank + 1]
        assert p.nnz_base == int(e_rp[nlo])
        assert np.array_equal(p.row_ptr.numpy() + p.nnz_base, e_rp[nlo:nhi + 1])
        ps = slice(int(e_rp[nlo]), int(e_rp[nhi]))
        assert np.array_equal(p.col.numpy(), e_col[ps]) and np.array_equal(p.vals.numpy(), e_val[ps])
        # the new-row blocks are nnz-balanced too
        assert abs((ps.stop - ps.start) - nnz / world) <= nnz / world * 0.5 + 64
        # ---- CSR -> CSC
        q = sharded.csr_to_csc(cpu_ops, s)
        c_cp, c_row, c_val = orc.csr_to_csc(n, n, g_rp, g_col, g_val)
        clo, chi = q.bounds[rank], q.bounds[rank + 1]
        assert q.nnz_base == int(c_cp[clo])
        assert np.array_equal(q.col_ptr.numpy() + q.nnz_base, c_cp[clo:chi + 1])
        cs = slice(int(c_cp[clo]), int(c_cp[chi]))
        assert np.array_equal(q.row.numpy(), c_row[cs]) and np.array_equal(q.vals.numpy(), c_val[cs])
        open(os.path.join(outdir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["poisson", "rmat", "er"])
def test_sharded_operators_world2_gloo(name):
    world = 2
    with tempfile.TemporaryDirectory() as d:
        initfile = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, name, initfile, d), nprocs=world, join=True)
        assert all(os.path.exists(os.path.join(d, f"ok{r}")) for r in range(world))


@pytest.mark.parametrize("name", ["poisson", "rmat"])
def test_sharded_operators_world4_gloo(name):
    """Four ranks: on the banded Poisson matrix distant row blocks exchange nothing (zero-size
    all_to_all splits); this is the rank count of the driver's scaling runs beyond two."""
    world = 4
    with tempfile.TemporaryDirectory() as d:
        initfile = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, name, initfile, d), nprocs=world, join=True)
        assert all(os.path.exists(os.path.join(d, f"ok{r}")) for r in range(world))


def test_sharded_single_process_matches_too():
    """world == 1 (no process group): the same code path degenerates to the local operators."""
    import cpu_ops
    import graphs
    import oracle_lib
    from sparsebase_b200 import sharded
    orc = oracle_lib.restated()
    n, r, c, vals = _case("rmat")
    t = torch.from_numpy
    s = sharded.coo_to_csr(cpu_ops, n, n, [0, n], t(r), t(c), t(vals))
    g_rp, g_col, g_val = orc.coo_to_csr(n, n, r, c, vals)
    assert np.array_equal(s.row_ptr.numpy(), g_rp)
    inv = sharded.degree_reorder(cpu_ops, s, False)
    assert np.array_equal(inv.numpy(), orc.degree_reorder(n, g_rp, g_col, False))
    p = sharded.permute2d(cpu_ops, s, inv, inv)
    e = orc.permute2d(n, n, g_rp, g_col, g_val, inv.numpy(), inv.numpy())
    assert np.array_equal(p.row_ptr.numpy(), e[0]) and np.array_equal(p.col.numpy(), e[1])
    q = sharded.csr_to_csc(cpu_ops, s)
    ce = orc.csr_to_csc(n, n, g_rp, g_col, g_val)
    assert np.array_equal(q.col_ptr.numpy(), ce[0]) and np.array_equal(q.row.numpy(), ce[1])
