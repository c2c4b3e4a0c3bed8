"""torchrun worker: the sharded operators with the real kernels over NCCL, compared
bit-for-bit with the single-GPU operators run on the whole matrix by the same rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/sharded_gpu_worker.py [--scale 16]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsebase_b200 import lib, sharded, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=15)
    ap.add_argument("--graph", default="rmat")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib.load()
    if args.graph == "rmat":
        n, row, col = synth.rmat(args.scale, 8, seed=21, device=dev)
    else:
        n, row, col = synth.erdos_renyi(1 << args.scale, 8, seed=22, device=dev)
    vals = synth.hash_vals(col.numel(), seed=5, device=dev)
    nnz = col.numel()
    eq = lambda a, b: a.dtype == b.dtype and a.shape == b.shape and bool((a == b).all())  # noqa

    # single-GPU results on the whole matrix (the parity-tested operators)
    g_rp, g_col, g_val = lib.coo_to_csr(n, n, row, col, vals)
    bounds = lib.partition_rows(n, nnz, g_rp, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    a, b = int(g_rp[lo]), int(g_rp[hi])
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)
    p = torch.randperm(b - a, generator=g, device=dev)
    s = sharded.coo_to_csr(lib, n, n, bounds, row[a:b][p], col[a:b][p], vals[a:b][p])
    assert s.nnz == nnz and s.nnz_base == a
    assert eq(s.global_row_ptr(), g_rp), "sharded COO->CSR row_ptr"
    assert eq(s.col, g_col[a:b]) and eq(s.vals, g_val[a:b]), "sharded COO->CSR col/vals"

    assert eq(sharded.degree_distribution(lib, s), lib.degree_distribution(n, nnz, g_rp)[lo:hi])
    assert eq(sharded.degrees(lib, s), lib.degrees(n, g_rp)[lo:hi])
    for asc in (True, False):
        assert eq(sharded.degree_reorder(lib, s, asc), lib.degree_reorder(n, g_rp, asc)), asc
    inv = sharded.degree_reorder(lib, s, True)

    ps = sharded.permute2d(lib, s, inv, inv)
    e_rp, e_col, e_val = lib.permute2d(n, n, g_rp, g_col, g_val, inv, inv)
    nlo, nhi = ps.bounds[rank], ps.bounds[rank + 1]
    a2, b2 = int(e_rp[nlo]), int(e_rp[nhi])
    assert ps.nnz_base == a2
    assert eq(ps.row_ptr + a2, e_rp[nlo:nhi + 1]), "sharded Permute2D row_ptr"
    assert eq(ps.col, e_col[a2:b2]) and eq(ps.vals, e_val[a2:b2]), "sharded Permute2D col/vals"

    cs = sharded.csr_to_csc(lib, s)
    c_cp, c_row, c_val = lib.csr_to_csc(n, n, g_rp, g_col, g_val)
    clo, chi = cs.bounds[rank], cs.bounds[rank + 1]
    a3, b3 = int(c_cp[clo]), int(c_cp[chi])
    assert cs.nnz_base == a3
    assert eq(cs.col_ptr + a3, c_cp[clo:chi + 1]), "sharded CSR->CSC col_ptr"
    assert eq(cs.row, c_row[a3:b3]) and eq(cs.vals, c_val[a3:b3]), "sharded CSR->CSC row/vals"
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"SHARDED OK world={world} n={n} nnz={nnz}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
