"""Tiny driver for ncu captures: runs selected operators once on a Poisson grid.

    ncu --set full --import-source on -k regex:<kernel> -c 2 -o gpurun_out/prof \
        python profiles/prof_driver.py --ops permute2d,csr_to_csc --grid 4096
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sparsebase_b200 import lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ops", default="permute2d,csr_to_csc,coo_to_csr,degree_reorder")
ap.add_argument("--grid", type=int, default=4096)
ap.add_argument("--graph", default="poisson")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--perm", default="random", help="random | degree (DegreeReorder of the matrix)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
if args.graph == "poisson":
    n, rp, col, vals = synth.poisson2d(args.grid, args.grid, device=dev)
elif args.graph == "rmat":  # --grid = scale
    n, row, col = synth.rmat(args.grid, 8, device=dev)
    rp = synth.csr_from_sorted_coo(n, row)
    vals = synth.hash_vals(col.numel(), device=dev)
else:  # er: --grid = log2(n)
    n, row, col = synth.erdos_renyi(1 << args.grid, 8, device=dev)
    rp = synth.csr_from_sorted_coo(n, row)
    vals = synth.hash_vals(col.numel(), device=dev)
nnz = col.numel()
row = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int32),
                              (rp[1:] - rp[:-1]).to(torch.int64))
g = torch.Generator(device=dev)
g.manual_seed(1)
perm = torch.randperm(n, generator=g, device=dev).to(torch.int32)
if args.perm == "degree":
    perm = lib.degree_reorder(n, rp, True)
ops = args.ops.split(",")
for _ in range(args.reps):
    if "rcm" in ops:
        perm = lib.rcm_reorder(n, rp, col)
    if "permute2d" in ops:
        lib.permute2d(n, n, rp, col, vals, perm, perm)
    if "csr_to_csc" in ops:
        lib.csr_to_csc(n, n, rp, col, vals)
    if "coo_to_csr" in ops:
        lib.coo_to_csr(n, n, row, col, vals)
    if "degree_reorder" in ops:
        lib.degree_reorder(n, rp, True)
    if "coo_sort" in ops:  # COO constructor on a shuffled edge list
        sh = torch.randperm(nnz, generator=g, device=dev)
        r2, c2, v2 = row[sh].contiguous(), col[sh].contiguous(), vals[sh].contiguous()
        lib.coo_sort_(n, n, r2, c2, v2)
    if "compressed_sort" in ops:  # CSR constructor on rows whose columns were renumbered
        c3 = perm[col.to(torch.int64)].contiguous()
        lib.compressed_sort_(n, n, rp, c3, vals.clone())
    if "features" in ops:
        lib.degrees(n, rp)
        lib.degree_distribution(n, nnz, rp)
        lib.permute1d(vals[:n].contiguous(), perm)
        lib.inverse_permutation(perm)
        lib.csr_to_coo(n, n, rp, col, vals)
        lib.coo_to_csc(n, n, row, col, vals)
torch.cuda.synchronize()
print("done", n, nnz)
