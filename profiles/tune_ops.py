"""Per-operator timings (CUDA events, device-resident inputs) on one synthetic graph.

    python profiles/tune_ops.py --graph poisson --size 4096 --ops csr_to_csc,coo_to_csr,...
Prints one JSON line; used for A/B runs of kernel variants selected by SB200_* env variables.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sparsebase_b200 import lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--graph", default="poisson")
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--ops", default="csr_to_csc,coo_to_csr,degree_reorder,permute2d_rand,coo_sort")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)
if args.graph == "poisson":
    n, rp, col, vals = synth.poisson2d(args.size, args.size, device=dev)
elif args.graph == "band":
    n, row, col = synth.band(1 << args.size, 31, 0.5, seed=45, shuffle_seed=46, device=dev)
    rp = synth.csr_from_sorted_coo(n, row)
    vals = synth.hash_vals(col.numel(), device=dev)
elif args.graph == "er":
    n, row, col = synth.erdos_renyi(1 << args.size, 8, device=dev)
    rp = synth.csr_from_sorted_coo(n, row)
    vals = synth.hash_vals(col.numel(), device=dev)
else:
    n, row, col = synth.rmat(args.size, 8, device=dev)
    rp = synth.csr_from_sorted_coo(n, row)
    vals = synth.hash_vals(col.numel(), device=dev)
nnz = col.numel()
row = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int32),
                              (rp[1:] - rp[:-1]).to(torch.int64))
g = torch.Generator(device=dev)
g.manual_seed(1)
perm = torch.randperm(n, generator=g, device=dev).to(torch.int32)
shuf = torch.randperm(nnz, generator=g, device=dev)
urow, ucol, uvals = row[shuf].contiguous(), col[shuf].contiguous(), vals[shuf].contiguous()
del shuf


def timed(fn, reps=args.reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def coo_sort():
    r, c, v = urow.clone(), ucol.clone(), uvals.clone()
    lib.coo_sort_(n, n, r, c, v)


FN = {
    "csr_to_csc": lambda: lib.csr_to_csc(n, n, rp, col, vals),
    "coo_to_csr": lambda: lib.coo_to_csr(n, n, row, col, vals),
    "degree_reorder": lambda: lib.degree_reorder(n, rp, True),
    "permute2d_rand": lambda: lib.permute2d(n, n, rp, col, vals, perm, perm),
    "coo_sort": coo_sort,
    "rcm": lambda: lib.rcm_reorder(n, rp, col),
    "csr_to_coo": lambda: lib.csr_to_coo(n, n, rp, col, vals),
}
if "permute2d_rcm" in args.ops:
    rcm_perm = lib.rcm_reorder(n, rp, col)
    FN["permute2d_rcm"] = lambda: lib.permute2d(n, n, rp, col, vals, rcm_perm, rcm_perm)
if "permute2d_deg" in args.ops:
    deg_perm = lib.degree_reorder(n, rp, True)
    FN["permute2d_deg"] = lambda: lib.permute2d(n, n, rp, col, vals, deg_perm, deg_perm)
out = {"graph": args.graph, "size": args.size, "n": n, "nnz": nnz,
       "env": {k: v for k, v in os.environ.items() if k.startswith("SB200_")}}
for op in args.ops.split(","):
    out[op + "_ms"] = round(timed(FN[op]), 4)
print(json.dumps(out))
