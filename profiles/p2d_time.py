"""Permute2D with the DegreeReorder permutation on an R-MAT graph (the C4 shape): ms per call.
    python profiles/p2d_time.py --scale 26 [--graph rmat|er]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsebase_b200 import lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=26)
ap.add_argument("--graph", default="rmat")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--classes", action="store_true", help="print the nnz share of every row-length class")
args = ap.parse_args()
dev = torch.device("cuda", 0)
lib.load()
if args.graph == "rmat":
    n, row, col = synth.rmat(args.scale, 8, seed=44, device=dev)
else:
    n, row, col = synth.erdos_renyi(1 << args.scale, 8, seed=43, device=dev)
vals = synth.hash_vals(col.numel(), seed=7, device=dev)
rp, cc, cv = lib.coo_to_csr(n, n, row, col, vals)
del row, col, vals
inv = lib.degree_reorder(n, rp, True)
if args.classes:
    deg = (rp[1:] - rp[:-1]).to(torch.int64)
    edges = [0, 8, 32, 64, 128, 256, 512, 1024, 2048, 4096, 6144, 8192, 12288, 16384, 32768, 65536,
             1 << 62]
    tot = int(deg.sum())
    for lo_, hi_ in zip(edges[:-1], edges[1:]):
        msk = (deg > lo_) & (deg <= hi_)
        print(f"P2D_CLASS ({lo_},{hi_}] rows={int(msk.sum())} nnz_share={float(deg[msk].sum()) / tot:.4f}")
    del deg
out = (torch.empty_like(rp), torch.empty_like(cc), torch.empty_like(cv))
lib.permute2d(n, n, rp, cc, cv, inv, inv, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()  # (ncu --profile-from-start off: only the timed calls are captured)
a.record()
for _ in range(args.reps):
    lib.permute2d(n, n, rp, cc, cv, inv, inv, out=out)
b.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"P2D_TIME graph={args.graph} scale={args.scale} nnz={cc.numel()} split={os.environ.get('SB200_P2D_SPLIT', 'auto')} "
      f"ms={a.elapsed_time(b) / args.reps:.3f} checksum={int(out[1][::1000003].to(torch.int64).sum())}")
