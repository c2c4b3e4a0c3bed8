"""The peer-memory operators with world = 1 (every "peer" store is local): isolates the kernel
structure from the NVLink path.  For `ncu --metrics gpu__time_duration.sum` launch lists.

    python profiles/mg_single_probe.py [--scale 23]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsebase_b200 import lib, mg, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=23)
args = ap.parse_args()
dev = torch.device("cuda", 0)
lib.load()
n, row, col = synth.rmat(args.scale, 8, seed=21, device=dev)
vals = synth.hash_vals(col.numel(), seed=5, device=dev)
nnz = col.numel()
comm = mg.Comm(3 * nnz * 8 + 8 * n + (256 << 20))
s = mg.coo_to_csr(comm, n, n, [0, n], row, col, vals, presorted=True)
inv = mg.degree_reorder(comm, s, True)
for _ in range(2):
    p = mg.permute2d(comm, s, inv, inv)
    t = mg.csr_to_csc(comm, s)
torch.cuda.synchronize()
e = lib.permute2d(n, n, s.row_ptr, s.col, s.vals, inv, inv)
print("permute2d equal:", torch.equal(p.col, e[1]) and torch.equal(p.vals, e[2]))
e = lib.csr_to_csc(n, n, s.row_ptr, s.col, s.vals)
print("csr_to_csc equal:", torch.equal(t.row, e[1]) and torch.equal(t.vals, e[2]))
comm.destroy()
