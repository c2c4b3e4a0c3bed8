"""Row a1: the COO constructor's check + (row, col) sort (sb200_coo_sort) on a randomly ordered
COO, and its check-only path on a sorted one.   python profiles/a1_time.py 16 26"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsebase_b200 import lib, synth  # noqa: E402

dev = torch.device("cuda", 0)
lib.load()
out = []
for scale in [int(a) for a in sys.argv[1:]] or [16]:
    n, row, col = synth.rmat(scale, 8 if scale > 20 else 16, seed=44 if scale > 20 else 42, device=dev)
    vals = synth.hash_vals(col.numel(), seed=7, device=dev)
    nnz = col.numel()
    res = {"graph": f"R-MAT scale {scale}", "n": n, "nnz": nnz}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for rep in range(2):
        g = torch.Generator(device=dev)
        g.manual_seed(100 + rep)
        perm = torch.randperm(nnz, generator=g, device=dev)
        r2, c2, v2 = row[perm], col[perm], vals[perm]
        del perm
        torch.cuda.synchronize()
        ev[0].record()
        was_sorted = lib.coo_sort_(n, n, r2, c2, v2)
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]))
        ok = (not was_sorted) and torch.equal(r2, row) and torch.equal(c2, col) and torch.equal(v2, vals)
        del r2, c2, v2
        torch.cuda.empty_cache()
    res["random_order_ms"] = min(times)
    res["random_order_gnnz_per_s"] = nnz / (min(times) * 1e-3) / 1e9
    res["restores_the_sorted_list"] = bool(ok)
    r2, c2, v2 = row.clone(), col.clone(), vals.clone()
    torch.cuda.synchronize()
    ev[0].record()
    was_sorted = lib.coo_sort_(n, n, r2, c2, v2)
    ev[1].record()
    torch.cuda.synchronize()
    res["already_sorted_ms"] = ev[0].elapsed_time(ev[1])
    res["already_sorted_detected"] = bool(was_sorted)
    out.append(res)
    del row, col, vals, r2, c2, v2
    torch.cuda.empty_cache()
    lib.trim()
print("A1_JSON " + json.dumps(out))
