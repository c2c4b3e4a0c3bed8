"""Times sb200_csr_to_csc_block on one GPU against sb200_csr_to_csc on the same row block."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparsebase_b200 import lib, synth
dev = torch.device("cuda", 0)
n, rp, col, vals = synth.poisson2d(4096, 4096, device=dev)
nnz = col.numel()
b = lib.partition_rows(n, nnz, rp, 2)
lo, hi = b[0], b[1]
a0, a1 = int(rp[lo]), int(rp[hi])
rp_l = (rp[lo:hi + 1] - a0).contiguous()
col_l, val_l = col[a0:a1].contiguous(), vals[a0:a1].contiguous()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {"block_ms": timed(lambda: lib.csr_to_csc_block(lo, hi - lo, n, rp_l, col_l, val_l)),
       "full_on_block_ms": timed(lambda: lib.csr_to_csc(hi - lo, n, rp_l, col_l, val_l)) if False else None,
       "full_ms": timed(lambda: lib.csr_to_csc(n, n, rp, col, vals))}
print(json.dumps(out))
