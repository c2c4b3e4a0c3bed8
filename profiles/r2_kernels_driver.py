"""Every kernel of libsb200.so once, for the ncu evidence (profiles/r2_kernels_*.md).

    ncu --profile-from-start off --set full --clock-control none -f -o gpurun_out/r2_kernels \
        python profiles/r2_kernels_driver.py

Graph generation happens before cudaProfilerStart; between start and stop only library calls run
(plus the few torch kernels that build their arguments).  Inputs are small relatives of the
BASELINE shapes: Poisson 1024^2 (rows <= 8 entries: short-row Permute2D, narrow RCM), shuffled
band n = 1 M (rows <= 64: mid-row Permute2D), R-MAT scale 20 (hub rows: tile + long-row paths,
wide RCM levels, DegreeReorder tail), Erdos-Renyi 2^20.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsebase_b200 import lib, mg, synth  # noqa: E402

dev = torch.device("cuda", 0)
lib.load()
g = torch.Generator(device=dev)
g.manual_seed(1)

pn, prp, pcol, pval = synth.poisson2d(1024, 1024, device=dev)
bn, brow, bcol = synth.band(1_000_000, 31, 0.5, seed=45, shuffle_seed=46, device=dev)
bval = synth.hash_vals(bcol.numel(), seed=7, device=dev)
rn, rrow, rcol = synth.rmat(20, 8, seed=44, device=dev)
rval = synth.hash_vals(rcol.numel(), seed=7, device=dev)
sn, srow, scol = synth.rmat(16, 8, seed=42, device=dev)   # (wide RCM levels: many launches)
en, erow, ecol = synth.erdos_renyi(1 << 20, 8, seed=43, device=dev)
eval_ = synth.hash_vals(ecol.numel(), seed=7, device=dev)
# a raw edge list for the reader path
eu = torch.randint(0, 1 << 20, (4_000_000,), generator=g, device=dev, dtype=torch.int32)
ev = torch.randint(0, 1 << 20, (4_000_000,), generator=g, device=dev, dtype=torch.int32)
ew = synth.hash_vals(4_000_000, seed=3, device=dev)
torch.cuda.synchronize()

torch.cuda.profiler.start()
# ---- Poisson: RCM (narrow regime), short-row Permute2D, conversions, features
inv = lib.rcm_reorder(pn, prp, pcol)
p = lib.permute2d(pn, pn, prp, pcol, pval, inv, inv)
prow, _, _ = lib.csr_to_coo(pn, pn, prp, pcol, pval)
lib.coo_to_csr(pn, pn, prow, pcol, pval)
lib.coo_to_csc(pn, pn, prow, pcol, pval)
lib.csr_to_csc(pn, pn, prp, pcol, pval)
lib.degrees(pn, prp)
lib.degree_distribution(pn, pcol.numel(), prp)
lib.degree_features(pn, pcol.numel(), p[0], p[1])
lib.degree_reorder(pn, prp, True)
x = pval[:pn].contiguous()
lib.permute1d(x, inv)
lib.inverse_permutation(inv)
lib.partition_rows(pn, pcol.numel(), prp, 8)
# ---- band: mid-row Permute2D + unsorted constructors
brp, bc, bv = lib.coo_to_csr(bn, bn, brow, bcol, bval)
binv = lib.degree_reorder(bn, brp, False)
lib.permute2d(bn, bn, brp, bc, bv, binv, binv)
sh = torch.randperm(bcol.numel(), generator=g, device=dev)
r2, c2, v2 = brow[sh].contiguous(), bcol[sh].contiguous(), bval[sh].contiguous()
lib.coo_sort_(bn, bn, r2, c2, v2)                       # COO constructor on a shuffled list
c3 = binv[bc.to(torch.int64)].contiguous()
lib.compressed_sort_(bn, bn, brp, c3, bv.clone())       # CSR constructor on renumbered rows
# ---- R-MAT: tile + long-row Permute2D, DegreeReorder tail, wide RCM levels, CSR->CSC
rrp, rc, rv = lib.coo_to_csr(rn, rn, rrow, rcol, rval)
rinv = lib.degree_reorder(rn, rrp, True)
lib.permute2d(rn, rn, rrp, rc, rv, rinv, rinv)
lib.csr_to_csc(rn, rn, rrp, rc, rv)
c4 = rinv[rc.to(torch.int64)].contiguous()
lib.compressed_sort_(rn, rn, rrp, c4, rv.clone())
lib.degree_features(rn, rc.numel(), rrp, rc)
# ---- edge list reader path
lib.edges_to_coo(eu, ev, ew, True, True, True, False)
# ---- multi-GPU building blocks and the peer-memory operators (world = 1: all stores local)
lib.coo_to_csr_block(0, en, en, erow, ecol, eval_)
erp, ec, ev2 = lib.coo_to_csr(en, en, erow, ecol, eval_)
lib.csr_to_csc_block(0, en, en, erp, ec, ev2)
lib.rank_keys(rinv, rn)
lib.degree_histogram(en, erp, lib.max_degree(en, erp) + 1)
comm = mg.Comm(3 * rc.numel() * 8 + 16 * rn + (256 << 20))
s = mg.coo_to_csr(comm, rn, rn, [0, rn], rrow, rcol, rval, presorted=True)
minv = mg.degree_reorder(comm, s, True)
mg.permute2d(comm, s, minv, minv)
mg.csr_to_csc(comm, s)
mg.permute1d(comm, [0, rn], rval[:rn].contiguous(), minv)
# ---- last (hundreds of launches of the same few kernels; the capture may stop inside): the
# wide regime of RCM
srp = synth.csr_from_sorted_coo(sn, srow)
lib.rcm_reorder(sn, srp, scol)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
comm.destroy()
print("done")
