#!/usr/bin/env python
"""Per-operator timings + size-independent parity properties on BASELINE.json's other configs
(C1 R-MAT-16, C3 ER 2^24 x 16, C4 R-MAT-26, C5 shuffled band) on ONE GPU.

    python profiles/bench_configs.py --config C3 [--scale S] [--reps R] [--rcm]

bench.py's headline stays C2 (the contract's N=1 workload); this script produces the
per-config tables kept under profiles/.  Every operator goes through the ctypes binding of the
C ABI with device-resident inputs, timed with CUDA events on the launching stream.  The parity
properties checked at full size (no CPU oracle at 1 B nnz): transpose twice = identity,
Permute2D by inv then by its inverse = identity, permutations are bijections, row_ptr/col_ptr
totals, degree order is monotone with the reference's tie rule, DegreeDistribution sums to ~1.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsebase_b200 import lib, synth  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(reps):
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None or ms < best else best
    return best, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--scale", type=int, default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rcm", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib.load()
    cfg = args.config.upper()
    if cfg == "C1":
        s = args.scale or 16
        n, row, col = synth.rmat(s, 16, seed=42, device=dev)
        name = f"C1 R-MAT scale {s}, edge factor 16"
    elif cfg == "C3":
        s = args.scale or 24
        n, row, col = synth.erdos_renyi(1 << s, 8, seed=43, device=dev)
        name = f"C3 Erdos-Renyi 2^{s} vertices, 8 undirected pairs per vertex"
    elif cfg == "C4":
        s = args.scale or 26
        n, row, col = synth.rmat(s, 8, seed=44, device=dev)
        name = f"C4 R-MAT scale {s}, edge factor 8 undirected"
    elif cfg == "C5":
        n = (args.scale and (1 << args.scale)) or 50_000_000
        n, row, col = synth.band(n, 31, 0.5, seed=45, shuffle_seed=46, device=dev)
        name = f"C5 band n={n}, |i-j|<=31 at density 0.5, random symmetric relabelling"
    else:
        raise SystemExit("config must be C1, C3, C4 or C5")
    torch.cuda.empty_cache()
    nnz = col.numel()
    vals = synth.hash_vals(nnz, seed=7, device=dev)
    pk = peak()
    I = 4
    alg = {
        "coo_to_csr": nnz * 20 + (n + 1) * 4,
        "csr_to_csc": nnz * 16 + 2 * (n + 1) * 4,
        "coo_to_csc": nnz * 20 + (n + 1) * 4,
        "csr_to_coo": nnz * 20 + (n + 1) * 4,
        "permute2d": nnz * 16 + 2 * (n + 1) * 4 + 2 * n * I,
        "degree_reorder": (n + 1) * 4 + n * I,
        "degree_distribution": (n + 1) * 4 + n * 4,
        "degrees": (n + 1) * 4 + n * I,
    }
    ops, checks = {}, {}

    def rec(op, ms):
        gbs = alg[op] / (ms * 1e-3) / 1e9
        ops[op] = {"ms": round(ms, 4), "gnnz_per_s": round(nnz / ms / 1e6, 3),
                   "alg_gb_per_s": round(gbs, 1), "roofline_frac": round(gbs / pk, 4)}

    ms, csr = timed(lambda: lib.coo_to_csr(n, n, row, col, vals), args.reps)
    rec("coo_to_csr", ms)
    rp, ccol, cval = csr
    checks["row_ptr_total"] = int(rp[-1]) == nnz and int(rp[0]) == 0
    checks["coo_to_csr_copies"] = bool((ccol == col).all()) and bool((cval == vals).all())
    checks["row_ptr_is_histogram"] = bool(
        (torch.bincount(row.to(torch.int64), minlength=n) == (rp[1:] - rp[:-1])).all())

    ms, csc = timed(lambda: lib.csr_to_csc(n, n, rp, ccol, cval), args.reps)
    rec("csr_to_csc", ms)
    cp, crow, cv = csc
    # the pattern is symmetric, so the CSC of A has the arrays of the CSR of A^T = A's pattern
    checks["csc_pattern_symmetric"] = bool((cp == rp).all()) and bool((crow == ccol).all())
    back = lib.csr_to_csc(n, n, cp, crow, cv)      # transpose of the transpose
    checks["transpose_twice_identity"] = all(bool((a == b).all()) for a, b in zip(back, csr))
    del back
    ms, csc2 = timed(lambda: lib.coo_to_csc(n, n, row, col, vals), max(1, args.reps // 2))
    rec("coo_to_csc", ms)
    checks["coo_to_csc_equals_csr_to_csc"] = all(bool((a == b).all()) for a, b in zip(csc2, csc))
    del csc2, csc, cp, crow, cv
    ms, coo = timed(lambda: lib.csr_to_coo(n, n, rp, ccol, cval), max(1, args.reps // 2))
    rec("csr_to_coo", ms)
    checks["csr_to_coo_row"] = bool((coo[0] == row).all())
    del coo

    ms, inv = timed(lambda: lib.degree_reorder(n, rp, True), args.reps)
    rec("degree_reorder", ms)
    deg = (rp[1:] - rp[:-1])
    order = lib.inverse_permutation(inv)           # order[new] = old
    checks["degree_perm_bijection"] = bool(
        (torch.sort(inv.to(torch.int64)).values == torch.arange(n, device=dev)).all())
    d_new = deg[order.to(torch.int64)]
    asc = bool((d_new[1:] >= d_new[:-1]).all())
    ties = d_new[1:] == d_new[:-1]
    tie_rule = bool((order[:-1][ties] > order[1:][ties]).all())   # equal degree: descending id
    checks["degree_order_monotone_ties_desc_id"] = asc and tie_rule
    ms, dist = timed(lambda: lib.degree_distribution(n, nnz, rp), args.reps)
    rec("degree_distribution", ms)
    checks["degree_distribution_sum"] = abs(float(dist.double().sum()) - 1.0) < 1e-3
    checks["degree_distribution_exact"] = bool(
        (dist == (deg.to(torch.float32) / torch.tensor(float(nnz), dtype=torch.float32,
                                                       device=dev))).all())
    ms, dg = timed(lambda: lib.degrees(n, rp), args.reps)
    rec("degrees", ms)
    checks["degrees"] = bool((dg == deg).all())
    del d_new, ties, dist, dg

    ms, p = timed(lambda: lib.permute2d(n, n, rp, ccol, cval, inv, inv), args.reps)
    rec("permute2d", ms)
    prp, pcol, pval = p
    checks["permute2d_row_ptr"] = bool(((prp[1:] - prp[:-1]) == deg[order.to(torch.int64)]).all())
    seg_sorted = pcol[1:] > pcol[:-1]
    starts = torch.zeros(nnz - 1, dtype=torch.bool, device=dev)
    b = prp[1:-1].to(torch.int64)
    b = b[(b > 0) & (b < nnz)]
    starts[b - 1] = True
    checks["permute2d_rows_sorted"] = bool((seg_sorted | starts).all())
    del seg_sorted, starts, b
    q = lib.permute2d(n, n, prp, pcol, pval, order, order)
    checks["permute2d_then_inverse_identity"] = all(bool((a == b).all()) for a, b in zip(q, csr))
    del q, p, prp, pcol, pval

    if args.rcm:
        ms, rinv = timed(lambda: lib.rcm_reorder(n, rp, ccol), 1)
        st = lib.rcm_last_stats()
        ops["rcm_reorder"] = {"ms": round(ms, 3), "levels_narrow": st["levels_narrow"],
                              "levels_wide": st["levels_wide"], "bfs": st["bfs"]}
        checks["rcm_perm_bijection"] = bool(
            (torch.sort(rinv.to(torch.int64)).values == torch.arange(n, device=dev)).all())
        bw0 = int((row.to(torch.int64) - col.to(torch.int64)).abs().max()) + 1
        r2 = rinv[row.to(torch.int64)].to(torch.int64)
        c2 = rinv[col.to(torch.int64)].to(torch.int64)
        bw1 = int((r2 - c2).abs().max()) + 1
        ops["rcm_reorder"]["bandwidth_before"] = bw0
        ops["rcm_reorder"]["bandwidth_after"] = bw1
        del r2, c2
        ms, _ = timed(lambda: lib.permute2d(n, n, rp, ccol, cval, rinv, rinv), 3)
        ops["permute2d_rcm"] = {"ms": round(ms, 4), "gnnz_per_s": round(nnz / ms / 1e6, 3),
                                "roofline_frac": round(alg["permute2d"] / (ms * 1e-3) / 1e9 / pk, 4)}

    line = {"config": name, "n": n, "nnz": nnz, "types": "IDType=NNZType=int32, ValueType=float32",
            "peak_gbs": pk, "ops": ops, "checks": checks, "all_checks_pass": all(checks.values()),
            "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
    s = json.dumps(line)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
