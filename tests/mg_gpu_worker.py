"""torchrun worker: the peer-memory multi-GPU operators (sparsebase_b200.mg -> sb200_mg_*),
compared bit-for-bit with the single-GPU operators run on the whole matrix by the same rank,
then (--time) timed next to the collective-based implementation (sparsebase_b200.sharded).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 tests/mg_gpu_worker.py [--graph rmat|er] [--scale 16] [--time]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsebase_b200 import lib, mg, sharded, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=15)
    ap.add_argument("--graph", default="rmat")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--out", default=None)
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

    g_rp, g_col, g_val = lib.coo_to_csr(n, n, row, col, vals)
    bounds = lib.partition_rows(n, nnz, g_rp, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    a, b = int(g_rp[lo]), int(g_rp[hi])
    # window: the largest exchange is a shard's (col, val) + the n+1 degree / pointer table, or
    # world col_ptr tables for the transpose
    shard = (nnz // world + n) * 8
    window = 2 * shard + (world + 2) * (n + 1) * 4 + (64 << 20)
    comm = mg.Comm(window)
    assert comm.allgather_i64(rank * 10 + 1) == [r * 10 + 1 for r in range(world)]

    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)
    p = torch.randperm(b - a, generator=g, device=dev)
    s = mg.coo_to_csr(comm, n, n, bounds, row[a:b][p], col[a:b][p], vals[a:b][p])
    assert s.nnz == nnz and s.nnz_base == a, (s.nnz, nnz, s.nnz_base, a)
    assert eq(s.row_ptr, g_rp[lo:hi + 1] - a), "mg COO->CSR row_ptr"
    assert eq(s.col, g_col[a:b]) and eq(s.vals, g_val[a:b]), "mg COO->CSR col/vals"

    for asc in (True, False):
        assert eq(mg.degree_reorder(comm, s, asc), lib.degree_reorder(n, g_rp, asc)), asc
    inv = mg.degree_reorder(comm, s, True)

    if not args.no_check:
        ps = mg.permute2d(comm, s, inv, inv)
        e_rp, e_col, e_val = lib.permute2d(n, n, g_rp, g_col, g_val, inv, inv)
        nlo, nhi = ps.bounds[rank], ps.bounds[rank + 1]
        a2, b2 = int(e_rp[nlo]), int(e_rp[nhi])
        assert ps.nnz_base == a2, (ps.nnz_base, a2)
        assert eq(ps.row_ptr + a2, e_rp[nlo:nhi + 1]), "mg Permute2D row_ptr"
        assert eq(ps.col, e_col[a2:b2]) and eq(ps.vals, e_val[a2:b2]), "mg Permute2D col/vals"
        # row order only / identity
        ps2 = mg.permute2d(comm, s, inv, None)
        e2 = lib.permute2d(n, n, g_rp, g_col, g_val, inv, None)
        a4, b4 = int(e2[0][ps2.bounds[rank]]), int(e2[0][ps2.bounds[rank + 1]])
        assert eq(ps2.col, e2[1][a4:b4]) and eq(ps2.vals, e2[2][a4:b4]), "mg Permute2D rows only"
        del ps, ps2, e_rp, e_col, e_val, e2

        cs = mg.csr_to_csc(comm, s)
        c_cp, c_row, c_val = lib.csr_to_csc(n, n, g_rp, g_col, g_val)
        clo, chi = cs.bounds[rank], cs.bounds[rank + 1]
        a3, b3 = int(c_cp[clo]), int(c_cp[chi])
        assert cs.nnz_base == a3
        assert eq(cs.col_ptr + a3, c_cp[clo:chi + 1]), "mg CSR->CSC col_ptr"
        assert eq(cs.row, c_row[a3:b3]) and eq(cs.vals, c_val[a3:b3]), "mg CSR->CSC row/vals"
        del cs, c_cp, c_row, c_val

        # Permute1D: a vertex-valued array permuted by inv, block by block
        eb = sharded.even_bounds(n, world)
        x = synth.hash_vals(n, seed=9, device=dev)
        out = mg.permute1d(comm, eb, x[eb[rank]:eb[rank + 1]].contiguous(),
                           inv[eb[rank]:eb[rank + 1]].contiguous())
        assert eq(out, lib.permute1d(x, inv)[eb[rank]:eb[rank + 1]]), "mg Permute1D"
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"MG OK world={world} n={n} nnz={nnz}", flush=True)

    if args.time:
        def timed(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return round(float(t.item()), 4)

        r_l, c_l, v_l = row[a:b].contiguous(), col[a:b].contiguous(), vals[a:b].contiguous()
        res = {"world": world, "graph": args.graph, "scale": args.scale, "n": n, "nnz": nnz}
        res["single_gpu"] = {
            "coo_to_csr": timed(lambda: lib.coo_to_csr(n, n, row, col, vals)),
            "degree_reorder": timed(lambda: lib.degree_reorder(n, g_rp, True)),
            "permute2d": timed(lambda: lib.permute2d(n, n, g_rp, g_col, g_val, inv, inv)),
            "csr_to_csc": timed(lambda: lib.csr_to_csc(n, n, g_rp, g_col, g_val)),
        }
        res["peer_memory"] = {
            "coo_to_csr": timed(lambda: mg.coo_to_csr(comm, n, n, bounds, r_l, c_l, v_l, copy=False)),
            "degree_reorder": timed(lambda: mg.degree_reorder(comm, s, True)),
            "permute2d": timed(lambda: mg.permute2d(comm, s, inv, inv)),
            "csr_to_csc": timed(lambda: mg.csr_to_csc(comm, s)),
        }
        s2 = sharded.coo_to_csr(lib, n, n, bounds, r_l, c_l, v_l)
        res["collectives"] = {
            "coo_to_csr": timed(lambda: sharded.coo_to_csr(lib, n, n, bounds, r_l, c_l, v_l,
                                                           copy=False)),
            "degree_reorder": timed(lambda: sharded.degree_reorder(lib, s2, True)),
            "permute2d": timed(lambda: sharded.permute2d(lib, s2, inv, inv)),
            "csr_to_csc": timed(lambda: sharded.csr_to_csc(lib, s2)),
        }
        os.environ["SB200_MG_TRACE"] = "1"     # stage times of one call each (rank 0, stderr)
        mg.permute2d(comm, s, inv, inv)
        mg.csr_to_csc(comm, s)
        torch.cuda.synchronize()
        os.environ.pop("SB200_MG_TRACE")
        if rank == 0:
            print("MG_TIMES " + json.dumps(res), flush=True)
            if args.out:
                with open(args.out, "w") as f:
                    json.dump(res, f, indent=1)
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
