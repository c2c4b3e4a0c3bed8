#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid G]

Workload (BASELINE.json configs[1], "C2"): RCMReorder + Permute2D on a synthetic 2-D Poisson
5-point stencil, G x G grid (default 4096 -> 16.7 M rows, 83.9 M nnz), IDType=NNZType=int32,
ValueType=float32.  A "step" is one pass of the path over that matrix:
    inv = RCMReorder(csr);  out = Permute2D(inv, csr)
`value` is nonzeros per second (GNNZ/s) of the whole step with the CSR resident in HBM;
`e2e` is the same step through the C ABI with HOST (pinned) buffers: H2D of the CSR, the two
operators, D2H of the permuted CSR and of the permutation, all inside the timed region.
With N > 1 ranks (torchrun) every rank runs an independent replica of the workload (RCM does
not shard: its levels serialise -- "replicas only"), value = N * nnz / max-time.

--impl reference times the reference's own CPU implementation (oracle/_ref/libsbref.so = the
unmodified reference compiled header-only; the oracle port if that is absent) on the host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

ALG_BYTES_PERMUTE2D = lambda n, nnz: nnz * 16 + 2 * (n + 1) * 4 + 2 * n * 4  # noqa: E731
ALG_BYTES = {  # SURVEY.md section 8(d), I = N = V = 4 bytes
    "coo_to_csr": lambda n, nnz: nnz * 20 + (n + 1) * 4,
    "csr_to_csc": lambda n, nnz: nnz * 16 + 2 * (n + 1) * 4,
    "permute2d": ALG_BYTES_PERMUTE2D,
    "degree_reorder": lambda n, nnz: (n + 1) * 4 + n * 4,
    "degree_distribution": lambda n, nnz: (n + 1) * 4 + n * 4,
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the Permute2D kernels of this workload,
    per launch, from the committed `ncu --set full` capture (profiles/*_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r1_e_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get("permute2d_dram_bytes_per_launch")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


# ------------------------------------------------------------------------ reference arm
def cpu_reference_step(grid, reps=1):
    """Times the reference's CPU path (RCMReorder + Permute2D) on a grid x grid Poisson matrix.
    Returns dict(value GNNZ/s, seconds, kind, cores, sample)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib
    from sparsebase_b200 import synth

    ref = oracle_lib.reference()
    kind = "reference"
    if ref is None:
        ref, kind = oracle_lib.restated(), "port"
    n, rp, col, vals = synth.poisson2d(grid, grid)
    rp, col, vals = rp.numpy(), col.numpy(), vals.numpy()
    nnz = len(col)
    best, parts = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        inv = ref.rcm_reorder(n, rp, col, vals)
        t1 = time.perf_counter()
        out = ref.permute2d(n, n, rp, col, vals, inv, inv)
        t2 = time.perf_counter()
        if best is None or t2 - t0 < best:
            best, parts = t2 - t0, (t1 - t0, t2 - t1)
        del out
    cores = os.cpu_count() if kind == "reference" else 1
    return {"value": nnz / best / 1e9, "unit": "GNNZ/s", "seconds": best, "rcm_s": parts[0],
            "permute2d_s": parts[1], "kind": kind, "cores": cores, "nnz": nnz, "n": n,
            "sample": f"RCMReorder+Permute2D on Poisson {grid}x{grid} ({n} rows, {nnz} nnz), "
                      f"best of {reps}; OpenMP threads = {cores} (only the CSR-ctor row sort is "
                      "parallel in the reference)"}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    total = args.steps + args.warmup
    grid = min(args.grid, 2048 if total <= 12 else 1024)
    for _ in range(args.warmup):
        cpu_reference_step(min(grid, 512))
    t = []
    r = None
    for _ in range(args.steps):
        r = cpu_reference_step(grid)
        t.append(r["seconds"])
    sec = sum(t) / len(t)
    value = r["nnz"] / sec / 1e9
    line = {
        "impl": "reference", "metric": "rcm_permute2d_gnnz_per_s", "value": value,
        "unit": "GNNZ/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"C2: RCMReorder+Permute2D, 2-D Poisson 5-point {args.grid}x"
                               f"{args.grid}; reference arm runs the bounded sample below"},
        "cpu_baseline": {"value": value, "unit": "GNNZ/s", "cores": r["cores"],
                         "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": value, "unit": "GNNZ/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sb200")
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-grid", type=int, default=2048)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from sparsebase_b200 import lib, synth
    lib.load()
    dev = torch.device("cuda", local_rank)
    W = max(args.warmup, 3)

    n, row_ptr, col, vals = synth.poisson2d(args.grid, args.grid, device=dev)
    nnz = col.numel()
    out = (torch.empty_like(row_ptr), torch.empty_like(col), torch.empty_like(vals))

    def step():
        inv = lib.rcm_reorder(n, row_ptr, col)
        lib.permute2d(n, n, row_ptr, col, vals, inv, inv, out=out)
        return inv

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    # ---- device-resident timing: K steps, CUDA events, per-operator split ----
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    lib.reset_launch_count()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record()
        inv = lib.rcm_reorder(n, row_ptr, col)
        ev[k][1].record()
        lib.permute2d(n, n, row_ptr, col, vals, inv, inv, out=out)
        ev[k][2].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.launch_count()
    clocks = sampler.stop()
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    rcm_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    p2d_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    rcm_stats = lib.rcm_last_stats()
    ms_per_step = total_ms / args.steps
    if world > 1:
        t = torch.tensor([ms_per_step], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())
    value = world * nnz / (ms_per_step * 1e-3) / 1e9

    # ---- end to end: host (pinned) buffers in, host buffers out ----
    h_in = [t.cpu().pin_memory() for t in (row_ptr, col, vals)]
    h_out = [torch.empty_like(t).pin_memory() for t in h_in]
    h_inv = torch.empty(n, dtype=torch.int32).pin_memory()
    d_in = [torch.empty_like(t) for t in (row_ptr, col, vals)]
    h2d = sum(t.numel() * t.element_size() for t in h_in)
    d2h = sum(t.numel() * t.element_size() for t in h_out) + h_inv.numel() * 4

    side = torch.cuda.Stream(device=dev)

    def e2e_step():
        # RCM needs the pattern only: the values travel on a second stream while it runs, and
        # the permutation goes back to the host while Permute2D runs
        main = torch.cuda.current_stream(dev)
        d_in[0].copy_(h_in[0], non_blocking=True)
        d_in[1].copy_(h_in[1], non_blocking=True)
        side.wait_stream(main)          # previous step's readers of d_in[2] are done
        with torch.cuda.stream(side):
            d_in[2].copy_(h_in[2], non_blocking=True)
        inv = lib.rcm_reorder(n, d_in[0], d_in[1])
        main.wait_stream(side)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            h_inv.copy_(inv, non_blocking=True)
        lib.permute2d(n, n, d_in[0], d_in[1], d_in[2], inv, inv, out=out)
        for h, d in zip(h_out, out):
            h.copy_(d, non_blocking=True)
        main.wait_stream(side)

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(2, min(args.steps, 5))
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * nnz / (e2e_ms * 1e-3) / 1e9

    # ---- other operators of the path on the same matrix (reported, not the headline) ----
    peak, peak_src = peaks()
    ops = {"rcm_reorder": {"ms": rcm_ms, "levels_narrow": rcm_stats["levels_narrow"],
                           "levels_wide": rcm_stats["levels_wide"], "bfs": rcm_stats["bfs"],
                           "phase_cycles": rcm_stats["phase_cycles"]}}

    def time_op(name, fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        gbs = ALG_BYTES[name](n, nnz) / (ms * 1e-3) / 1e9
        ops[name] = {"ms": ms, "gnnz_per_s": nnz / (ms * 1e-3) / 1e9, "alg_gb_per_s": gbs,
                     "roofline_frac": gbs / peak}

    if rank == 0:
        row = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int32),
                                      (row_ptr[1:] - row_ptr[:-1]).to(torch.int64))
        inv_fixed = inv
        time_op("permute2d", lambda: lib.permute2d(n, n, row_ptr, col, vals, inv_fixed, inv_fixed,
                                                   out=out))
        time_op("coo_to_csr", lambda: lib.coo_to_csr(n, n, row, col, vals))
        time_op("csr_to_csc", lambda: lib.csr_to_csc(n, n, row_ptr, col, vals))
        time_op("degree_reorder", lambda: lib.degree_reorder(n, row_ptr, True))
        time_op("degree_distribution", lambda: lib.degree_distribution(n, nnz, row_ptr))
        del row

    # ---- N > 1: the row-block sharded operators on the same matrix (strong scaling: the C2
    #      matrix split into `world` nnz-balanced row blocks; max time over ranks, exchanges
    #      included; aggregate roofline = world x per-GPU peak) ----
    ops_sharded = None
    if world > 1:
        from sparsebase_b200 import sharded
        bounds = lib.partition_rows(n, nnz, row_ptr, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        a, b = int(row_ptr[lo]), int(row_ptr[hi])
        deg_l = (row_ptr[lo + 1:hi + 1] - row_ptr[lo:hi]).to(torch.int64)
        row_l = torch.repeat_interleave(torch.arange(lo, hi, device=dev, dtype=torch.int32), deg_l)
        col_l, val_l = col[a:b].contiguous(), vals[a:b].contiguous()
        ops_sharded = {}

        def time_sharded(name, fn, reps=3):
            fn()
            barrier()
            e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_a.record()
            for _ in range(reps):
                fn()
            e_b.record()
            barrier()
            t = torch.tensor([e_a.elapsed_time(e_b) / reps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            gbs = ALG_BYTES[name](n, nnz) / (ms * 1e-3) / 1e9
            ops_sharded[name] = {"ms": ms, "gnnz_per_s": nnz / (ms * 1e-3) / 1e9,
                                 "alg_gb_per_s": gbs, "roofline_frac": gbs / (peak * world)}

        shard = None
        try:  # a failure here must not cost the headline line (the error is reported instead)
            shard = sharded.coo_to_csr(lib, n, n, bounds, row_l.clone(), col_l.clone(),
                                       val_l.clone())
            time_sharded("coo_to_csr", lambda: sharded.coo_to_csr(lib, n, n, bounds, row_l, col_l,
                                                                  val_l, copy=False))
            time_sharded("csr_to_csc", lambda: sharded.csr_to_csc(lib, shard))
            time_sharded("permute2d", lambda: sharded.permute2d(lib, shard, inv, inv))
            time_sharded("degree_reorder", lambda: sharded.degree_reorder(lib, shard, True))
            time_sharded("degree_distribution", lambda: sharded.degree_distribution(lib, shard))
        except Exception as exc:  # noqa: BLE001
            ops_sharded["error"] = f"{type(exc).__name__}: {exc}"[:300]
        del row_l, col_l, val_l, shard

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    p2d_bytes = ALG_BYTES_PERMUTE2D(n, nnz)
    achieved = p2d_bytes / (p2d_ms * 1e-3) / 1e9
    line = {
        "metric": "rcm_permute2d_gnnz_per_s", "value": value, "unit": "GNNZ/s",
        "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": f"C2: RCMReorder+Permute2D on 2-D Poisson 5-point {args.grid}x"
                               f"{args.grid} ({n} rows, {nnz} nnz), IDType=NNZType=int32, "
                               "ValueType=float32",
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} independent replicas (RCM does not shard)",
                   "l2_policy": "inputs (1.0 GB CSR) larger than the 126 MB L2; no flush needed"},
        "rcm_ms": rcm_ms, "permute2d_ms": p2d_ms, "wall_ms_per_step": t_wall * 1e3 / args.steps,
        "e2e": {"value": e2e_value, "unit": "GNNZ/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "permute_short_rows_kernel (fused gather / "
                     "renumber / row-sort of Permute2D) + permute_prepare_kernel + "
                     "scan_lookback_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": p2d_bytes,
                     "note": "RCM (rcm_narrow_kernel) is latency-bound level walking and is "
                             "reported in ms (rcm_ms), not against the HBM roofline"},
        "ops": ops,
    }
    if ops_sharded is not None:
        line["ops_sharded"] = ops_sharded
        line["ops_sharded_note"] = ("row-block sharded operators on the same C2 matrix split over "
                                    f"{world} GPUs (strong scaling), exchanges included, max over "
                                    "ranks; roofline_frac is against world x per-GPU peak")
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_step(args.cpu_grid)
        line["cpu_baseline"] = {"value": cb["value"], "unit": "GNNZ/s", "cores": cb["cores"],
                                "kind": cb["kind"], "sample": cb["sample"],
                                "rcm_s": cb["rcm_s"], "permute2d_s": cb["permute2d_s"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
