#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: the task statement / DESIGN.md 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale S]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], "C4"): COO->CSR + DegreeReorder + Permute2D on a power-law
R-MAT scale-26 graph (67.1 M vertices, ~1.07 B nnz, IDType = NNZType = int32, ValueType =
float32), row-block sharded over the N GPUs.  A "step" is one pass of that path over the matrix:

    csr = COO->CSR(coo);  inv = DegreeReorder(csr);  out = Permute2D(inv, csr)

At N = 1 these are the single-GPU operators; at N > 1 the peer-memory operators (sb200_mg_*:
every rank owns an nnz-balanced row block and stores its part of every exchange straight into
the destination GPU's window over NVLink).  The matrix is the same at every N (strong scaling);
`value` = nnz / (step time, max over ranks) with the shards resident in HBM; `e2e` is the same
step with HOST (pinned) shards in and out.  Beside the headline the line carries
  ops      per-operator ms / GNNZ/s / fraction of the HBM roofline (x N GPUs)
  c3       configs[2]: CSR->CSC + DegreeDistribution on Erdos-Renyi 2^24 x 16 (268 M nnz)
  rcm      configs[1] (N = 1): RCMReorder + Permute2D on 2-D Poisson 4096^2 in ms, checked
           byte for byte against the compiled reference; bandwidth before / after
  parity   the headline result checked against the oracle (sampled rows + checksums)

--impl reference times the reference's own CPU implementation of the same path
(oracle/_ref/libsbref.so = the unmodified reference compiled header-only; the C restatement if
that is absent) on a bounded R-MAT sample of the workload, with all host threads.
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


def _set_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference's OpenMP regions (the constructor row
    sort) get every host core, and the line says how many."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


HOST_THREADS = _set_host_threads()

import torch  # noqa: E402

WORKLOAD = ("C4: COO->CSR + DegreeReorder + Permute2D on R-MAT scale {scale} (edge factor 8 "
            "undirected, a,b,c = 0.57,0.19,0.19, seed 44), IDType=NNZType=int32, ValueType=float32, "
            "nnz-balanced row blocks over the GPUs")
METRIC = "coo_csr_degreereorder_permute2d_gnnz_per_s"

ALG_BYTES = {  # SURVEY.md section 8(d), I = N = V = 4 bytes
    "coo_to_csr": lambda n, nnz: nnz * 20 + (n + 1) * 4,
    "csr_to_csc": lambda n, nnz: nnz * 16 + 2 * (n + 1) * 4,
    "permute2d": lambda n, nnz: nnz * 16 + 2 * (n + 1) * 4 + 2 * n * 4,
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
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the
    committed `ncu --set full` capture (profiles/r2_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get("permute2d_gather_dram_bytes_per_launch")


_POLLER = r"""
import sys, time
import pynvml as nv
idx = [int(x) for x in sys.argv[1].split(',')]
period = float(sys.argv[2]) / 1000.0
nv.nvmlInit()
hs = [nv.nvmlDeviceGetHandleByIndex(i) for i in idx]
mx = [nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM) for h in hs]
print('max', *mx, flush=True)
while True:
    t = time.time()
    for i, h in zip(idx, hs):
        try:
            print(t, i, nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                  nv.nvmlDeviceGetCurrentClocksThrottleReasons(h), flush=True)
        except Exception:
            pass
    time.sleep(period)
"""


class ClockSampler:
    """SM clocks / throttle reasons of all the job's GPUs sampled DURING the timed region by ONE
    poller: a child process of rank 0 calling NVML every SB200_BENCH_CLOCK_MS (50) milliseconds;
    samples carry wall-clock stamps and the timed region is cut out afterwards.  What was tried
    and dropped: one `nvidia-smi -lms 20` per rank (eight concurrent pollers stalled the CUDA
    calls of the 8-GPU run for milliseconds at a time: DegreeReorder 13 ms between the events,
    2.4 ms inside the operator); an NVML thread inside rank 0 (10 ms period: DegreeReorder 0.9 ->
    5.3 ms at N = 1)."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
               "sw_power_cap": 0x4}

    def __init__(self, indices, enabled=True):
        self.indices = list(indices)
        self.enabled = enabled
        self.period_ms = int(os.environ.get("SB200_BENCH_CLOCK_MS", "50"))
        self.rows, self.mx = [], []
        self.proc, self.thread = None, None

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(
                [sys.executable, "-c", _POLLER, ",".join(map(str, self.indices)),
                 str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = line.split()
            try:
                if f[0] == "max":
                    self.mx = [float(x) for x in f[1:]]
                else:
                    self.rows.append((float(f[0]), int(f[1]), float(f[2]), int(f[3])))
            except (ValueError, IndexError):
                continue

    def stop(self, t0=None, t1=None):
        """t0, t1 = time.time() at the two ends of the timed region."""
        if not self.enabled:
            return None
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML samples"]}
        timed = [r for r in self.rows if t0 is not None and t0 <= r[0] <= t1]
        use = timed if timed else self.rows
        sm = sorted(r[2] for r in use)
        bits = 0
        for r in use:
            bits |= r[3]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0],
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "samples": len(use), "samples_in_timed_region": len(timed),
                "samples_total": len(self.rows), "gpus": self.indices,
                "period_ms": self.period_ms, "source": "NVML poller process of rank 0",
                "reasons": sorted(k for k, v in self.REASONS.items() if bits & v)}


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


# ------------------------------------------------------------------------ CPU arms
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    ref = oracle_lib.reference()
    if ref is not None:
        return ref, "reference", oracle_lib
    return oracle_lib.restated(), "port", oracle_lib


def cpu_pipeline_step(scale, graph=None):
    """The reference's CPU path for the headline workload on an R-MAT sample of `scale`:
    COO->CSR (converter_order_two.cc:162-212), DegreeReorder (degree_reorder.cc:22-62),
    Permute2D (permute_order_two.cc:21-79 + the CSR constructor sort)."""
    from sparsebase_b200 import synth
    ref, kind, _ = _oracle()
    if graph is None:
        n, row, col = synth.rmat(scale, 8, seed=44)
        vals = synth.hash_vals(col.numel(), seed=7)
        graph = (n, row.numpy(), col.numpy(), vals.numpy())
    n, row, col, vals = graph
    nnz = len(col)
    t0 = time.perf_counter()
    rp, cc, vv = ref.coo_to_csr(n, n, row, col, vals)
    t1 = time.perf_counter()
    inv = ref.degree_reorder(n, rp, cc, True, vv)
    t2 = time.perf_counter()
    out = ref.permute2d(n, n, rp, cc, vv, inv, inv)
    t3 = time.perf_counter()
    del out
    cores = HOST_THREADS if kind == "reference" else 1
    return {"value": nnz / (t3 - t0) / 1e9, "unit": "GNNZ/s", "seconds": t3 - t0,
            "coo_to_csr_s": t1 - t0, "degree_reorder_s": t2 - t1, "permute2d_s": t3 - t2,
            "kind": kind, "cores": cores, "nnz": nnz, "n": n,
            "sample": f"the same path on R-MAT scale {scale} ({n} rows, {nnz} nnz; same generator "
                      f"and seed as the workload); OpenMP threads = {cores} (set explicitly; only "
                      "the CSR-constructor row check / sort is parallel in the reference); the "
                      "harness replaces operator new[] by calloc + 64 B (reference heap overflow "
                      "in degree_reorder.cc:41-45), which is also the timed allocator"}, graph


def run_reference(args):
    rank, _, _ = dist_env()
    if rank != 0:
        return
    total = args.steps + args.warmup
    scale = args.ref_scale if total <= 10 else max(16, args.ref_scale - 1)
    for _ in range(args.warmup):
        cpu_pipeline_step(min(scale, 16))
    graph, t, r = None, [], None
    for _ in range(args.steps):
        r, graph = cpu_pipeline_step(scale, graph)
        t.append(r["seconds"])
    sec = sum(t) / len(t)
    value = r["nnz"] / sec / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GNNZ/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(scale=args.scale)},
        "cpu_baseline": {"value": value, "unit": "GNNZ/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"], "coo_to_csr_s": r["coo_to_csr_s"],
                         "degree_reorder_s": r["degree_reorder_s"], "permute2d_s": r["permute2d_s"]},
        "e2e": {"value": value, "unit": "GNNZ/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------ parity helpers
def check_headline(lib, n, g_rp, g_col, g_val, inv, out_rows, out_shard, nsamp=500):
    """The headline result against the oracle: DegreeReorder's order and tie rule over the whole
    permutation, Permute2D's row_ptr / checksums over the rank's whole block, and `nsamp` rows of
    the block against the C restatement run on a matrix made of exactly those rows.
    out_rows = (nlo, nhi); out_shard = (row_ptr block-local, col, vals)."""
    import ctypes
    import numpy as np
    _, _, oracle_lib = _oracle()
    orc = oracle_lib.restated()
    dev = g_rp.device
    res = {}
    deg = (g_rp[1:] - g_rp[:-1]).to(torch.int64)
    order = torch.empty(n, dtype=torch.int64, device=dev)
    order[inv.to(torch.int64)] = torch.arange(n, device=dev)
    res["inv_is_permutation"] = bool((torch.bincount(inv.to(torch.int64), minlength=n) == 1).all())
    d_sorted = deg[order]
    tie = d_sorted[1:] == d_sorted[:-1]
    res["degree_order_and_tie_rule"] = bool((d_sorted[1:] >= d_sorted[:-1]).all()) and \
        bool((order[1:][tie] < order[:-1][tie]).all())
    nlo, nhi = out_rows
    orp, ocol, oval = out_shard
    res["row_ptr"] = bool(((orp[1:] - orp[:-1]).to(torch.int64) == d_sorted[nlo:nhi]).all())
    del d_sorted, tie
    # sampled rows of my block
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    new_rows = (torch.randint(0, max(1, nhi - nlo), (nsamp,), generator=g, device=dev) + nlo).unique()
    old_rows = order[new_rows]
    h_old, h_new = old_rows.cpu().numpy(), new_rows.cpu().numpy()
    h_rp = g_rp.cpu().numpy()
    h_orp = orp.cpu().numpy()
    h_inv = inv.cpu().numpy()
    sub_rp = np.zeros(len(h_old) + 1, dtype=h_rp.dtype)
    pc, pv = [], []
    for k, r in enumerate(h_old):
        a, b = int(h_rp[r]), int(h_rp[r + 1])
        pc.append(g_col[a:b].cpu().numpy())
        pv.append(g_val[a:b].cpu().numpy())
        sub_rp[k + 1] = sub_rp[k] + (b - a)
    sub_col, sub_val = np.concatenate(pc), np.concatenate(pv)
    ident = np.arange(len(h_old), dtype=h_inv.dtype)
    t = oracle_lib.tag_of(sub_col.dtype, sub_rp.dtype, sub_val.dtype)
    e_rp = np.empty(len(h_old) + 1, sub_rp.dtype)
    e_col, e_val = np.empty_like(sub_col), np.empty_like(sub_val)
    P = oracle_lib._ptr
    rc = orc._fn(f"permute2d_{t}")(ctypes.c_int64(len(h_old)), ctypes.c_int64(n), P(sub_rp),
                                   P(sub_col), P(sub_val), P(ident), P(h_inv), P(e_rp), P(e_col),
                                   P(e_val))
    ok = rc == 0
    for k, j in enumerate(h_new):
        a, b = int(h_orp[j - nlo]), int(h_orp[j - nlo + 1])
        ea, eb = int(e_rp[k]), int(e_rp[k + 1])
        ok = ok and (b - a == eb - ea) and np.array_equal(ocol[a:b].cpu().numpy(), e_col[ea:eb]) \
            and np.array_equal(oval[a:b].cpu().numpy().view(np.uint32), e_val[ea:eb].view(np.uint32))
    res["sampled_rows_vs_oracle"] = bool(ok)
    res["rows_sampled"] = int(len(h_new))
    return res


def rcm_block(lib, dev, peak, grid=4096, reps=3):
    """configs[1]: RCMReorder + Permute2D on the 2-D Poisson grid, with the permutation and the
    permuted matrix compared byte for byte with the reference's CPU result."""
    import numpy as np
    from sparsebase_b200 import synth
    n, rp, col, vals = synth.poisson2d(grid, grid, device=dev)
    nnz = col.numel()
    out = (torch.empty_like(rp), torch.empty_like(col), torch.empty_like(vals))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
    inv = lib.rcm_reorder(n, rp, col)
    lib.permute2d(n, n, rp, col, vals, inv, inv, out=out)
    torch.cuda.synchronize()
    for k in range(reps):
        ev[k][0].record()
        inv = lib.rcm_reorder(n, rp, col)
        ev[k][1].record()
        lib.permute2d(n, n, rp, col, vals, inv, inv, out=out)
        ev[k][2].record()
    torch.cuda.synchronize()
    rcm_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / reps
    p2d_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / reps
    st = lib.rcm_last_stats()
    _, _, before = lib.degree_features(n, nnz, rp, col, want_arrays=False)
    _, _, after = lib.degree_features(n, nnz, out[0], out[1], want_arrays=False)
    # ReorderHeatmap (8 x 8) of the matrix as it is and as the RCM permutation arranges it
    hb = 8
    heat0 = lib.reorder_heatmap(n, n, rp, col, None, None, hb)
    hev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    hev[0].record()
    heat1 = lib.reorder_heatmap(n, n, rp, col, inv, inv, hb)
    hev[1].record()
    torch.cuda.synchronize()
    heatmap = {"num_parts": hb, "ms": hev[0].elapsed_time(hev[1]),
               "diagonal_share_before": float(heat0.view(hb, hb).diagonal().sum()),
               "diagonal_share_after": float(heat1.view(hb, hb).diagonal().sum())}
    ref, kind, _ = _oracle()
    h = [t.cpu().numpy() for t in (rp, col, vals)]
    t0 = time.perf_counter()
    e_inv = ref.rcm_reorder(n, h[0], h[1], h[2])
    t1 = time.perf_counter()
    e_out = ref.permute2d(n, n, h[0], h[1], h[2], e_inv, e_inv)
    t2 = time.perf_counter()
    same = np.array_equal(inv.cpu().numpy(), e_inv) and all(
        np.array_equal(a.cpu().numpy().view(np.uint8), b.view(np.uint8)) for a, b in zip(out, e_out))
    p2d_gbs = ALG_BYTES["permute2d"](n, nnz) / (p2d_ms * 1e-3) / 1e9
    # the same step as a SparseBase user gets it: host format::CSR in, host arrays out, through
    # the UNMODIFIED reference's dispatch with the sb200 plugin registered (oracle/plugin_bench.cc)
    e2e_plugin = None
    exe = os.path.join(ROOT, "oracle", "_ref", "plugin_bench")
    if os.path.exists(exe):
        del out, rp, col, vals
        torch.cuda.empty_cache()
        lib.trim()
        try:
            r = subprocess.run([exe, str(grid), "3"], capture_output=True, text=True, timeout=600)
            e2e_plugin = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else \
                {"error": (r.stdout + r.stderr)[-300:]}
        except Exception as exc:  # noqa: BLE001
            e2e_plugin = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    return {"e2e_plugin": e2e_plugin,
            "workload": f"C2: RCMReorder + Permute2D on 2-D Poisson 5-point {grid}x{grid} "
                        f"({n} rows, {nnz} nnz)",
            "rcm_ms": rcm_ms, "permute2d_ms": p2d_ms, "gnnz_per_s": nnz / ((rcm_ms + p2d_ms) * 1e-3) / 1e9,
            "permute2d_roofline_frac": p2d_gbs / peak,
            "levels": st["levels_narrow"] + st["levels_wide"], "levels_wide": st["levels_wide"],
            "us_per_level": rcm_ms * 1e3 / max(1, st["levels_narrow"] + st["levels_wide"]),
            "bfs": st["bfs"], "cluster_resizes": st["resizes"], "share_resplits": st["resplits"],
            "bandwidth_before": before["bandwidth"], "bandwidth_after": after["bandwidth"],
            "profile_before": before["profile"], "profile_after": after["profile"],
            "heatmap": heatmap,
            "parity_vs_reference": bool(same), "parity_checker": kind,
            "cpu": {"rcm_s": t1 - t0, "permute2d_s": t2 - t1, "kind": kind, "cores": HOST_THREADS,
                    "gnnz_per_s": nnz / (t2 - t0) / 1e9}}


# ------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sb200")
    ap.add_argument("--scale", type=int, default=26)
    ap.add_argument("--ref-scale", type=int, default=21)
    ap.add_argument("--cpu-scale", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rcm", action="store_true")
    ap.add_argument("--no-c3", action="store_true")
    ap.add_argument("--c3-log2n", type=int, default=24)
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
    from sparsebase_b200 import lib, mg, sharded, synth
    lib.load()
    dev = torch.device("cuda", local_rank)
    W = max(args.warmup, 3)
    peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # (the clock poller is started here so that it is up and sampling long before the warm-up)
    sampler = ClockSampler(range(world), enabled=(rank == 0))
    sampler.start()
    # ---- the matrix (identical on every rank: same generator, same seed, same hardware)
    n, row, col = synth.rmat(args.scale, 8, seed=44, device=dev)
    nnz = col.numel()
    vals = synth.hash_vals(nnz, seed=7, device=dev)
    torch.cuda.empty_cache()
    g_rp, g_col, g_val = lib.coo_to_csr(n, n, row, col, vals)   # set-up: blocks + parity checks
    bounds = lib.partition_rows(n, nnz, g_rp, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    a, b = int(g_rp[lo]), int(g_rp[hi])
    if world > 1:
        r_l, c_l, v_l = row[a:b].clone(), col[a:b].clone(), vals[a:b].clone()
        del row, col, vals
        torch.cuda.empty_cache()
        shard_bytes = (nnz // world + n // world + 4096) * 8
        window = int(shard_bytes * 1.15) + (n + 1) * 4 + (256 << 20)
        comm = mg.Comm(window)
    else:
        r_l, c_l, v_l = row, col, vals
        comm = None

    ev = None

    def step(record=None):
        """-> (inv, (nlo, nhi), (row_ptr', col', vals') of this rank's block)"""
        if record:
            record[0].record()
        if world == 1:
            csr = lib.coo_to_csr(n, n, r_l, c_l, v_l)
            if record:
                record[1].record()
            inv = lib.degree_reorder(n, csr[0], True)
            if record:
                record[2].record()
            out = lib.permute2d(n, n, csr[0], csr[1], csr[2], inv, inv)
            if record:
                record[3].record()
            return inv, (0, n), out
        s = mg.coo_to_csr(comm, n, n, bounds, r_l, c_l, v_l, presorted=True)
        if record:
            record[1].record()
        inv = mg.degree_reorder(comm, s, True)
        if record:
            record[2].record()
        p = mg.permute2d(comm, s, inv, inv)
        if record:
            record[3].record()
        return inv, (p.bounds[rank], p.bounds[rank + 1]), (p.row_ptr, p.col, p.vals)

    # (clocks: samples stamped inside the timed region are reported; when that region is shorter
    # than the polling period -- 8 GPUs -- all samples since start-up are used and the line
    # says so in samples_in_timed_region)
    for _ in range(W):
        step()
    # ---- device-resident timing: K steps, CUDA events on the launching stream
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier()
    lib.reset_launch_count()
    t_clock0 = time.time()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        inv, out_rows, out = step(ev[k])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.launch_count()
    clocks = sampler.stop(t_clock0, time.time())
    ms_per_step = max_over_ranks(ev[0][0].elapsed_time(ev[-1][3]) / args.steps)
    op_ms = {name: max_over_ranks(sum(e[i].elapsed_time(e[i + 1]) for e in ev) / args.steps)
             for i, name in enumerate(("coo_to_csr", "degree_reorder", "permute2d"))}
    value = nnz / (ms_per_step * 1e-3) / 1e9

    # ---- parity of the headline result (outside the timed region)
    parity = check_headline(lib, n, g_rp, g_col, g_val, inv, out_rows, out)
    parity_ok = all(v for k, v in parity.items() if isinstance(v, bool))
    if world > 1:
        t = torch.tensor([1.0 if parity_ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        parity_ok = bool(t.item() > 0.5)
    out_sizes = [(t.numel(), t.dtype) for t in out]   # (the same every step)
    del out

    # ---- end to end: host (pinned) shards in, host shards out
    h_in = [t.cpu().pin_memory() for t in (r_l, c_l, v_l)]
    d_in = [torch.empty_like(t) for t in (r_l, c_l, v_l)]
    h2d = sum(t.numel() * t.element_size() for t in h_in)
    h_inv = torch.empty(n, dtype=torch.int32).pin_memory() if rank == 0 else None
    h_out = [torch.empty(cnt, dtype=dt).pin_memory() for cnt, dt in out_sizes]
    d2h_box = [0]

    def e2e_step():
        nonlocal r_l, c_l, v_l
        for d, h in zip(d_in, h_in):
            d.copy_(h, non_blocking=True)
        keep = (r_l, c_l, v_l)
        r_l, c_l, v_l = d_in
        inv2, _, o = step()
        r_l, c_l, v_l = keep
        bytes_out = 0
        for h, d in zip(h_out, o):
            h[: d.numel()].copy_(d, non_blocking=True)
            bytes_out += d.numel() * d.element_size()
        if h_inv is not None:
            h_inv.copy_(inv2, non_blocking=True)
            bytes_out += n * 4
        d2h_box[0] = bytes_out

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(2, min(args.steps, 3))
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1) / e2e_steps)
    e2e_value = nnz / (e2e_ms * 1e-3) / 1e9
    h2d_total, d2h_total = h2d, d2h_box[0]
    if world > 1:
        t = torch.tensor([float(h2d), float(d2h_box[0])], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        h2d_total, d2h_total = int(t[0].item()), int(t[1].item())
    del h_in, h_out, d_in, h_inv

    # ---- per-operator figures of the headline workload
    ops = {}
    for name, ms in op_ms.items():
        gbs = ALG_BYTES[name](n, nnz) / (ms * 1e-3) / 1e9
        ops[name] = {"ms": ms, "gnnz_per_s": nnz / (ms * 1e-3) / 1e9, "alg_gb_per_s": gbs,
                     "roofline_frac": gbs / (peak * world)}

    # ---- configs[2] (C3): CSR->CSC + DegreeDistribution on Erdos-Renyi 2^24 x 16
    c3 = None
    if not args.no_c3:
        del g_col, g_val, r_l, c_l, v_l
        torch.cuda.empty_cache()
        lib.trim()
        n3, row3, col3 = synth.erdos_renyi(1 << args.c3_log2n, 8, seed=43, device=dev)
        nnz3 = col3.numel()
        val3 = synth.hash_vals(nnz3, seed=7, device=dev)
        rp3, cc3, cv3 = lib.coo_to_csr(n3, n3, row3, col3, val3)
        del row3, col3, val3

        def timed(fn, reps=5):
            fn()
            barrier()
            x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x.record()
            for _ in range(reps):
                r = fn()
            y.record()
            barrier()
            return max_over_ranks(x.elapsed_time(y) / reps), r

        if world == 1:
            ms_t, csc = timed(lambda: lib.csr_to_csc(n3, n3, rp3, cc3, cv3))
            ms_d, _ = timed(lambda: lib.degree_distribution(n3, nnz3, rp3))
            back = lib.csr_to_csc(n3, n3, csc[0], csc[1], csc[2])
            ok3 = all(torch.equal(x, y) for x, y in zip(back, (rp3, cc3, cv3)))
        else:
            s3 = sharded.shard_csr(lib, n3, n3, rp3, cc3, cv3, rank, world)
            ms_t, cs = timed(lambda: mg.csr_to_csc(comm, s3))
            ms_d, _ = timed(lambda: sharded.degree_distribution(lib, s3))
            full = lib.csr_to_csc(n3, n3, rp3, cc3, cv3)     # single-GPU result on every rank
            clo, chi = cs.bounds[rank], cs.bounds[rank + 1]
            a3, b3 = int(full[0][clo]), int(full[0][chi])
            ok3 = torch.equal(cs.col_ptr + a3, full[0][clo:chi + 1]) and \
                torch.equal(cs.row, full[1][a3:b3]) and torch.equal(cs.vals, full[2][a3:b3])
            t = torch.tensor([1.0 if ok3 else 0.0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok3 = bool(t.item() > 0.5)
        c3 = {"workload": f"C3: CSR->CSC + DegreeDistribution on Erdos-Renyi 2^{args.c3_log2n} "
                          f"vertices, 8 undirected pairs per vertex ({nnz3} nnz)",
              "parity": bool(ok3),
              "parity_check": "transpose of the transpose == input" if world == 1 else
                              "every rank's column block == the slice of the single-GPU result"}
        for name, ms in (("csr_to_csc", ms_t), ("degree_distribution", ms_d)):
            gbs = ALG_BYTES[name](n3, nnz3) / (ms * 1e-3) / 1e9
            c3[name] = {"ms": ms, "gnnz_per_s": nnz3 / (ms * 1e-3) / 1e9, "alg_gb_per_s": gbs,
                        "roofline_frac": gbs / (peak * world)}
        del rp3, cc3, cv3
        torch.cuda.empty_cache()

    if comm is not None:
        comm.destroy()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- configs[1] (C2): RCM in ms (one GPU: its levels serialise)
    rcm = None
    if world == 1 and not args.no_rcm:
        lib.trim()
        rcm = rcm_block(lib, dev, peak)

    p2d = ops["permute2d"]
    line = {
        "metric": METRIC, "value": value, "unit": "GNNZ/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(scale=args.scale)},
        "matrix": {"n": n, "nnz": nnz,
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} nnz-balanced row blocks, peer-memory exchanges (sb200_mg_*)",
                   "l2_policy": "inputs (12.8 GB COO) far larger than the 126 MB L2; no flush needed"},
        "wall_ms_per_step": t_wall * 1e3 / args.steps,
        "e2e": {"value": e2e_value, "unit": "GNNZ/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total},
        "gpu_launches": launches,
        "clocks": clocks,
        "parity": parity_ok, "parity_detail": parity,
        "roofline": {"bound": "hbm",
                     "kernel": "Permute2D gather / renumber / row-sort kernels (ss_tile_kernel + "
                               "segmented long-row sort) -- the dominant operator of the step",
                     "achieved": p2d["alg_gb_per_s"] / world, "peak": peak, "unit": "GB/s",
                     "frac": p2d["roofline_frac"],
                     "traffic": ncu_traffic() if world == 1 else None,  # (captured on one GPU)
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALG_BYTES["permute2d"](n, nnz) // world,
                     "note": "achieved is per GPU: algorithmic bytes of Permute2D / N / its time"},
        "ops": ops,
    }
    if c3 is not None:
        line["c3"] = c3
    if rcm is not None:
        line["rcm"] = rcm
        line["rcm_ms"] = rcm["rcm_ms"]
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_pipeline_step(args.cpu_scale)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample",
                                                   "coo_to_csr_s", "degree_reorder_s",
                                                   "permute2d_s")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
