#!/usr/bin/env python
"""RCM narrow-regime sweep: bit-exactness against the oracle on small relatives, then per-level
time at pinned cluster sizes (SB200_RCM_CLUSTER) and with adaptive sizing.

    python profiles/rcm_sweep.py [--quick] [--out gpurun_out/rcm_sweep.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sparsebase_b200 import lib, synth  # noqa: E402


def time_rcm(n, rp, col, reps=2):
    os.environ["SB200_RCM_PROFILE"] = "1"   # one profiled run for the per-phase cycles
    lib.rcm_reorder(n, rp, col)
    prof = lib.rcm_last_stats()
    os.environ["SB200_RCM_PROFILE"] = "0"
    lib.rcm_reorder(n, rp, col)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        inv = lib.rcm_reorder(n, rp, col)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) / reps * 1e3
    st = lib.rcm_last_stats()
    lv = max(1, st["levels_narrow"])
    return inv, {"ms": round(ms, 2), "us_per_level": round(ms * 1e3 / (lv + st["levels_wide"]), 3),
                 "levels_narrow": st["levels_narrow"], "levels_wide": st["levels_wide"],
                 "bfs": st["bfs"], "resplits": st["resplits"], "resizes": st["resizes"],
                 "cyc_per_level": {k: round(v / max(1, prof["levels_narrow"]))
                                   for k, v in prof["phase_cycles"].items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib.load()
    import oracle_lib
    orc = oracle_lib.restated()
    res = {"parity": {}, "timing": {}}

    # ---- parity on small relatives, every cluster size
    small = {}
    n, rp, col, _ = synth.poisson2d(700, 300)
    small["poisson700x300"] = (n, rp, col)
    n, r, c = synth.band(300_000, 31, 0.5, seed=45, shuffle_seed=46)
    small["band300k"] = (n, synth.csr_from_sorted_coo(n, r), c)
    n, r, c = synth.band(60_000, 5, 0.9, seed=3, shuffle_seed=4)
    small["band60k_hb5"] = (n, synth.csr_from_sorted_coo(n, r), c)
    for name, (n, rp, col) in small.items():
        exp = orc.rcm_reorder(n, rp.numpy(), col.numpy())
        rpd, cold = rp.to(dev), col.to(dev)
        for cl in ("auto", "1", "2", "4", "8", "16"):
            if cl == "auto":
                os.environ.pop("SB200_RCM_CLUSTER", None)
            else:
                os.environ["SB200_RCM_CLUSTER"] = cl
            got = lib.rcm_reorder(n, rpd, cold).cpu().numpy()
            st = lib.rcm_last_stats()
            ok = bool(np.array_equal(got, exp))
            res["parity"][f"{name}/cluster={cl}"] = {
                "ok": ok, "levels_narrow": st["levels_narrow"], "levels_wide": st["levels_wide"],
                "resplits": st["resplits"], "resizes": st["resizes"]}
            print(f"parity {name} cluster={cl}: {'OK' if ok else 'MISMATCH'} {st}", flush=True)
    os.environ.pop("SB200_RCM_CLUSTER", None)

    # ---- timing
    big = {}
    g = 2048 if args.quick else 4096
    n, rp, col, _ = synth.poisson2d(g, g, device=dev)
    big[f"poisson{g}"] = (n, rp, col)
    nb = 2_000_000 if args.quick else 4_000_000
    n, r, c = synth.band(nb, 31, 0.5, seed=45, shuffle_seed=46, device=dev)
    big[f"band{nb}"] = (n, synth.csr_from_sorted_coo(n, r), c)
    del r
    for name, (n, rp, col) in big.items():
        ref = None
        # (small pinned clusters send most levels of the wide grid to the host-driven regime)
        for cl in (("auto", "8", "16") if name.startswith("poisson") else
                   ("auto", "1", "2", "4", "8", "16")):
            if cl == "auto":
                os.environ.pop("SB200_RCM_CLUSTER", None)
            else:
                os.environ["SB200_RCM_CLUSTER"] = cl
            inv, t = time_rcm(n, rp, col)
            if ref is None:
                ref = inv
            t["same_as_auto"] = bool(torch.equal(ref, inv))
            res["timing"][f"{name}/cluster={cl}"] = t
            print(f"timing {name} cluster={cl}: {t}", flush=True)
    os.environ.pop("SB200_RCM_CLUSTER", None)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
