"""Aggregates the per-launch table of profiles/ncu_summary.py by kernel name.

    python profiles/ncu_by_kernel.py gpurun_out/r2_kernels_full.md > profiles/r2_kernels_by_kernel.md
"""
import collections
import sys

rows = [ln.strip().strip("|").split("|") for ln in open(sys.argv[1]) if ln.startswith("|")]
hdr, body = [c.strip() for c in rows[0]], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in body:
    r = [c.strip() for c in r]
    if "sb200" not in r[ix["kernel"]] and "::" in r[ix["kernel"]]:
        continue  # torch / library kernels that built the arguments
    a = agg.setdefault(r[ix["kernel"]], {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "best": None})
    t = float(r[ix["time us"]])
    a["n"] += 1
    a["t"] += t
    a["rd"] += float(r[ix["DRAM read MB"]])
    a["wr"] += float(r[ix["DRAM write MB"]])
    if a["best"] is None or t > a["best"][0]:
        a["best"] = (t, r)
print("| kernel | launches | total us | DRAM read MB | DRAM write MB | longest launch: us | grid x block "
      "| regs | DRAM GB/s | sm % | warps active % |")
print("|---|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|")
for k, a in agg.items():
    t, r = a["best"]
    print(f"| {k} | {a['n']} | {a['t']:.1f} | {a['rd']:.1f} | {a['wr']:.1f} | {t:.1f} | "
          f"{r[ix['grid x block']]} | {r[ix['regs']]} | {r[ix['DRAM GB/s']]} | {r[ix['sm %']]} | "
          f"{r[ix['warps active %']]} |")
