"""profiles/ncu_summary.py table of one Permute2D call -> the JSON bench.py reads.

    python profiles/traffic_json.py gpurun_out/r2_traffic_full.md > profiles/r2_traffic.json
"""
import json
import sys

rows = [ln.strip().strip("|").split("|") for ln in open(sys.argv[1]) if ln.startswith("|")]
hdr, body = [c.strip() for c in rows[0]], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
kernels, rd_total, wr_total, t_total = [], 0.0, 0.0, 0.0
for r in body:
    r = [c.strip() for c in r]
    rd, wr, t = float(r[ix["DRAM read MB"]]), float(r[ix["DRAM write MB"]]), float(r[ix["time us"]])
    kernels.append({"kernel": r[ix["kernel"]], "time_us": t, "dram_read_MB": rd, "dram_write_MB": wr})
    rd_total += rd
    wr_total += wr
    t_total += t
print(json.dumps({
    "source": "ncu --clock-control none (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum) of one sb200_permute2d call on C4 (R-MAT scale 26, "
              "DegreeReorder permutation): profiles/r2_traffic.sh; dram__bytes_read.sum + "
              "dram__bytes_write.sum summed over the kernels of the call",
    "permute2d_gather_dram_bytes_per_launch": int((rd_total + wr_total) * 1e6),
    "dram_read_bytes": int(rd_total * 1e6), "dram_write_bytes": int(wr_total * 1e6),
    "kernel_time_us_under_ncu": t_total, "kernels": kernels}, indent=1))
