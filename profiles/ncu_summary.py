"""Per-kernel table from an ncu report (one `ncu --set full` capture).

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sectors.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3,
         "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--metrics",
                          ",".join(METRICS)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def get(r, k):
        if k not in ix:
            return float("nan")
        try:
            return float(r[ix[k]].replace(",", "")) * SCALE.get(units[ix[k]], 1)
        except ValueError:
            return float("nan")

    print("| # | kernel | grid x block | regs | time us | DRAM read MB | DRAM write MB | "
          "DRAM GB/s | dram % | L2 sectors M | sm % | warps active % |")
    print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for n, r in enumerate(rows[2:]):
        name = r[ix["Kernel Name"]].replace("sb200::", "").split("(")[0][:70]
        t = get(r, METRICS[0])
        rd, wr = get(r, METRICS[1]), get(r, METRICS[2])
        print(f"| {n} | {name} | {r[ix['Grid Size']].strip()} x {r[ix['Block Size']].strip()} | "
              f"{get(r, METRICS[7]):.0f} | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
              f"{(rd + wr) / t / 1e3:.0f} | {get(r, METRICS[3]):.1f} | {get(r, METRICS[4]) / 1e6:.1f} | "
              f"{get(r, METRICS[5]):.1f} | {get(r, METRICS[6]):.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
