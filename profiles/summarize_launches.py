"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md
"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        t = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        t = t / 1e3 if unit == "ns" else t * 1e3 if unit == "ms" else t
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ours = "sb200::" in k
        name = k.split("(")[0].replace("void ", "")[:90]
        print(f"| {'' if ours else '(torch) '}{name} | {c} | {t:.1f} | {t / c:.1f} | {100 * t / total:.2f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
