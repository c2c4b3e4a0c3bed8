"""Per-source-line stall samples from an ncu report.

    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv -k regex:<kernel> > x.csv
    python profiles/hotlines.py x.csv [top] [inst]   # "inst": rank by executed instructions
"""
import csv
import sys


def _int(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    fname, hdr, cur = None, None, None
    agg = {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            si = hdr.index("Warp Stall Sampling (All Samples)")
            ie = hdr.index("Instructions Executed")
            stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) <= si:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip())
            agg.setdefault(cur, [0, 0, {}])
            continue  # the source row carries the SUM of its sass rows; count sass rows only
        if cur is None:
            continue
        a = agg[cur]
        a[0] += _int(r[si])
        a[1] += _int(r[ie])
        for i, h in stalls:
            v = _int(r[i])
            if v:
                a[2][h] = a[2].get(h, 0) + v
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print(f"samples {tot}  warp-instructions {toti}")
    by_inst = len(sys.argv) > 3 and sys.argv[3] == "inst"
    for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_inst else 0])[:top]:
        st = ",".join(f"{k[6:]}:{v}" for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
        print(f"{100 * a[0] / tot:5.1f}% {100 * a[1] / toti:5.1f}%i {f}:{ln:<4d} {src[:70]:70s} [{st}]")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
