"""Opcode evidence from the built library: which kernels use the Blackwell-specific machinery.

    python profiles/sass_opcodes.py > profiles/r2_sass_opcodes.md      (needs cuobjdump, no GPU)
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparsebase_b200", "libsb200.so")
WATCH = ["UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "STAS", "CCTL", "MEMBAR", "FENCE", "REDUX",
         "VOTE", "MATCH", "ATOMS", "ATOMG", "REDG", "LDL", "STL", "BAR", "SHFL", "LDGSTS", "UTMALDG"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fn, per, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        base = op.split(".")[0]
        total[base] += 1
        per[fn][base] += 1
dem = subprocess.run(["cu++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(per, dem)) if len(dem) == len(per) else {k: k for k in per}
print("# SASS opcode evidence (`cuobjdump -sass sparsebase_b200/libsb200.so`, sm_100a)\n")
print(f"{len(per)} kernels, {sum(total.values())} instructions.  Library-wide counts of the opcodes "
      "that matter here:\n")
print("| opcode | count | meaning |")
print("|---|---:|---|")
MEAN = {"UBLKCP": "cp.async.bulk (bulk-copy engine: radix-sort tile fetch)",
        "SYNCS": "mbarrier arrive / try_wait (bulk-copy completion, RCM cluster exchange)",
        "UCGABAR_ARV": "barrier.cluster.arrive", "UCGABAR_WAIT": "barrier.cluster.wait",
        "STAS": "st.async into a peer CTA's shared memory, completing on its mbarrier (RCM level exchange)",
        "REDUX": "warp reduce in one instruction",
        "VOTE": "ballots (stable ranking, compaction)", "MATCH": "match.any", "ATOMS": "shared-memory atomics",
        "ATOMG": "global atomics with result (claims)", "REDG": "global reductions", "LDL": "local-memory loads",
        "STL": "local-memory stores", "MEMBAR": "memory barriers", "FENCE": "proxy / mbarrier-init fences",
        "CCTL": "cache control", "BAR": "CTA barriers", "SHFL": "shuffles", "LDGSTS": "cp.async (not used)",
        "UTMALDG": "TMA tensor copies (not used: all copies are 1-D)"}
for op in WATCH:
    print(f"| {op} | {total.get(op, 0)} | {MEAN.get(op, '')} |")
print("\nNo tcgen05 / UTCMMA / HMMA instructions: nothing on this path is a contraction.\n")
print("Kernels using the bulk-copy engine, mbarriers, cluster instructions or local memory:\n")
print("| kernel | instructions | UBLKCP | SYNCS | UCGABAR_ARV | STAS | REDUX | VOTE | ATOMS | ATOMG | LDL | STL |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
cols = ["UBLKCP", "SYNCS", "UCGABAR_ARV", "STAS", "REDUX", "VOTE", "ATOMS", "ATOMG", "LDL", "STL"]
for k, c in per.items():
    if any(c.get(x, 0) for x in ["UBLKCP", "SYNCS", "UCGABAR_ARV", "STAS", "LDL", "STL"]):
        nm = names[k].replace("sb200::", "")
        nm = re.sub(r"\(.*", "", nm)
        nm = (nm[:100] + "...") if len(nm) > 100 else nm
        print(f"| {nm} | {sum(c.values())} | " + " | ".join(str(c.get(x, 0)) for x in cols) + " |")
