#!/usr/bin/env python
"""Per-instruction stall picture of the hottest straight-line stretch of a kernel in an .ncu-rep (source page):
for every SASS instruction its samples and the dominant stall reason -- for a lone warp this reads as the timeline of
one bin's dependent chain.   python tools/ncu_chain.py <rep> <kernel-substring> [n_instr=130] [start_index]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
n_show = int(sys.argv[3]) if len(sys.argv) > 3 else 130
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr, data = None, None, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]; data[cur] = []
    elif cur and r and r[0] == "Address":
        hdr = r
    elif cur and r and r[0].startswith("0x"):
        data[cur].append(r)
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
for k, v in data.items():
    if pat not in k:
        continue
    ns, ie, src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = sum(int(r[ns] or 0) for r in v)
    agg = {c: sum(int(r[hdr.index(c)] or 0) for r in v) for c in stalls}
    print(f"== {k[:70]}: {tot} samples; by reason: " + ", ".join(f"{c[6:]}={100.0 * n / max(tot, 1):.1f}%" for c, n in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    # hottest window of n_show consecutive instructions
    s = [int(r[ns] or 0) for r in v]
    if len(sys.argv) > 4:
        best = int(sys.argv[4])
    else:
        best, bsum, run = 0, -1, sum(s[:n_show])
        for i in range(0, max(1, len(s) - n_show)):
            if run > bsum:
                best, bsum = i, run
            run += s[i + n_show] - s[i] if i + n_show < len(s) else 0
    for i in range(best, min(len(v), best + n_show)):
        r = v[i]
        top = max(stalls, key=lambda c: int(r[hdr.index(c)] or 0))
        tn = int(r[hdr.index(top)] or 0)
        print(f"  [{i:5d}] x{int(r[ie]):>8d} {int(r[ns] or 0):6d}  {top[6:] if tn else '':14s} {r[src].strip()[:80]}")
