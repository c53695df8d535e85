#!/usr/bin/env python
"""Top SASS instructions of one kernel by a warp-stall reason, from the source page of an .ncu-rep.
  python tools/ncu_stalls.py <rep> <kernel-substring> <stall column, e.g. stall_long_sb> [top N]"""
import csv
import io
import subprocess
import sys

rep, pat, col = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 15
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr, data = None, None, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]
        data[cur] = []
    elif cur and r and r[0] == "Address":
        hdr = r
    elif cur and r and r[0].startswith("0x"):
        data[cur].append(r)
for k, v in data.items():
    if pat not in k:
        continue
    ci, si, ni = hdr.index(col), hdr.index("Source"), hdr.index("# Samples")
    tot = sum(int(r[ci] or 0) for r in v)
    alls = sum(int(r[ni] or 0) for r in v)
    print(f"== {k[:60]}: {col} = {tot} of {alls} samples ({100.0 * tot / max(alls, 1):.1f} %)")
    order = sorted(range(len(v)), key=lambda i: -int(v[i][ci] or 0))[:top]
    for i in sorted(order):
        print(f"  [{i:5d}] {int(v[i][ci] or 0):7d}  {v[i][si][:90]}")
