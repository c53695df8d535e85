#!/usr/bin/env python
"""Per-SASS-instruction execution counts from an .ncu-rep (source page): prints, for one kernel,
the instructions grouped into runs with equal execution count, so hot-loop cost can be read off.

  python tools/ncu_hot.py gpurun_out/x.ncu-rep <kernel-substring> [--all]
"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, kernels = None, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]
        kernels[cur] = []
    elif cur and r and r[0].startswith("0x"):
        kernels[cur].append(r)
    elif cur and r and r[0] == "Address":
        kernels[cur + "#hdr"] = r
for k, v in kernels.items():
    if k.endswith("#hdr") or pat not in k:
        continue
    hdr = kernels[k + "#hdr"]
    ie, st, sm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in v)
    print(f"== {k}: {len(v)} SASS instructions, {tot} warp-instructions executed")
    if "--all" in sys.argv:
        for i, r in enumerate(v):
            print(f"{i:5d} {int(r[ie]):>12d} {int(r[sm]):>7d}  {r[st].strip()}")
        continue
    # runs of (roughly) equal execution count
    i = 0
    while i < len(v):
        c = int(v[i][ie])
        j = i
        while j + 1 < len(v) and abs(int(v[j + 1][ie]) - c) <= max(1, c // 50):
            j += 1
        n = j - i + 1
        samples = sum(int(r[sm]) for r in v[i:j + 1])
        if c * n > tot // 500:
            print(f"  [{i:5d}..{j:5d}] {n:4d} instr x {c:>11d} = {100.0 * c * n / tot:5.1f}% of issue, {samples:7d} samples   first: {v[i][st].strip()[:60]}")
        i = j + 1
