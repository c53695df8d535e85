#!/usr/bin/env python
"""Per-kernel share of a bench step from the ncu launch list tools/profile_round.sh writes (the last 10 launches = the two
timed steps of `bench.py --steps 2 --warmup 1`).   python tools/launch_summary.py gpurun_out/<tag>_launches.csv > profiles/<tag>_launches_summary.txt"""
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
last = rows[-10:]
order, tot = [], {}
for r in last:
    name = r[kn].split("(")[0].split("::")[-1]
    if name not in tot:
        order.append(name)
        tot[name] = 0.0
    tot[name] += float(r[mv]) / 1e6 / 2
step = sum(tot.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_(encode|decode|scan|compact), python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs")
print("# the two timed steps = the last 10 launches (cold-cache, serialised under the profiler: shares, not absolutes)")
for n in order:
    print(f"{n:30s} {tot[n]:8.3f} ms/step  {100 * tot[n] / step:5.1f} %")
print(f"{'step total':30s} {step:8.3f} ms")
