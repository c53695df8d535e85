#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full report: DRAM bytes (read + write) per launch of
each profiled kernel.   python tools/traffic_from_ncu.py <rep> <bins_per_launch> [out.json]"""
import csv
import io
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import hot_kernel_hash  # noqa: E402  (the stamp bench.py checks before it cites these numbers)

rep, bins = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = name.split("(")[0].split("::")[-1].split("<")[0].strip()
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        tot += float(r[i]) * scale[units[i]]
    res[short] = {"dram_bytes_per_launch": tot, "bins_per_launch": bins,
                  "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}",
                  "csrc_hash": hot_kernel_hash()}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
