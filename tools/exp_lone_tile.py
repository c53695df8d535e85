#!/usr/bin/env python
"""One encode + one decode launch of a lone tile (32 streams x N bins): the launch pair `ncu --set full --import-source on`
is pointed at to read per-instruction stall samples of a warp that runs alone on its scheduler.
   ISSCABAC_LAT=1 python tools/exp_lone_tile.py [--streams 32] [--bins 65536]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
import isscabac_b200 as I  # noqa: E402

a = sys.argv[1:]
S = int(a[a.index("--streams") + 1]) if "--streams" in a else 32
bins = int(a[a.index("--bins") + 1]) if "--bins" in a else 65536
dev = torch.device("cuda")
ctx = torch.full((23,), 1, dtype=torch.uint8, device=dev)
ops = B.gen_ops_device(torch, 7, S, bins, dev)
off = torch.arange(S + 1, dtype=torch.int64, device=dev) * bins
for _ in range(2):
    enc = I.encode_ops(ops, off, ctx, slab_stride=(bins // 4 + 64 + 15) & ~15)
    pay = I.compact(enc)
    out, ok = I.decode_ops(pay, ops, off, ctx)
torch.cuda.synchronize()
assert bool(ok.all().item()) and bool((out == (ops & 1)).all().item())
print("ok")
