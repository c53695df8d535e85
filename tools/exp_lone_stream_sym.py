#!/usr/bin/env python
"""One fused encode + decode launch of a few long symbol streams (C5's profile: EG2, bypass-coded suffixes): the launch
`ncu --set full --import-source on` is pointed at to read per-instruction stall samples of the symbol kernels when one
stream's serial chain is all there is.   python tools/exp_lone_stream_sym.py [--streams 1] [--symbols 34000]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import isscabac_b200 as I  # noqa: E402

a = sys.argv[1:]
S = int(a[a.index("--streams") + 1]) if "--streams" in a else 1
L = int(a[a.index("--symbols") + 1]) if "--symbols" in a else 34000
dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(4)
cfg = I.make_cfg(I.PROFILE_FLAT_EPSUF, I.BIN_EG2, 256, 3, 0, rows=0)
ctx = torch.full((4,), 1, dtype=torch.uint8, device=dev)
sym = torch.floor(-6.0 * torch.log(torch.rand(S * L, generator=g, device=dev))).clamp_(0, 255).to(torch.uint8)
off = torch.arange(S + 1, dtype=torch.int64, device=dev) * L
for _ in range(2):
    enc = I.encode_symbols(cfg, sym, off, ctx, slab_stride=((L * 5) // 4 + 64 + 15) & ~15)
    pay = I.compact(enc)
    dec, ok = I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8)
torch.cuda.synchronize()
assert bool(ok.all().item()) and bool((dec == sym).all().item())
print("ok")
