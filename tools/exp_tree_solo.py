#!/usr/bin/env python
"""Fused decoder with / without the lone-SM option (ISSCABAC_TREE_SOLO), alternating in one process: C4 (2^20 equally long
streams) and C5 (2^20 ragged streams).   python tools/exp_tree_solo.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isscabac_b200 as I  # noqa: E402

dev = torch.device("cuda")


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def job(name):
    g = torch.Generator(device=dev)
    if name == "c4":
        g.manual_seed(3)
        n, per = 1 << 20, 1024
        sym = torch.floor(torch.log(torch.rand(n * per, generator=g, device=dev)) / np.log(0.5)).clamp_(0, 15).to(torch.uint8)
        off = torch.arange(n + 1, dtype=torch.int64, device=dev) * per
        return I.make_cfg(I.PROFILE_FLAT, I.BIN_EG0, 16, 3, 0, rows=0), sym, off, torch.full((8,), 1, dtype=torch.uint8, device=dev), 1024
    rng = np.random.default_rng(4)
    n = 1 << 20
    lens = np.clip(np.round(rng.lognormal(np.log(256), 1.0, size=n)), 1, 65536).astype(np.int64)
    offn = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offn[1:])
    g.manual_seed(4)
    sym = torch.floor(-6.0 * torch.log(torch.rand(int(offn[-1]), generator=g, device=dev))).clamp_(0, 255).to(torch.uint8)
    return (I.make_cfg(I.PROFILE_FLAT_EPSUF, I.BIN_EG2, 256, 3, 0, rows=0), sym, torch.as_tensor(offn, device=dev),
            torch.full((4,), 1, dtype=torch.uint8, device=dev), (int(lens.max()) * 5 // 4 + 64 + 15) & ~15)


for name in ("c4", "c5"):
    cfg, sym, off, ctx, stride = job(name)
    enc = I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride)
    pay = I.compact(enc)
    del enc
    res = {"0": [], "1": []}
    for it in range(5):
        for solo in ("0", "1"):
            os.environ["ISSCABAC_TREE_SOLO"] = solo
            res[solo].append(round(timed(lambda: I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8)), 2))
    dec, ok = I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8)
    assert bool(ok.all().item()) and bool((dec == sym).all().item())
    print(name, "default", res["0"], "lone SM", res["1"])
    del sym, pay, dec
