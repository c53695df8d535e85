#!/usr/bin/env python
"""Does the C3 step gain from running step k's decode CONCURRENTLY with step k+1's encode (two CUDA streams, double-buffered
slab / payload)?  Both kernels leave issue slots and ALU-pipe time unused at 3.5 warps per scheduler.
   python tools/exp_overlap.py [--streams 65536] [--bins 65536] [--steps 10]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
import isscabac_b200 as I  # noqa: E402

a = sys.argv[1:]
S = int(a[a.index("--streams") + 1]) if "--streams" in a else 65536
bins = int(a[a.index("--bins") + 1]) if "--bins" in a else 65536
steps = int(a[a.index("--steps") + 1]) if "--steps" in a else 10
dev = torch.device("cuda")
ops = B.gen_ops_device(torch, 3, S, bins, dev)
off = torch.arange(S + 1, dtype=torch.int64, device=dev) * bins
ctx = torch.full((23,), 1, dtype=torch.uint8, device=dev)
stride = (bins // 4 + 64 + 15) & ~15
L = I.lib()
bufs = []
for k in range(2):
    bufs.append(dict(enc=I.Encoded(torch.empty((S, stride), dtype=torch.uint8, device=dev), torch.empty(S, dtype=torch.int32, device=dev),
                                   torch.zeros(4, dtype=torch.int32, device=dev)),
                     scratch=torch.empty(int(L.cabac_compact_scratch_bytes(S)), dtype=torch.uint8, device=dev),
                     boff=torch.empty(S + 1, dtype=torch.int64, device=dev), pay=torch.empty(S * (bins // 6 + 64), dtype=torch.uint8, device=dev)))
out = torch.empty(S * bins, dtype=torch.uint8, device=dev)
ok = torch.empty(S, dtype=torch.uint8, device=dev)


def enc_part(b):
    I.encode_ops(ops, off, ctx, out=b["enc"])
    return I.compact(b["enc"], payload=b["pay"], byte_off=b["boff"], scratch=b["scratch"])


def seq(n):
    for k in range(n):
        p = enc_part(bufs[k & 1])
        I.decode_ops(p, ops, off, ctx, bins=out, finish_ok=ok)


sA, sB = torch.cuda.Stream(), torch.cuda.Stream()


def overlapped(n):
    """stream A: encode + compact of step k; stream B: decode of step k (after A's event), while A goes on with step k+1"""
    done_dec = [None, None]
    for k in range(n):
        b = bufs[k & 1]
        with torch.cuda.stream(sA):
            if done_dec[k & 1] is not None:
                sA.wait_event(done_dec[k & 1])       # the buffer pair is free again
            p = enc_part(b)
            ev = torch.cuda.Event()
            ev.record(sA)
        with torch.cuda.stream(sB):
            sB.wait_event(ev)
            I.decode_ops(p, ops, off, ctx, bins=out, finish_ok=ok)
            d = torch.cuda.Event()
            d.record(sB)
            done_dec[k & 1] = d
    torch.cuda.current_stream().wait_stream(sA)
    torch.cuda.current_stream().wait_stream(sB)


def timed(fn):
    fn(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


t_seq = timed(seq)
assert bool(ok.all().item()) and bool(((ops & 1) == out).all().item())
t_ovl = timed(overlapped)
assert bool(ok.all().item()) and bool(((ops & 1) == out).all().item())
print(json.dumps({"streams": S, "bins": bins, "sequential_ms_per_step": t_seq, "overlapped_ms_per_step": t_ovl,
                  "gbins_sequential": 2 * S * bins / (t_seq * 1e-3) / 1e9, "gbins_overlapped": 2 * S * bins / (t_ovl * 1e-3) / 1e9}))
