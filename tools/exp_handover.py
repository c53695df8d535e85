#!/usr/bin/env python
"""The wide encoder with one hand-over per SM (k_encode_ops_wide_ho, ISSCABAC_HANDOVER=1) against the plain kernel:
bytes and lengths of ALL streams compared, both timed.   python tools/exp_handover.py [n_streams] [ops_per_stream] [ragged]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isscabac_b200 as I  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    ragged = len(sys.argv) > 3 and sys.argv[3] == "ragged"
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    if ragged:
        lens = torch.randint(L // 2, L + 1, (n,), generator=g, device=dev, dtype=torch.int64)
    else:
        lens = torch.full((n,), L, dtype=torch.int64, device=dev)
    off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(lens, 0)
    tot = int(off[-1].item())
    code = torch.randint(0, 23, (tot,), generator=g, device=dev, dtype=torch.uint8)
    u = torch.rand(tot, generator=g, device=dev)
    bins = (u < (0.2 + 0.1 * (code % 5).float())).to(torch.uint8)
    code[torch.rand(tot, generator=g, device=dev) < 0.25] = 126
    ops = (code << 1) | bins
    del code, u, bins
    ctx = torch.randint(0, 126, (23,), generator=g, device=dev, dtype=torch.uint8)
    stride = (L // 4 + 64 + 15) & ~15
    os.environ["ISSCABAC_ENC_SPLIT"] = "0"
    res = {}
    outs = {}
    fracs = [a for a in sys.argv[4:] if a.isdigit()] or ["5"]
    for ho in ["0"] + fracs:
        os.environ["ISSCABAC_HANDOVER"] = "0" if ho == "0" else "1"
        os.environ["ISSCABAC_HANDOVER_EIGHTHS"] = ho if ho != "0" else "5"
        enc = I.Encoded(torch.zeros((n, stride), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int32, device=dev),
                        torch.zeros(4, dtype=torch.int32, device=dev))
        I.encode_ops(ops, off, ctx, out=enc)
        torch.cuda.synchronize()
        outs[ho] = enc
        scratch = I.Encoded(torch.empty((n, stride), dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
                            torch.zeros(4, dtype=torch.int32, device=dev))
        res["plain" if ho == "0" else "handover at %s/8" % ho] = timed(lambda: I.encode_ops(ops, off, ctx, out=scratch))
        del scratch
    same_len, same_bytes = True, True
    w = int(outs["0"].lengths.max().item())
    live = torch.arange(w, device=dev)[None, :] < outs["0"].lengths[:, None]
    for k in fracs:
        same_len = same_len and bool((outs["0"].lengths == outs[k].lengths).all().item())
        same_bytes = same_bytes and bool(((outs["0"].slab[:, :w] == outs[k].slab[:, :w]) | ~live).all().item())
    # decode: the plain encoder's payload through the plain and the hand-over decoder, bins and finish flags of all streams
    os.environ["ISSCABAC_HANDOVER"] = "0"
    pay = I.compact(outs["0"])
    dres, douts = {}, {}
    for ho in ("0", "1"):
        os.environ["ISSCABAC_HANDOVER"] = ho
        os.environ["ISSCABAC_HANDOVER_EIGHTHS"] = fracs[-1]
        bins_o, ok = I.decode_ops(pay, ops, off, ctx)
        torch.cuda.synchronize()
        douts[ho] = (bins_o, ok)
        dres["decode plain" if ho == "0" else "decode handover at %s/8" % fracs[-1]] = timed(lambda: I.decode_ops(pay, ops, off, ctx))
    dec_ok = bool(douts["1"][1].all().item()) and bool((douts["1"][0] == (ops & 1)).all().item()) and bool((douts["0"][0] == douts["1"][0]).all().item())
    res.update(dres)
    print(json.dumps({"streams": n, "ops_per_stream": L, "ragged": ragged, "decode_identical": dec_ok, "decode_kernel": I.lib().cabac_decode_ops_kernel(n, 23).decode(), "ms": res, "gbins": {k: tot / (v * 1e-3) / 1e9 for k, v in res.items()},
                      "identical_lengths": same_len, "identical_bytes": same_bytes}), flush=True)
    assert same_len and same_bytes and dec_ok


if __name__ == "__main__":
    main()
