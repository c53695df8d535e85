#!/usr/bin/env python
"""Symbol-level configurations of BASELINE.json (parity-test cases, not the headline bench line):
times the device kernels on C2 (ISS column streams), C4 (1M fixed segments, reset contexts) and
C5 (skewed lengths, bypass suffixes, decode) and checks the round trip.   python tools/bench_symbols.py [c2|c4|c5 ...] [--scale S]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isscabac_b200 as I  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


def run(name, cfg, sym, sym_off, ctx, n_ctx):
    dev = torch.device("cuda")
    n_streams = sym_off.numel() - 1
    n_sym = sym.numel()
    out = {"config": name, "streams": n_streams, "symbols": n_sym}
    # two-pass encode: symbol-parallel binarizer -> op array -> wide encode kernel
    ms_bin, (ops, op_off) = timed(lambda: I.binarize_symbols(cfg, sym, sym_off), 2)
    n_bins = ops.numel()
    out["bins"] = n_bins
    longest = int((op_off[1:] - op_off[:-1]).max().item())
    stride = (longest // 4 + 64 + 15) & ~15
    enc = I.Encoded(torch.empty((n_streams, stride), dtype=torch.uint8, device=dev),
                    torch.empty(n_streams, dtype=torch.int32, device=dev), torch.zeros(4, dtype=torch.int32, device=dev))
    ms_enc, _ = timed(lambda: I.encode_ops(ops, op_off, ctx, out=enc))
    enc.check_overflow()
    pay = I.compact(enc)                                   # sizes the payload (one host sync), not timed
    out["payload_bytes"] = int(pay.byte_off[-1].item())
    L = I.lib()
    scratch = torch.empty(int(L.cabac_compact_scratch_bytes(n_streams)), dtype=torch.uint8, device=dev)
    ms_cmp, pay = timed(lambda: I.compact(enc, payload=pay.payload, byte_off=pay.byte_off, scratch=scratch))
    # fused encode (binarize + select + code in one kernel)
    ms_fused, enc2 = timed(lambda: I.encode_symbols(cfg, sym, sym_off, ctx, slab_stride=stride))
    assert bool((enc2.lengths == enc.lengths).all().item()), "fused and two-pass encoders disagree"
    ms_dec, (dec, ok) = timed(lambda: I.decode_symbols(cfg, pay, sym_off, ctx, sym_dtype=torch.uint8))
    assert bool(ok.all().item()) and bool((dec == sym).all().item()), "round trip failed"
    g = lambda ms: n_bins / (ms * 1e-3) / 1e9
    out.update(ms={"binarize": ms_bin, "encode_ops": ms_enc, "compact": ms_cmp, "encode_fused": ms_fused, "decode": ms_dec},
               gbins={"encode_two_pass": g(ms_bin + ms_enc + ms_cmp), "encode_fused": g(ms_fused + ms_cmp), "decode": g(ms_dec)},
               bits_per_symbol=8.0 * out["payload_bytes"] / max(n_sym, 1))
    print(json.dumps(out))


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["q", "c2", "c4", "c5"]
    scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    T = I.CM_COND0 | I.CM_COND1 | I.CM_CONDS0 | I.CM_CONDS1
    if "q" in which:    # the quantiser in front of C2: 1,638 tracks x (W 400 x 20, H 109 x 20), N = 8, dead zone 0.7, Lloyd-Max
        from isscabac_b200 import quantizer as QZ
        g.manual_seed(5)
        sizes = np.tile(np.array([400 * 20, 109 * 20], dtype=np.int64), int(1638 * scale))
        off = np.zeros(len(sizes) + 1, dtype=np.int64)
        np.cumsum(sizes, out=off[1:])
        # log(theta + eps) of gamma(0.6)-distributed factors (theta = Gamma(1)^(1/0.6) is close enough for a timing input)
        x = torch.log(torch.empty(int(off[-1]), dtype=torch.float64, device=dev).exponential_(1.0, generator=g) ** (1 / 0.6) + 1e-5)
        cfg = QZ.make_quant_cfg(N=8, GMM=1, deadzoneQuant=0.7)
        ms_q, (grp, cent, it) = timed(lambda: QZ.quantize_matrices((x, off), cfg, want_iters=True))
        print(json.dumps({"config": "quantiser for C2 (quantizeWrapper: dead zone 0.7 + Lloyd-Max, N = 8)", "matrices": len(sizes),
                          "elements": int(off[-1]), "ms": ms_q, "gelem_per_s": int(off[-1]) / (ms_q * 1e-3) / 1e9,
                          "lloyd_iterations_mean": float(it.float().mean().item()), "p_symbol0": float((grp == 0).float().mean().item())}))
    if "c2" in which:   # 65,520 column streams of 400 symbols, ISS profile, Nq = 8, P(0) = 0.7
        g.manual_seed(1)
        n_streams, rows = int(65520 * scale), 400
        u = torch.rand(n_streams * rows, generator=g, device=dev)
        sym = torch.where(u < 0.7, torch.zeros_like(u), 1 + torch.floor(torch.log(torch.rand(u.shape, generator=g, device=dev)) / np.log(0.6))).clamp_(0, 7).to(torch.uint8)
        off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * rows
        cfg = I.make_cfg(I.PROFILE_ISS, I.BIN_EG0, 8, 3, T, rows=rows)
        # the context-init statistics in front of the encoder (cabacInitContextModel.m): counters per matrix (20 column streams)
        ms_st, cnt = timed(lambda: I.iss_ctx_stats(cfg, sym, off, streams_per_group=20))
        print(json.dumps({"config": "C2 context-init statistics (cabacInitContextModel: counters per matrix, 20 column streams each)",
                          "groups": int(cnt.shape[0]), "counters_per_group": int(cnt.shape[1]), "ms": ms_st}))
        run("C2 ISS column streams", cfg, sym, off, torch.full((23,), 1, dtype=torch.uint8, device=dev), 23)
    if "c4" in which:   # 2^20 segments x 1024 symbols, Nq = 16, geometric, FLAT profile (8 contexts reset per segment)
        g.manual_seed(3)
        n_streams, per = int((1 << 20) * scale), 1024
        sym = torch.floor(torch.log(torch.rand(n_streams * per, generator=g, device=dev)) / np.log(0.5)).clamp_(0, 15).to(torch.uint8)
        off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * per
        cfg = I.make_cfg(I.PROFILE_FLAT, I.BIN_EG0, 16, 3, 0, rows=0)
        run("C4 1M fixed segments", cfg, sym, off, torch.full((8,), 1, dtype=torch.uint8, device=dev), 8)
    if "c5" in which:   # 2^20 streams, lognormal lengths, EG2, bypass suffixes
        rng = np.random.default_rng(4)
        n_streams = int((1 << 20) * scale)
        lens = np.clip(np.round(rng.lognormal(np.log(256), 1.0, size=n_streams)), 1, 65536).astype(np.int64)
        offn = np.zeros(n_streams + 1, dtype=np.int64)
        np.cumsum(lens, out=offn[1:])
        g.manual_seed(4)
        sym = torch.floor(-6.0 * torch.log(torch.rand(int(offn[-1]), generator=g, device=dev))).clamp_(0, 255).to(torch.uint8)
        cfg = I.make_cfg(I.PROFILE_FLAT_EPSUF, I.BIN_EG2, 256, 3, 0, rows=0)
        run("C5 skewed lengths", cfg, sym, torch.as_tensor(offn, device=dev), torch.full((4,), 1, dtype=torch.uint8, device=dev), 4)


if __name__ == "__main__":
    main()
