#!/usr/bin/env python
"""Op-array kernels in the latency-bound regime (strong scaling of C3, VERDICT r1 task 2): encode / decode time of
n_streams x n_bins for a few stream counts, and the cycles one bin costs a lone tile (32 streams on one warp).
   python tools/exp_ops_latency.py [--bins 65536] [--streams 32,1024,4096,8192,16384,32768] [--lib path.so ...]
Every (library, size) line also re-checks the round trip; with ISSCABAC_ENC_SPLIT unset the library picks the encoder."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(bins, streams):
    import torch
    import isscabac_b200 as I
    import bench as B
    dev = torch.device("cuda")
    ctx = torch.full((23,), 1, dtype=torch.uint8, device=dev)
    clk = 1.965e9

    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    for S in streams:
        ops = B.gen_ops_device(torch, 7, S, bins, dev)
        off = torch.arange(S + 1, dtype=torch.int64, device=dev) * bins
        stride = (bins // 4 + 64 + 15) & ~15
        enc = I.Encoded(torch.empty((S, stride), dtype=torch.uint8, device=dev), torch.empty(S, dtype=torch.int32, device=dev),
                        torch.zeros(4, dtype=torch.int32, device=dev))
        ms_e, _ = t(lambda: I.encode_ops(ops, off, ctx, out=enc))
        pay = I.compact(enc)
        out_b = torch.empty(S * bins, dtype=torch.uint8, device=dev)
        ok = torch.empty(S, dtype=torch.uint8, device=dev)
        ms_d, _ = t(lambda: I.decode_ops(pay, ops, off, ctx, bins=out_b, finish_ok=ok))
        good = bool(ok.all().item()) and bool((out_b == (ops & 1)).all().item()) and int(enc.overflow[0].item()) == 0
        print(json.dumps({"streams": S, "bins": bins, "encode_ms": round(ms_e, 4), "decode_ms": round(ms_d, 4),
                          "encode_cycles_per_bin": round(ms_e * 1e-3 * clk / bins, 1), "decode_cycles_per_bin": round(ms_d * 1e-3 * clk / bins, 1),
                          "gbins_enc_dec": round(2 * S * bins / ((ms_e + ms_d) * 1e-3) / 1e9, 1), "round_trip_ok": good,
                          "lib": os.path.basename(os.environ.get("ISSCABAC_LIB", "default")),
                          "enc_split": os.environ.get("ISSCABAC_ENC_SPLIT", "auto")}), flush=True)
        del ops, enc, pay, out_b


def main():
    a = sys.argv[1:]
    bins = int(a[a.index("--bins") + 1]) if "--bins" in a else 65536
    streams = [int(x) for x in (a[a.index("--streams") + 1] if "--streams" in a else "32,1024,4096,8192,16384,32768").split(",")]
    if "--child" in a:
        return child(bins, streams)
    libs = [None]
    i = 0
    while i < len(a):
        if a[i] == "--lib":
            libs.append(a[i + 1]); i += 1
        i += 1
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["ISSCABAC_LIB"] = lib
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--bins", str(bins), "--streams", ",".join(map(str, streams))], env=env)


if __name__ == "__main__":
    main()
