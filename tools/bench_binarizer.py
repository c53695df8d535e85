#!/usr/bin/env python
"""The symbol-parallel binarizer (cabac_binarize_symbols) on the symbol-level BASELINE shapes: the offsets-only call and
the full call timed separately, the u8 table kernels (k_bin_count8 / k_bin_emit8) beside the closed-form ones
(ISSCABAC_BIN8=0), ops and offsets of the two compared.   python tools/bench_binarizer.py [c2|c4|c5|u32 ...] [--scale S]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isscabac_b200 as I  # noqa: E402
from isscabac_b200 import engine as E  # noqa: E402


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


def run(name, cfg, sym, off):
    dev = sym.device
    L = I.lib()
    n, n_sym = off.numel() - 1, sym.numel()
    scratch = torch.empty(int(L.cabac_binarize_scratch_bytes(C.c_uint64(n_sym), C.c_uint32(n))), dtype=torch.uint8, device=dev)
    res = {"config": name, "streams": n, "symbols": n_sym}
    keep = {}
    for variant in ("1", "0"):
        os.environ["ISSCABAC_BIN8"] = variant
        op_off = torch.empty(n + 1, dtype=torch.int64, device=dev)

        def call(ops, cap):
            E.check(L.cabac_binarize_symbols(C.byref(cfg), C.c_uint32(n), E.vp(off), E.vp(sym), sym.element_size(), C.c_uint64(n_sym),
                                             E.vp(op_off), E.vp(ops) if ops is not None else None, C.c_uint64(cap), E.vp(scratch),
                                             E._stream_ptr()))
        call(None, 0)
        total = int(op_off[-1].item())
        ops = torch.empty(total + 64, dtype=torch.uint8, device=dev)
        ms1 = timed(lambda: call(None, 0))
        ms2 = timed(lambda: call(ops, total))
        keep[variant] = (ops[:total].clone(), op_off.clone())
        res["table_kernels" if variant == "1" else "closed_form_kernels"] = {
            "offsets_only_ms": ms1, "full_call_ms": ms2, "two_call_api_ms": ms1 + ms2,
            "gops_per_s_full_call": total / (ms2 * 1e-3) / 1e9,
            "hbm_frac_full_call": (2 * n_sym * sym.element_size() + total) / (ms2 * 1e-3) / 6.56e12}
        res["ops"] = total
    os.environ.pop("ISSCABAC_BIN8")
    res["identical"] = bool((keep["0"][0] == keep["1"][0]).all().item()) and bool((keep["0"][1] == keep["1"][1]).all().item())
    print(json.dumps(res), flush=True)
    assert res["identical"], "the two binarizer formulations disagree"


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c2", "c4", "c5", "u32"]
    scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    T = I.CM_COND0 | I.CM_COND1 | I.CM_CONDS0 | I.CM_CONDS1
    if "c2" in which:   # the shapes of tools/bench_symbols.py
        g.manual_seed(1)
        n_streams, rows = int(65520 * scale), 400
        u = torch.rand(n_streams * rows, generator=g, device=dev)
        sym = torch.where(u < 0.7, torch.zeros_like(u), 1 + torch.floor(torch.log(torch.rand(u.shape, generator=g, device=dev)) / np.log(0.6))).clamp_(0, 7).to(torch.uint8)
        off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * rows
        run("C2 ISS column streams", I.make_cfg(I.PROFILE_ISS, I.BIN_EG0, 8, 3, T, rows=rows), sym, off)
    if "c4" in which:
        g.manual_seed(3)
        n_streams, per = int((1 << 20) * scale), 1024
        sym = torch.floor(torch.log(torch.rand(n_streams * per, generator=g, device=dev)) / np.log(0.5)).clamp_(0, 15).to(torch.uint8)
        off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * per
        run("C4 1M fixed segments", I.make_cfg(I.PROFILE_FLAT, I.BIN_EG0, 16, 3, 0, rows=0), sym, off)
    if "c5" in which:
        rng = np.random.default_rng(4)
        n_streams = int((1 << 20) * scale)
        lens = np.clip(np.round(rng.lognormal(np.log(256), 1.0, size=n_streams)), 1, 65536).astype(np.int64)
        offn = np.zeros(n_streams + 1, dtype=np.int64)
        np.cumsum(lens, out=offn[1:])
        g.manual_seed(4)
        sym = torch.floor(-6.0 * torch.log(torch.rand(int(offn[-1]), generator=g, device=dev))).clamp_(0, 255).to(torch.uint8)
        run("C5 skewed lengths", I.make_cfg(I.PROFILE_FLAT_EPSUF, I.BIN_EG2, 256, 3, 0, rows=0), sym, torch.as_tensor(offn, device=dev))
    if "u32" in which:  # uniform over 32 values: half the strings have 9 ops (the table's long entries, two appends)
        g.manual_seed(6)
        n_streams, per = int((1 << 18) * scale), 1024
        sym = torch.randint(0, 32, (n_streams * per,), generator=g, device=dev, dtype=torch.uint8)
        off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * per
        run("uniform 32 values", I.make_cfg(I.PROFILE_FLAT, I.BIN_EG0, 32, 3, 0, rows=0), sym, off)


if __name__ == "__main__":
    main()
