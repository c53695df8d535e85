#!/usr/bin/env python
"""Randomised GPU-vs-oracle sweep (test infrastructure, like tests/): many small random jobs with random shapes --
ragged and empty streams, terminate ops in the middle, misaligned op buffers, per-stream context inits, both encoder
formulations, every symbol profile / binarization -- each compared bit for bit with the oracle.
  python tests/gpu_tools/fuzz_gpu.py [--iters N] [--seed S]      (on a GPU box; prints one JSON line)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import isscabac_b200 as I  # noqa: E402
import oracle as O  # noqa: E402


def fuzz_ops(rng):
    n_streams = int(rng.choice([1, 2, 31, 32, 33, 100, 700, 3000]))
    n_ctx = int(rng.choice([1, 2, 3, 23, 60, 124]))
    max_len = int(rng.choice([0, 1, 15, 16, 17, 100, 1000, 5000]))
    lens = rng.integers(0, max_len + 1, size=n_streams)
    if rng.random() < 0.3:
        lens[rng.integers(0, n_streams, size=max(1, n_streams // 4))] = 0
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    code = rng.integers(0, n_ctx, size=n).astype(np.uint8)
    code[rng.random(n) < rng.choice([0.0, 0.25, 0.9])] = O.OP8_EP
    code[rng.random(n) < rng.choice([0.0, 0.001, 0.05])] = O.OP8_TRM
    bins = (rng.random(n) < rng.choice([0.05, 0.3, 0.5, 0.95])).astype(np.uint8)
    ops = ((code << 1) | bins).astype(np.uint8)
    per = rng.random() < 0.5
    ci = rng.integers(0, 126, size=(n_streams, n_ctx) if per else n_ctx).astype(np.uint8)
    stride = ((max_len * 7 // 8 + 80) + 15) & ~15
    s_ref, l_ref = O.encode_ops(ops, off, ci, out_stride=stride, n_threads=8)
    p_ref, b_ref = O.compact(s_ref, l_ref)
    # a terminate-1 in the middle of a stream ends what the reference decoder can read: decode parity only without them
    decodable = not ((code == O.OP8_TRM) & (bins == 1)).any()
    shift = int(rng.integers(0, 16))
    buf = torch.zeros(n + 32, dtype=torch.uint8, device="cuda")
    buf[shift:shift + n] = torch.as_tensor(ops, device="cuda")
    d_ops = buf[shift:shift + n] if n else buf[:0]
    for split in ("0", "1", "lat"):
        os.environ["ISSCABAC_ENC_SPLIT"] = "0" if split == "lat" else split
        os.environ["ISSCABAC_LAT"] = "1" if split == "lat" else "0"
        enc = I.encode_ops(d_ops, off.astype(np.int64), ci, slab_stride=stride)
        pay = I.compact(enc)
        torch.cuda.synchronize()
        enc.check_overflow()
        lens_g = enc.lengths.cpu().numpy().astype(np.uint32)
        assert (lens_g == l_ref).all(), ("lengths", split)
        assert (pay.byte_off.cpu().numpy().astype(np.uint64) == b_ref).all()
        assert (pay.payload.cpu().numpy()[:len(p_ref)] == p_ref).all(), ("payload", split)
    os.environ.pop("ISSCABAC_ENC_SPLIT")
    if decodable and n:
        for lat in ("0", "1"):
            os.environ["ISSCABAC_LAT"] = lat
            dbins, ok = I.decode_ops(pay, d_ops, off.astype(np.int64), ci)
            assert bool(ok.all().item()) and (dbins.cpu().numpy() == bins).all(), ("decode", lat)
    os.environ.pop("ISSCABAC_LAT")
    return n


def fuzz_symbols(rng):
    cases = [(O.PROFILE_DEMO, O.BIN_TU, 4, np.uint8), (O.PROFILE_DEMO, O.BIN_EG0, 4, np.uint8),
             (O.PROFILE_ISS, O.BIN_EG0, 8, np.uint8), (O.PROFILE_ISS, O.BIN_EG1, 40, np.uint16),
             (O.PROFILE_ISS, O.BIN_TU, 8, np.uint8), (O.PROFILE_FLAT, O.BIN_EG0, 16, np.uint8),
             (O.PROFILE_FLAT, O.BIN_EG2, 300, np.uint16), (O.PROFILE_FLAT_EPSUF, O.BIN_EG2, 256, np.uint8),
             (O.PROFILE_FLAT_EPSUF, O.BIN_EG0, 70000, np.uint32), (O.PROFILE_FLAT, O.BIN_FL32, 0, np.uint32),
             # small alphabets of the value-only profiles: symbol PAIRS in the ring encoder (with strings too long for the
             # pair table among them), bypass RUNS of every length in the tree decoder
             (O.PROFILE_FLAT_EPSUF, O.BIN_EG0, 16, np.uint8), (O.PROFILE_FLAT, O.BIN_EG1, 30, np.uint8),
             (O.PROFILE_FLAT_EPSUF, O.BIN_EG1, 200, np.uint8), (O.PROFILE_FLAT, O.BIN_TU, 7, np.uint8)]
    prof, meth, Nq, dt = cases[int(rng.integers(0, len(cases)))]
    rows = int(rng.choice([0, 1, 7, 109])) if prof == O.PROFILE_ISS else 0
    n_streams = int(rng.choice([1, 5, 33, 300]))
    counts = rng.integers(0, int(rng.choice([2, 50, 900])), size=n_streams)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    if n == 0:
        return 0
    T = int(rng.integers(0, 32))
    if meth == O.BIN_FL32:
        sym = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    else:
        sym = np.minimum(np.floor(rng.exponential(Nq / rng.choice([2.0, 6.0]) + 0.5, size=n)), Nq - 1).astype(dt)
    nctx = O.num_ctx(prof, 3)
    per = rng.random() < 0.5
    ci = rng.integers(0, 126, size=(n_streams, nctx) if per else nctx).astype(np.uint8)
    cfg, ocfg = I.make_cfg(prof, meth, Nq, 3, T, rows), O.make_cfg(prof, meth, Nq, 3, T, rows)
    stride = 16 * 1024
    s_ref, l_ref = O.encode_symbols(ocfg, sym, off, ci, stride, n_threads=8)
    w = int(l_ref.max()) if n_streams else 0
    live = np.arange(w)[None, :] < l_ref[:, None]
    for ring in ("0", "1"):              # both fused encoders: per-bin state machine / binarizer ahead of the coder through a ring
        os.environ["ISSCABAC_SYM_RING"] = ring
        enc = I.encode_symbols(cfg, sym, off.astype(np.int64), ci, slab_stride=stride)
        torch.cuda.synchronize()
        os.environ.pop("ISSCABAC_SYM_RING")
        enc.check_overflow()
        assert (enc.lengths.cpu().numpy().astype(np.uint32) == l_ref).all(), ("symbol lengths", ring)
        assert (enc.slab[:, :w].cpu().numpy()[live] == s_ref[:, :w][live]).all(), ("symbol bytes", ring)
    pay = I.compact(enc)
    tdt = {np.uint8: torch.uint8, np.uint16: torch.int16, np.uint32: torch.int32}[dt]
    for tree in ("0", "1"):              # both fused decoders: closed-form state machine / code-tree walk
        os.environ["ISSCABAC_SYM_TREE"] = tree
        dec, ok = I.decode_symbols(cfg, pay, off.astype(np.int64), ci, sym_dtype=tdt)
        torch.cuda.synchronize()
        os.environ.pop("ISSCABAC_SYM_TREE")
        assert bool(ok.all().item()) and (dec.cpu().numpy().view(dt)[:n] == sym).all(), ("symbol decode", tree)
    if n:
        per_stream = [O.symbols_to_ops(ocfg, sym[int(off[s]):int(off[s + 1])]) for s in range(n_streams)]
        want = np.concatenate(per_stream)
        want_off = np.zeros(n_streams + 1, dtype=np.int64)
        np.cumsum([len(x) for x in per_stream], out=want_off[1:])
        for bin8 in (("1", "0") if dt == np.uint8 else ("1",)):     # u8 symbols: the table kernels and the closed-form ones
            os.environ["ISSCABAC_BIN8"] = bin8
            ops, op_off = I.binarize_symbols(cfg, sym, off.astype(np.int64))
            os.environ.pop("ISSCABAC_BIN8")
            assert (ops.cpu().numpy() == want).all(), ("binarizer", bin8)
            assert (op_off.cpu().numpy() == want_off).all(), ("binarizer offsets", bin8)
    return n


def main():
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 200
    seed = int(sys.argv[sys.argv.index("--seed") + 1]) if "--seed" in sys.argv else 1
    t0 = time.time()
    done = {"ops": 0, "symbols": 0, "units": 0}
    for it in range(iters):
        rng = np.random.default_rng([seed, it])
        kind = "ops" if it % 2 == 0 else "symbols"
        try:
            done["units"] += fuzz_ops(rng) if kind == "ops" else fuzz_symbols(rng)
        except AssertionError as e:
            print(json.dumps({"fuzz": "FAILED", "seed": seed, "iteration": it, "kind": kind, "what": str(e)}))
            raise
        done[kind] += 1
    print(json.dumps({"fuzz": "ok", "seed": seed, "iterations": iters, **done, "seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
