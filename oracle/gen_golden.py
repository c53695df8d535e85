"""TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/* from the UNMODIFIED
reference engine (oracle/_ref, built from /root/reference by oracle/Makefile).

    python oracle/gen_golden.py          # run in the build container (needs /root/reference)

The reference ships no golden bitstreams, so the pinned
vectors are outputs of the reference itself:

* kat.json            K0-K8 of SURVEY.md 4.1 (engine-level scripts -> bytes, final ctx states)
* random_ops.npz      many short random streams (u8 ops incl. terminate bins, shared and
                      per-stream context init, ragged/empty streams) -> reference bytes
* random_ops16.npz    the same with the u16 op format and 300 contexts
* prob_to_state.json  initContextModelsByP0Prob over k/255 and edge probabilities
* mex_session.json    a transcript of the reference mexFunction (stub mex.h): commands,
                      outputs, error strings, resulting file bytes
* symbols_refengine.npz  symbol-level cases: ops produced by the oracle's restatement of the
                      MATLAB binarizer/context rules (no reference vectors exist for those),
                      then encoded by the REFERENCE engine.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def ctxb(mps, state):
    return (state << 1) | mps


def kat_scripts():
    k = {}
    k["K0"] = dict(script=[], ctx=[ctxb(1, 0)])
    k["K1"] = dict(script=[(0, b, 0) for b in (0, 0, 1, 0, 1, 1)] + [(1, b, 0) for b in (1, 0, 0, 1, 0)] +
                   [(2, 18, 5)] + [(0, b, 1) for b in (1, 1, 0, 1, 1, 1)], ctx=[ctxb(0, 20), ctxb(1, 0)])
    k["K2"] = dict(script=[(0, b, 0) for b in (0, 0, 1, 0, 1, 1)] + [(0, b, 1) for b in (1, 1, 0, 1, 1, 1)],
                   ctx=[ctxb(0, 20), ctxb(1, 0)])
    v = 0xDEADBEEF
    k["K3"] = dict(script=[(1, (v >> (31 - i)) & 1, 0) for i in range(32)], ctx=[ctxb(1, 0)])
    k["K4"] = dict(script=[(2, v, 32)], ctx=[ctxb(1, 0)])
    k["K5"] = dict(script=[(0, 0, 0)] * 100, ctx=[ctxb(1, 0)])
    k["K6"] = dict(script=[(0, i & 1, 0) for i in range(64)], ctx=[ctxb(1, 0)])
    for name, seed in (("K7", 135), ("K8", 534)):
        s = []
        for o in O.xorshift_ops(seed):
            code, b = int(o) >> 1, int(o) & 1
            s.append((1, b, 0) if code == O.OP8_EP else (0, b, code))
        k[name] = dict(script=s, ctx=[ctxb(1, 0)] * 4)
    # extra: terminate bins inside a stream, state-63 context, mps=0/state=0 init (gotcha 12)
    k["X_trm"] = dict(script=[(0, 1, 0), (3, 0, 0), (1, 1, 0), (3, 0, 0), (0, 0, 0), (3, 0, 0), (0, 1, 0), (1, 0, 0)] +
                      [(0, i % 3 == 0, 0) for i in range(40)] + [(3, 0, 0)] * 30, ctx=[ctxb(1, 5)])
    # a terminate-1 bin in mid-stream is encodable but not decodable by the reference (its
    # decodeBinTrm(1) does not renormalise, CABAC_ArithmeticDecoder.cpp:427-436): encode-only vector
    k["X_trm1"] = dict(script=[(0, 1, 0), (3, 1, 0), (0, 1, 0), (1, 0, 0), (3, 1, 0), (0, 0, 0)], ctx=[ctxb(1, 5)])
    k["X_st63"] = dict(script=[(0, 1, 0)] * 5 + [(0, 0, 0)] + [(1, 1, 0)] * 3, ctx=[ctxb(1, 63)])
    k["X_eq0"] = dict(script=[(0, b, 0) for b in (1, 1, 0, 1, 1, 1)], ctx=[ctxb(0, 0)])
    k["X_eq1"] = dict(script=[(0, b, 0) for b in (1, 1, 0, 1, 1, 1)], ctx=[ctxb(1, 0)])
    k["X_binsep"] = dict(script=[(0, 1, 0), (2, 0x1ABCDE, 21), (0, 0, 0), (2, 1, 1), (2, 0x3FF, 10), (2, 0, 9)],
                         ctx=[ctxb(0, 10)])
    return k


def gen_kat():
    out = {}
    for name, k in kat_scripts().items():
        data, fin = O.encode_script(k["script"], k["ctx"], impl="ref")
        # decode transcript from the reference decoder
        # (state 63 drives the reference decoder out of its table bounds after an LPS --
        #  range drops to 128 and (range>>6)-4 goes negative -- so no decode transcript there)
        dscript = [(kk, 0, b) for kk, a, b in k["script"]]
        dec = None if name in ("X_st63", "X_trm1") else O.decode_script(dscript, k["ctx"], data, impl="ref")
        out[name] = dict(script=[list(map(int, e)) for e in k["script"]], ctx=[int(c) for c in k["ctx"]],
                         bytes=data.hex(), ctx_final=[int(c) for c in fin], decoded=dec)
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(out, f, indent=0)
    return out


def random_ops(seed, n_streams, max_len, n_ctx, width, p_ep=0.25, p_trm=0.01):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, max_len + 1, size=n_streams)
    lens[:4] = (0, 1, 2, max_len)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    ctx = rng.integers(0, n_ctx, size=n)
    pone = 0.05 + 0.9 * rng.random(n_ctx)
    bins = (rng.random(n) < pone[ctx]).astype(np.uint32)
    kind = rng.random(n)
    ep, trm = (O.OP8_EP, O.OP8_TRM) if width == 1 else (O.OP16_EP, O.OP16_TRM)
    code = ctx.astype(np.uint32)
    code[kind < p_ep] = ep
    is_trm = (kind >= p_ep) & (kind < p_ep + p_trm)
    code[is_trm] = trm
    bins[is_trm] = 0  # terminate bin 1 only ends a stream
    ops = ((code << 1) | bins).astype(np.uint8 if width == 1 else np.uint16)
    return ops, off, rng


def gen_random(fname, seed, n_streams, max_len, n_ctx, width):
    ops, off, rng = random_ops(seed, n_streams, max_len, n_ctx, width)
    shared = rng.integers(0, 126, size=n_ctx).astype(np.uint8)           # states 0..62, both mps
    per = rng.integers(0, 126, size=(n_streams, n_ctx)).astype(np.uint8)
    res = {"ops": ops, "op_off": off, "ctx_shared": shared, "ctx_per": per}
    for tag, ci in (("shared", shared), ("per", per)):
        slab, lens = O.encode_ops(ops, off, ci, impl="ref", n_threads=8)
        payload, boff = O.compact(slab, lens)
        bins, ok = O.decode_ops(payload, boff, ops, off, ci, impl="ref", n_threads=8)
        assert ok.all() and (bins == (ops & 1)).all()
        res["payload_" + tag] = payload
        res["lens_" + tag] = lens
    np.savez_compressed(os.path.join(OUT, fname), **res)


def gen_prob():
    ps = [k / 255.0 for k in range(256)] + [0.5, 0.01875, 0.0187, 0.98125, 0.9813, 0.25, 0.75, 1.0, 0.0, 0.4999999, 0.3, 0.1]
    st = O.ref_prob_to_state(ps)
    with open(os.path.join(OUT, "prob_to_state.json"), "w") as f:
        json.dump({"p0": ps, "ctx": [int(x) for x in st]}, f)


def gen_mex():
    fn = os.path.join(O.tmpdir(), "golden_mex.bin")
    log = []

    def call(nlhs, *args, hide_handle=True):
        rc, out, err = O.mex_call(nlhs, *args)
        shown = []
        for a in args:
            if isinstance(a, str):
                shown.append(a if a != fn else "<fn>")
            else:
                shown.append(np.asarray(a, dtype=float).tolist())
        log.append(dict(nlhs=nlhs, args=shown, rc=rc, out=out, err=err))
        return rc, out, err

    # errors first
    call(0, "bogus")
    call(0)
    call(1, "initByProb")
    call(1, "initByProb", fn, [0.5], [1.0])
    call(1, "initByProb", [1.0], [0.5])
    call(1, "initByProb", fn, "notdouble")
    call(0, "encodeStart")
    call(0, "encodeStart", [0.0])
    # K2 session (initByState) + getNumBits
    rc, out, _ = call(1, "initByState", fn, np.array([[0, 0, 20], [1, 1, 0]]).T)
    h = out[0]
    log[-1]["out"] = ["<handle>"]
    H = [h]

    def hcall(nlhs, cmd, *rest):
        rc, out, err = O.mex_call(nlhs, cmd, H, *rest)
        log.append(dict(nlhs=nlhs, args=[cmd, "<handle>"] + [np.asarray(r, dtype=float).tolist() for r in rest],
                        rc=rc, out=out, err=err))
        return rc, out, err

    hcall(0, "encodeStart")
    hcall(0, "encodeBin", [1.0])                  # wrong arity
    hcall(0, "encodeBin", [2.0], [0.0])           # bad bin
    for b in (0, 0, 1, 0, 1, 1):
        hcall(0, "encodeBin", [float(b)], [0.0])
    hcall(1, "getNumBits")
    for b in (1, 1, 0, 1, 1, 1):
        hcall(0, "encodeBin", [float(b)], [1.0])
    hcall(1, "getNumBits")
    hcall(1, "getNumBits", [1.0])                 # wrong arity
    hcall(0, "encodeFinish")
    hcall(1, "getNumBits")
    log.append(dict(file=open(fn, "rb").read().hex()))
    hcall(0, "decodeStart")
    hcall(0, "decodeBin", [0.0])                  # nlhs != 1
    for _ in range(6):
        hcall(1, "decodeBin", [0.0])
    for _ in range(6):
        hcall(1, "decodeBin", [1.0])
    hcall(0, "decodeFinish")
    # initByProb session, longer, exercising getNumBits lag
    rng = np.random.default_rng(7)
    p0 = np.array([0.5, 0.9, 0.2])
    rc, out, _ = call(1, "initByProb", fn, p0)
    H[0] = out[0]
    log[-1]["out"] = ["<handle>"]
    hcall(0, "encodeStart")
    bins = (rng.random(200) < 0.3).astype(int)
    ctxs = rng.integers(0, 3, size=200)
    for i, (b, c) in enumerate(zip(bins, ctxs)):
        hcall(0, "encodeBin", [float(b)], [float(c)])
        if i % 10 == 9:
            hcall(1, "getNumBits")
    hcall(0, "encodeFinish")
    hcall(1, "getNumBits")
    log.append(dict(file=open(fn, "rb").read().hex()))
    hcall(0, "decodeStart")
    for c in ctxs:
        hcall(1, "decodeBin", [float(c)])
    hcall(0, "decodeFinish")
    # the SAME handle codes a second stream: contexts are NOT re-initialised by encodeStart
    # (SimpleCABACMex.cpp:186-209) and the bit counter keeps running (CABAC_BitstreamFile.h:70)
    hcall(0, "encodeStart")
    bins2 = (rng.random(150) < 0.3).astype(int)
    ctxs2 = rng.integers(0, 5, size=150)          # contexts 3,4 were never initialised: constructor default
    for i, (b, c) in enumerate(zip(bins2, ctxs2)):
        hcall(0, "encodeBin", [float(b)], [float(c)])
        if i % 25 == 24:
            hcall(1, "getNumBits")
    hcall(0, "encodeFinish")
    hcall(1, "getNumBits")
    log.append(dict(file=open(fn, "rb").read().hex()))
    os.unlink(fn)
    with open(os.path.join(OUT, "mex_session.json"), "w") as f:
        json.dump(log, f)


def gen_symbols():
    """Symbol-level cases: oracle binarizer/ctx rules -> ops -> REFERENCE engine bytes."""
    rng = np.random.default_rng(11)
    cases = {}
    allt = O.CM_COND0 | O.CM_COND1 | O.CM_CONDS0 | O.CM_CONDS1
    specs = [
        ("demo_tu", O.PROFILE_DEMO, O.BIN_TU, 4, 3, 0, 0, 2000),
        ("demo_eg0", O.PROFILE_DEMO, O.BIN_EG0, 4, 3, 0, 0, 2000),
        ("iss_eg0", O.PROFILE_ISS, O.BIN_EG0, 8, 3, allt, 50, 1000),
        ("iss_eg1_all", O.PROFILE_ISS, O.BIN_EG1, 16, 2, allt | O.CM_CONDBINLFT, 25, 500),
        ("iss_tu", O.PROFILE_ISS, O.BIN_TU, 6, 3, allt, 40, 400),
        ("flat_eg0", O.PROFILE_FLAT, O.BIN_EG0, 16, 3, 0, 0, 1024),
        ("flat_epsuf_eg2", O.PROFILE_FLAT_EPSUF, O.BIN_EG2, 256, 3, 0, 0, 700),
    ]
    for name, prof, meth, Nq, Nlbp, types, rows, n in specs:
        p = 0.55 ** np.arange(Nq)
        p /= p.sum()
        sym = rng.choice(Nq, size=n, p=p).astype(np.uint32)
        cfg = O.make_cfg(prof, meth, Nq, Nlbp, types, rows)
        ops = O.symbols_to_ops(cfg, sym)
        nctx = O.num_ctx(prof, Nlbp)
        ci = O.ctx_from_p0(O.matlab_uint8(rng.random(nctx) * 255) / 255.0)
        off = np.array([0, len(ops)], dtype=np.uint64)
        slab, lens = O.encode_ops(ops, off, ci, impl="ref")
        cases[name + "_sym"] = sym
        cases[name + "_cfg"] = np.array([prof, meth, Nq, Nlbp, types, rows], dtype=np.int64)
        cases[name + "_ctx"] = ci
        cases[name + "_ops"] = ops
        cases[name + "_bytes"] = slab[0, :lens[0]].copy()
    np.savez_compressed(os.path.join(OUT, "symbols_refengine.npz"), **cases)


def gen_trace():
    """getEncoderStats / getDecoderStats of the reference's trace build (oracle/_ref/libref_mex_trace.so =
    the unmodified SimpleCABACMex.cpp with -D_WIN32, CommonDef.h:39-40): step logs and transition counts
    of every context, plus getNumBits() after every bin (the ctxCost of cabacEncode.m:63-65)."""
    rng = np.random.default_rng(77)
    p0 = [0.3, 0.7, 0.5, 0.92]
    fn = os.path.join(O.tmpdir(), "golden_trace.bin")
    rc, out, err = O.mex_call(1, "initByProb", fn, p0, trace_build=True)
    assert rc == 0, err
    h = out[0]
    O.mex_call(0, "encodeStart", [h], trace_build=True)
    seq, bits = [], []
    pr = [0.3, 0.7, 0.5, 0.05]
    for _ in range(1200):
        c = int(rng.integers(0, 4))
        b = int(rng.random() < pr[c])
        seq.append([b, c])
        rc, _, err = O.mex_call(0, "encodeBin", [h], [b], [c], trace_build=True)
        assert rc == 0, err
        bits.append(int(O.mex_call(1, "getNumBits", [h], trace_build=True)[1][0]))
    O.mex_call(0, "encodeFinish", [h], trace_build=True)
    data = open(fn, "rb").read()
    O.mex_call(0, "decodeStart", [h], trace_build=True)
    for b, c in seq:
        rc, out, err = O.mex_call(1, "decodeBin", [h], [c], trace_build=True)
        assert rc == 0 and int(out[0]) == b
    O.mex_call(0, "decodeFinish", [h], trace_build=True)
    res = {"p0": p0, "seq": seq, "bits_after_bin": bits, "bytes": data.hex(), "enc": [], "dec": []}
    for c in range(4):
        for key, dec in (("enc", False), ("dec", True)):
            steps, trans = O.mex_stats(h, c, decoder=dec)
            nz = np.argwhere(trans)
            res[key].append({"steps": steps.reshape(-1).tolist(),
                             "trans": [[int(a), int(b), int(trans[a, b])] for a, b in nz]})
    with open(os.path.join(OUT, "mex_trace.json"), "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    assert O.ref() is not None, "needs oracle/_ref (build container with /root/reference)"
    os.makedirs(OUT, exist_ok=True)
    k = gen_kat()
    for name in ("K0", "K1", "K2", "K3", "K4", "K5", "K6", "K7", "K8"):
        print(name, k[name]["bytes"])
    gen_random("random_ops.npz", 101, 600, 300, 23, 1)
    gen_random("random_ops16.npz", 102, 120, 400, 300, 2)
    gen_prob()
    gen_mex()
    gen_symbols()
    gen_trace()
    print("golden vectors written to", OUT)
