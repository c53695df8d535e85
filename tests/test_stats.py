"""Statistics outputs.  CPU: the restatement of the reference's trace members
(oracle.trace_context) against the golden vectors recorded from the reference's own trace build
(oracle/_ref/libref_mex_trace.so = unmodified SimpleCABACMex.cpp compiled with -D_WIN32).  GPU: the
device trace kernel, the handle-level getEncoderStats/getDecoderStats and the per-context bit cost
against the same golden vectors and against the oracle."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle as O


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "mex_trace.json")) as f:
        return json.load(f)


def _dense(tr):
    m = np.zeros((128, 128), dtype=np.uint32)
    for a, b, c in tr:
        m[a, b] = c
    return m


def test_restatement_matches_reference_trace_build(golden):
    init = O.ctx_from_p0(golden["p0"])
    for c in range(4):
        bins = [b for b, cc in golden["seq"] if cc == c]
        for key, dec in (("enc", False), ("dec", True)):
            steps, trans, hist = O.trace_context(init[c], bins, decoder=dec)
            want = np.array(golden[key][c]["steps"], dtype=np.uint8).reshape(-1, 5)
            assert np.array_equal(steps, want)
            assert np.array_equal(trans, _dense(golden[key][c]["trans"]))
            assert hist.sum() == len(bins) and np.array_equal(hist, trans.sum(1))


def test_oracle_num_bits_matches_reference_per_bin(golden):
    """getNumBits() after every bin (what ctxCost of cabacEncode.m:63-65 is made of)."""
    L = O.lib()
    ci = O.ctx_from_p0(golden["p0"]).copy()
    out = np.zeros(4096, dtype=np.uint8)
    e = O._Enc()
    L.orc_enc_attach(C.byref(e), O._p(out, O._u8p), C.c_uint64(out.size))
    L.orc_enc_start(C.byref(e))
    got = []
    for b, c in golden["seq"]:
        L.orc_enc_bin(C.byref(e), int(b), C.byref(C.c_uint8.from_buffer(ci, int(c))))
        got.append(int(L.orc_enc_num_bits(C.byref(e))))
    assert got == golden["bits_after_bin"]


@pytest.mark.skipif(O.ref_mex_trace() is None, reason="oracle/_ref trace build not present")
def test_trace_build_live_random_session():
    rng = np.random.default_rng(5)
    p0 = list(rng.random(6))
    rc, out, err = O.mex_call(1, "initByProb", os.path.join(O.tmpdir(), "trace_live.bin"), p0, trace_build=True)
    assert rc == 0, err
    h = out[0]
    O.mex_call(0, "encodeStart", [h], trace_build=True)
    per = [[] for _ in p0]
    for _ in range(2000):
        c = int(rng.integers(0, 6))
        b = int(rng.random() < 0.4)
        per[c].append(b)
        assert O.mex_call(0, "encodeBin", [h], [b], [c], trace_build=True)[0] == 0
    O.mex_call(0, "encodeFinish", [h], trace_build=True)
    init = O.ctx_from_p0(p0)
    for c in range(6):
        steps, trans = O.mex_stats(h, c)
        s2, t2, _ = O.trace_context(init[c], per[c])
        assert np.array_equal(steps, s2) and np.array_equal(trans, t2)


# ----------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_handle_stats_equal_reference(golden, tmp_path):
    from isscabac_b200.matlab_api import MexError, SimpleCABACMexStats, cabacWrapper
    c = cabacWrapper(golden["p0"], str(tmp_path / "t.bin"))
    with pytest.raises(MexError, match="tracing is off"):
        c.getEncoderStats(0)
    c.setTrace(1)
    with pytest.raises(MexError, match="provide two variables"):
        SimpleCABACMexStats("getEncoderStats", c.cabac_handle, 0, nargout=1)
    c.encodeStart()
    bits = []
    for b, ctx in golden["seq"]:
        c.encodeBin(b, ctx)
    c.encodeFinish()
    assert open(tmp_path / "t.bin", "rb").read().hex() == golden["bytes"]
    c.decodeStart()
    for b, ctx in golden["seq"]:
        assert c.decodeBin(ctx) == b
    c.decodeFinish()
    for ctx in range(4):
        for key, fn in (("enc", c.getEncoderStats), ("dec", c.getDecoderStats)):
            trace, stats = fn(ctx)
            want = np.array(golden[key][ctx]["steps"], dtype=np.uint8).reshape(-1, 5)
            assert np.array_equal(trace.T, want)
            assert np.array_equal(stats.T, _dense(golden[key][ctx]["trans"]))
    # a context that never coded a bin: empty trace, zero matrix
    trace, stats = c.getEncoderStats(7)
    assert trace.shape == (5, 0) and not stats.any()
    c.close()


@pytest.mark.gpu
def test_gpu_batch_trace_and_cost(golden):
    import isscabac_b200 as I
    # (1) the golden session as one stream: transitions, usage, cost
    seq = np.array(golden["seq"], dtype=np.int64)
    ops = ((seq[:, 1] << 1) | seq[:, 0]).astype(np.uint8)
    init = O.ctx_from_p0(golden["p0"])
    t = I.ctx_trace_ops(ops, np.array([0, len(ops)]), init, want_trans=True, want_cost=True, want_steps=True, want_final=True)
    tr = t.trans.cpu().numpy().view(np.uint32)[0]
    for c in range(4):
        assert np.array_equal(tr[c], _dense(golden["enc"][c]["trans"]))
        assert int(t.usage[0, c].item()) == int((seq[:, 1] == c).sum())
    bits = np.array(golden["bits_after_bin"], dtype=np.int64)
    d = np.diff(np.concatenate([[0], bits]))
    want_cost = np.array([d[seq[:, 1] == c].sum() for c in range(4)] + [0])
    assert np.array_equal(t.cost_bits.cpu().numpy()[0], want_cost)
    st = t.step_states.cpu().numpy()
    for c in range(4):
        want = np.array(golden["enc"][c]["steps"], dtype=np.uint8).reshape(-1, 5)
        mine = st[seq[:, 1] == c]
        assert np.array_equal(I.trace_state(mine[:, 0]), want[:, 1]) and np.array_equal(mine[:, 0] & 1, want[:, 2])
        assert np.array_equal(I.trace_state(mine[:, 1]), want[:, 3]) and np.array_equal(mine[:, 1] & 1, want[:, 4])
    # (2) many ragged streams with bypass / terminate ops, pooled in groups of 3; u16 op format as well
    rng = np.random.default_rng(8)
    for width, n_ctx in ((1, 9), (2, 300)):
        n_streams = 41
        lens = rng.integers(0, 200, size=n_streams)
        off = np.zeros(n_streams + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        n = int(off[-1])
        code = rng.integers(0, n_ctx, size=n).astype(np.uint16)
        EP, TRM = (I.OP8_EP, I.OP8_TRM) if width == 1 else (I.OP16_EP, I.OP16_TRM)
        r = rng.random(n)
        code[r < 0.2] = EP
        code[r < 0.02] = TRM
        bins = (rng.random(n) < 0.35).astype(np.uint16)
        bins[code == TRM] = 0
        ops = ((code << 1) | bins).astype(np.uint8 if width == 1 else np.uint16)
        ci = rng.integers(0, 126, size=(n_streams, n_ctx)).astype(np.uint8)
        t = I.ctx_trace_ops(ops, off, ci, streams_per_group=3, want_trans=(width == 1), want_cost=True, want_final=True)
        groups = (n_streams + 2) // 3
        hist = np.zeros((groups, n_ctx, 128), dtype=np.int64)
        cost = np.zeros((groups, n_ctx + 1), dtype=np.int64)
        trans = np.zeros((groups, n_ctx, 128, 128), dtype=np.uint32)
        final = ci.copy()
        L = O.lib()
        for s in range(n_streams):
            g = s // 3
            out = np.zeros(1024, dtype=np.uint8)
            e = O._Enc()
            L.orc_enc_attach(C.byref(e), O._p(out, O._u8p), C.c_uint64(out.size))
            L.orc_enc_start(C.byref(e))
            ctx = final[s]
            for i in range(int(off[s]), int(off[s + 1])):
                before = int(L.orc_enc_num_bits(C.byref(e)))
                cd, b = int(code[i]), int(bins[i])
                if cd == EP:
                    L.orc_enc_ep(C.byref(e), b); slot = n_ctx
                elif cd == TRM:
                    L.orc_enc_trm(C.byref(e), b); slot = n_ctx
                else:
                    p = int(ctx[cd])
                    L.orc_enc_bin(C.byref(e), b, C.byref(C.c_uint8.from_buffer(ctx, cd)))
                    a = int(ctx[cd])
                    hist[g, cd, I.trace_state(p)] += 1
                    trans[g, cd, I.trace_state(p), I.trace_state(a)] += 1
                    slot = cd
                cost[g, slot] += int(L.orc_enc_num_bits(C.byref(e))) - before
        assert np.array_equal(t.state_hist.cpu().numpy(), hist)
        assert np.array_equal(t.cost_bits.cpu().numpy(), cost)
        assert np.array_equal(t.final_ctx.cpu().numpy(), final)
        if width == 1:
            assert np.array_equal(t.trans.cpu().numpy().view(np.uint32), trans)
