"""GPU: the ISS application flow (cabacInitContextModel -> uint8 side info -> encode -> decode) through
isscabac_b200/coder.py against the oracle's restatement of the same MATLAB code, with the stream
bytes checked against the reference engine coding the same (bin, ctx) sequence."""
import numpy as np
import pytest

import oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ALLT = O.CM_COND0 | O.CM_COND1 | O.CM_CONDS0 | O.CM_CONDS1


def iss_matrix(rng, rows, cols, Nq=8):
    """Dead-zone-like statistics: P(0) ~ 0.7, geometric tail, AR(1) correlation down the columns."""
    x = rng.standard_normal((rows, cols))
    for d in range(1, rows):
        x[d] = 0.8 * x[d - 1] + 0.6 * x[d]
    a = np.abs(x)
    thr = np.quantile(a, 0.7)
    q = np.floor(np.maximum(a - thr, 0) / (thr * 0.5 + 1e-9) + (a > thr)).astype(np.int64)
    return np.minimum(q, Nq - 1).astype(np.uint32)


@pytest.mark.parametrize("types,method,Nq", [
    (["cond0", "cond1", "conds0", "conds1"], "DEC2EG0", 8),
    (["cond0", "cond1", "condbinlft", "conds0", "conds1"], "DEC2EG1", 16),
    (["cond0"], "DEC2TU", 4),
    ([], "DEC2EG0", 8),
])
def test_ctx_stats_match_oracle(types, method, Nq):
    import isscabac_b200 as I
    rng = np.random.default_rng(3)
    mats = [iss_matrix(rng, 400, 20, Nq), iss_matrix(rng, 400, 7, Nq), iss_matrix(rng, 400, 1, Nq)]
    cfg = I.make_cfg(I.PROFILE_ISS, method, Nq, 3, types, rows=400)
    sym = np.concatenate([m.T.reshape(-1) for m in mats]).astype(np.uint32)
    off = np.concatenate([[0], np.cumsum([m.size for m in mats])]).astype(np.int64)
    cnt = I.iss_ctx_stats(cfg, sym, off, 1)
    p0, q, st = I.iss_ctx_from_counters(cfg, cnt)
    for g, m in enumerate(mats):
        want = O.iss_ctx_init(m, Nq, I.METHODS[method], 3, cfg.types)
        assert np.array_equal(p0[g], want), (g, p0[g], want)           # same integer counts, same double arithmetic
        wq = O.matlab_uint8(want * 255)
        assert np.array_equal(q[g], wq)
        assert np.array_equal(st[g], O.ctx_from_p0(wq.astype(np.float64) / 255))
    # column streams pooled per matrix = the matrix statistics
    coff = np.concatenate([[0], np.cumsum(np.full(20, 400))]).astype(np.int64)
    cnt2 = I.iss_ctx_stats(cfg, mats[0].T.reshape(-1).astype(np.uint32), coff, 20)
    assert np.array_equal(cnt2.cpu().numpy()[0], cnt.cpu().numpy()[0])
    # equalProb: every context starts at p = 0.5 -> uint8 128 -> (mps 0, state 0)
    _, qe, ste = I.iss_ctx_from_counters(cfg, cnt, equal_prob=True)
    assert (qe == 128).all() and (ste == 0).all()
    # the device-side finalisation (no host round trip) gives the same doubles, side information and states
    p0d, qd, std = I.iss_ctx_from_counters_device(cfg, cnt, want_p0=True)
    assert np.array_equal(p0d.cpu().numpy(), p0) and np.array_equal(qd.cpu().numpy(), q) and np.array_equal(std.cpu().numpy(), st)
    _, qde, stde = I.iss_ctx_from_counters_device(cfg, cnt, equal_prob=True)
    assert (qde.cpu().numpy() == 128).all() and (stde.cpu().numpy() == 0).all()


def test_ctx_finalise_device_equals_host_on_random_counters():
    """counters -> p(0) -> uint8 side info -> state bytes on the device against the host function (which is pinned to the
    reference's xMapProbabilityToState): random counter tables incl. empty and degenerate contexts."""
    import isscabac_b200 as I
    rng = np.random.default_rng(31)
    for Nlbp, types in ((3, ALLT), (1, 0), (8, ALLT | O.CM_CONDBINLFT)):
        cfg = I.make_cfg(I.PROFILE_ISS, I.BIN_EG0, 8, Nlbp, types, rows=0)
        K = int(I.lib().cabac_iss_num_counters(Nlbp))
        g = 500
        tot = rng.integers(0, 5000, size=(g, K))
        cnt = (tot * rng.random((g, K))).astype(np.int64)
        cnt[rng.random((g, K)) < 0.1] = 0
        # hits never exceed their totals in real data; keep a few arbitrary rows too (the formulas must still agree)
        cnt[:400] = np.minimum(cnt[:400], tot[:400])
        p0, q, st = I.iss_ctx_from_counters(cfg, cnt)
        p0d, qd, std = I.iss_ctx_from_counters_device(cfg, torch.as_tensor(cnt, device="cuda"), want_p0=True)
        assert np.array_equal(p0d.cpu().numpy(), p0, equal_nan=True)
        assert np.array_equal(qd.cpu().numpy(), q) and np.array_equal(std.cpu().numpy(), st)


def test_cabacEncode_cabacDecode_like_the_reference(tmp_path):
    """ISS.m:192-211 (DEMO mode): encode W (400x20) and H (109x20), decode, compare; the file bytes are what
    the reference engine writes for the oracle's (bin, ctx) sequence with the same side information."""
    from isscabac_b200 import coder
    rng = np.random.default_rng(4)
    param = dict(binMethod="DEC2EG0", cmTypes=["cond0", "cond1", "conds0", "conds1"], Nlbp=3, equalProb=False)
    for name, shape in (("W", (400, 20)), ("H", (109, 20))):
        G = iss_matrix(rng, *shape)
        param["fn"] = str(tmp_path / f"{name}.bit")
        nbits, ctx0 = coder.cabacEncode(G, 8, param)
        p = O.iss_ctx_init(G, 8, O.BIN_EG0, 3, ALLT)
        q = O.matlab_uint8(p * 255)
        assert np.array_equal(ctx0, q)
        st = O.ctx_from_p0(q.astype(np.float64) / 255)
        ocfg = O.make_cfg(O.PROFILE_ISS, O.BIN_EG0, 8, 3, ALLT, shape[0])
        flat = G.T.reshape(-1).astype(np.uint32)
        ops = O.symbols_to_ops(ocfg, flat)
        slab, lens = O.encode_ops(ops, np.array([0, len(ops)], dtype=np.uint64), st, out_stride=1 << 16,
                                  impl="ref" if O.ref() is not None else "oracle")
        data = open(param["fn"], "rb").read()
        assert nbits == 8 * len(data) and data == bytes(slab[0, :lens[0]])
        assert np.array_equal(coder.cabacDecode(8, param, ctx0, G.shape), G)


def test_batched_matrices_and_column_streams():
    from isscabac_b200 import coder
    rng = np.random.default_rng(5)
    param = dict(binMethod="DEC2EG0", cmTypes=["cond0", "cond1", "conds0", "conds1"], Nlbp=3)
    mats = [iss_matrix(rng, 109, 20) for _ in range(12)] + [iss_matrix(rng, 109, 3)]
    single = [coder.encode_matrices([m], 8, param) for m in mats[:3]]
    for per_column in (False, True):
        enc = coder.encode_matrices(mats, 8, param, per_column=per_column, want_heat=True)
        dec = coder.decode_matrices(enc, [m.shape for m in mats], per_column=per_column)
        assert all(np.array_equal(a, b) for a, b in zip(dec, mats))
        boff = enc["byte_off"].cpu().numpy()
        assert int(enc["heat"].sum()) <= 8 * int(boff[-1])
        if not per_column:   # a batch member is byte-identical to the matrix coded alone
            pay = enc["payload"].cpu().numpy()
            for i, s in enumerate(single):
                assert bytes(pay[boff[i]:boff[i + 1]]) == s["payload"].cpu().numpy().tobytes()
        else:
            assert len(boff) - 1 == sum(m.shape[1] for m in mats)


def test_matrices_through_the_container():
    """encode_matrices -> container blob (payload + offsets + symbol counts + uint8 side info) -> decode;
    a single column stream cut out of the blob is decoded by the oracle with the states derived from
    the side information bytes, the way cabacDecode.m:13 derives them."""
    from isscabac_b200 import coder, container as K
    rng = np.random.default_rng(6)
    param = dict(binMethod="DEC2EG0", cmTypes=["cond0", "cond1", "conds0", "conds1"], Nlbp=3)
    mats = [iss_matrix(rng, 109, 20) for _ in range(5)]
    enc = coder.encode_matrices(mats, 8, param, per_column=True)
    blob = coder.to_container(enc)
    dec = coder.decode_matrices(coder.from_container(blob.tobytes()), [m.shape for m in mats], per_column=True)
    assert all(np.array_equal(a, b) for a, b in zip(dec, mats))
    c = K.unpack(blob)
    assert c.ctx_is_prob and c.n_streams == 100 and c.ctx_init.shape == (100, 23)
    s = 47                                                   # column 7 of matrix 2
    ocfg = O.make_cfg(O.PROFILE_ISS, O.BIN_EG0, 8, 3, ALLT, 109)
    st = O.ctx_from_p0(c.stream_ctx(s).astype(np.float64) / 255)
    one = c.stream(s)
    out, ok = O.decode_symbols(ocfg, one, np.array([0, one.size], dtype=np.uint64), np.array([0, 109], dtype=np.uint64), st)
    assert ok.all() and np.array_equal(out, mats[2][:, 7])


def test_two_level_matrices_are_coded_without_binarisation(tmp_path):
    """Nq = 2 (the default of cabacEncode.m:12): every symbol is its own single bin (cabacEncode.m:16,
    cabacDecode.m:62-66), whatever binMethod says; the bytes are what the reference engine writes for
    those bins with the ISS context rule."""
    from isscabac_b200 import coder
    rng = np.random.default_rng(8)
    G = (rng.random((60, 7)) < 0.3).astype(np.uint32)
    param = dict(binMethod="DEC2EG0", cmTypes=["cond0", "cond1", "conds0", "conds1"], Nlbp=3, fn=str(tmp_path / "b.bit"))
    nbits, ctx0 = coder.cabacEncode(G, 2, param)
    assert np.array_equal(coder.cabacDecode(2, param, ctx0, G.shape), G)
    ocfg = O.make_cfg(O.PROFILE_ISS, O.BIN_TU, 2, 3, ALLT, 60)
    flat = G.T.reshape(-1).astype(np.uint32)
    ops = O.symbols_to_ops(ocfg, flat)
    assert len(ops) == flat.size and np.array_equal(ops & 1, flat)          # one bin per symbol, the value itself
    st = O.ctx_from_p0(ctx0.astype(np.float64) / 255)
    slab, lens = O.encode_ops(ops, np.array([0, len(ops)], dtype=np.uint64), st, out_stride=4096,
                              impl="ref" if O.ref() is not None else "oracle")
    assert open(param["fn"], "rb").read() == bytes(slab[0, :lens[0]]) and nbits == 8 * int(lens[0])
