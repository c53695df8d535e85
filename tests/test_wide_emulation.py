"""CPU: the wide-window lane formulation (isscabac_b200/csrc/cabac_wide.cuh -- the code the hot
kernels execute: 64-bit low window with word-wise emission, 64-bit value window with word-wise
refill, bypass folded into the context step) compiled for the host, driven with the kernels'
block schedule and checked bit-exactly against the golden vectors and the oracle."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle as O
from test_lane_emulation import emul, p, script_to_ops, u8p, u32p, u64p  # noqa: F401


# the formulation under test: "wide" (cabac_wide.cuh, throughput kernels) or "spec" (cabac_spec.cuh, latency kernels);
# every test of the block schedule runs with both
FORM = ["wide"]


@pytest.fixture(params=["wide", "spec"], autouse=True)
def _formulation(request):
    FORM[0] = request.param
    yield
    FORM[0] = "wide"


def wide_encode(L, ops, off, ci, stride, misalign=0, form=None):
    """misalign: byte offset of the op array inside a 64-byte aligned buffer (exercises head/tail)."""
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    buf = np.zeros(len(ops) + 128, dtype=np.uint8)
    base = (-buf.ctypes.data) % 64 + misalign
    buf[base:base + len(ops)] = ops
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = len(off) - 1
    ci = np.ascontiguousarray(ci, dtype=np.uint8)
    per = int(ci.ndim == 2)
    slab = np.zeros((n, stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint32)
    fn = L.emul_encode_ops_spec if (form or FORM[0]) == "spec" else L.emul_encode_ops_wide
    ovf = fn(C.c_uint32(n), p(off, u64p), C.c_void_p(buf.ctypes.data + base), p(ci.reshape(-1), u8p),
                                 C.c_uint32(ci.shape[-1]), per, p(slab, u8p), C.c_uint64(stride), p(lens, u32p))
    assert ovf == 0
    return slab, lens


def wide_decode(L, payload, boff, ops, off, ci, misalign=0, pay_misalign=0, form=None):
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    buf = np.zeros(len(ops) + 128, dtype=np.uint8)
    base = (-buf.ctypes.data) % 64 + misalign
    buf[base:base + len(ops)] = ops
    out = np.zeros(len(ops) + 128, dtype=np.uint8)
    obase = (-out.ctypes.data) % 64 + misalign
    pb = np.full(len(payload) + 128, 0xA5, dtype=np.uint8)   # junk around the payload: must never influence decoding
    pbase = (-pb.ctypes.data) % 64 + pay_misalign
    pb[pbase:pbase + len(payload)] = payload
    ci = np.ascontiguousarray(ci, dtype=np.uint8)
    per = int(ci.ndim == 2)
    n = len(off) - 1
    ok = np.zeros(n, dtype=np.uint8)
    fn = L.emul_decode_ops_spec if (form or FORM[0]) == "spec" else L.emul_decode_ops_wide
    fn(C.c_uint32(n), p(np.ascontiguousarray(boff, dtype=np.uint64), u64p), C.c_void_p(pb.ctypes.data + pbase),
                           p(np.ascontiguousarray(off, dtype=np.uint64), u64p), C.c_void_p(buf.ctypes.data + base),
                           p(ci.reshape(-1), u8p), C.c_uint32(ci.shape[-1]), per, C.c_void_p(out.ctypes.data + obase), p(ok, u8p))
    return out[obase:obase + len(ops)].copy(), ok


def test_wide_kats(emul, golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        kat = json.load(f)
    for name, k in kat.items():
        ops = script_to_ops(k["script"])
        off = np.array([0, len(ops)], dtype=np.uint64)
        ci = np.array(k["ctx"], dtype=np.uint8)
        if len(ci) == 0:
            ci = np.array([1], dtype=np.uint8)
        for mis in (0, 5):
            slab, lens = wide_encode(emul, ops, off, ci, 64, misalign=mis)
            assert bytes(slab[0, :lens[0]]).hex() == k["bytes"], name
            if k["decoded"] is not None:
                data = np.frombuffer(bytes.fromhex(k["bytes"]), dtype=np.uint8)
                for pm in (0, 1, 2, 3):
                    bins, ok = wide_decode(emul, data, [0, len(data)], ops, off, ci, misalign=mis, pay_misalign=pm)
                    assert ok[0] == 1 and (bins == (ops & 1)).all(), (name, mis, pm)


def test_wide_random_golden(emul, golden_dir):
    z = np.load(os.path.join(golden_dir, "random_ops.npz"))
    ops, off = z["ops"], z["op_off"]
    for tag, ci in (("shared", z["ctx_shared"]), ("per", z["ctx_per"])):
        slab, lens = wide_encode(emul, ops, off, ci, 512, misalign=3)
        assert (lens == z["lens_" + tag]).all()
        payload, boff = O.compact(slab, lens)
        assert (payload == z["payload_" + tag]).all()
        bins, ok = wide_decode(emul, payload, boff, ops, off, ci, misalign=3, pay_misalign=1)
        assert ok.all() and (bins == (ops & 1)).all()


def _mixed_ops(rng, n, n_ctx, p_ep, p1):
    code = rng.integers(0, n_ctx, size=n).astype(np.uint8)
    bins = (rng.random(n) < p1).astype(np.uint8)
    ep = rng.random(n) < p_ep
    code[ep] = O.OP8_EP
    bins[ep] = rng.integers(0, 2, size=int(ep.sum()))
    return ((code << 1) | bins).astype(np.uint8)


def test_wide_vs_oracle_bulk_short_streams(emul):
    # ~3M bins in short ragged streams: carries into the pending word, finish() carries, every head/tail length
    rng = np.random.default_rng(19)
    n_streams = 40000
    lens_ops = rng.integers(0, 150, size=n_streams)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(lens_ops, out=off[1:])
    ops = _mixed_ops(rng, int(off[-1]), 4, 0.25, 0.15)
    ci = np.full(4, 1, dtype=np.uint8)
    s1, l1 = wide_encode(emul, ops, off, ci, 64)
    s0, l0 = O.encode_ops(ops, off, ci, out_stride=64, n_threads=8)
    assert (l0 == l1).all()
    live = np.arange(64)[None, :] < l0[:, None]
    assert (s0[live] == s1[live]).all()
    payload, boff = O.compact(s0, l0)
    bins, ok = wide_decode(emul, payload, boff, ops, off, ci)
    assert ok.all() and (bins == (ops & 1)).all()


def test_wide_long_streams_all_mixes(emul):
    rng = np.random.default_rng(23)
    for n_ctx, p_ep, p1 in ((23, 0.25, 0.3), (3, 0.0, 0.05), (1, 1.0, 0.5), (124, 0.1, 0.5), (2, 0.0, 0.5)):
        n_streams, n_ops = 64, 20000
        off = (np.arange(n_streams + 1) * n_ops).astype(np.uint64)
        ops = _mixed_ops(rng, n_streams * n_ops, n_ctx, p_ep, p1)
        ci = rng.integers(0, 126, size=(n_streams, n_ctx)).astype(np.uint8)
        stride = n_ops // 4 + 4096
        s1, l1 = wide_encode(emul, ops, off, ci, stride, misalign=7)
        s0, l0 = O.encode_ops(ops, off, ci, out_stride=stride, n_threads=8)
        assert (l0 == l1).all()
        live = np.arange(stride)[None, :] < l0[:, None]
        assert (s0[live] == s1[live]).all()
        payload, boff = O.compact(s0, l0)
        bins, ok = wide_decode(emul, payload, boff, ops, off, ci, misalign=7, pay_misalign=2)
        assert ok.all() and (bins == (ops & 1)).all()


def test_wide_worst_case_growth(emul):
    """Every bin an LPS on a context parked at state 60..62 (6 shift bits per bin): the early-emit
    guard of the 16-op block and the decoder's look-ahead budget at their limits.  State 63 (lps = 2,
    renorm shift 6 by table, CABAC_ArithmeticEncoder.cpp:479,484) is exercised with MPS bins and one
    closing LPS: after an LPS at state 63 the range is 128 and the reference decoder indexes its
    table out of bounds (CABAC_ArithmeticDecoder.cpp:94), so nothing may follow it but finish()."""
    n_ctx, n_ops = 120, 4000
    ctx_seq = np.arange(n_ops) % n_ctx
    off = np.array([0, n_ops], dtype=np.uint64)
    stride = 8192
    for state in (62, 61, 60, 63):
        ci = np.full(n_ctx, (state << 1) | 1, dtype=np.uint8)      # mps = 1
        ops = ((ctx_seq << 1) | 0).astype(np.uint8)                # bin 0 = LPS
        if state == 63:
            ops[:-1] |= 1                                          # MPS bins, then a single LPS
        s1, l1 = wide_encode(emul, ops, off, ci, stride)
        s0, l0 = O.encode_ops(ops, off, ci, out_stride=stride, n_threads=1)
        assert l0[0] == l1[0] and (s0[0, :l0[0]] == s1[0, :l0[0]]).all(), state
        bins, ok = wide_decode(emul, s0[0, :l0[0]], [0, int(l0[0])], ops, off, ci)
        assert (bins == (ops & 1)).all() and (ok[0] == 1 or state == 63), state


def test_wide_terminate_bins_mid_stream(emul):
    rng = np.random.default_rng(29)
    n_streams, n_ops = 500, 400
    off = (np.arange(n_streams + 1) * n_ops).astype(np.uint64)
    ops = _mixed_ops(rng, n_streams * n_ops, 5, 0.2, 0.3)
    trm = rng.random(len(ops)) < 0.02
    ops[trm] = (O.OP8_TRM << 1) | 0          # encodeBinTrm(0) inside the stream
    ci = rng.integers(0, 126, size=5).astype(np.uint8)
    s1, l1 = wide_encode(emul, ops, off, ci, 256, misalign=9)
    s0, l0 = O.encode_ops(ops, off, ci, out_stride=256, n_threads=8)
    assert (l0 == l1).all()
    live = np.arange(256)[None, :] < l0[:, None]
    assert (s0[live] == s1[live]).all()
    payload, boff = O.compact(s0, l0)
    bins, ok = wide_decode(emul, payload, boff, ops, off, ci, misalign=9, pay_misalign=3)
    assert ok.all() and (bins == (ops & 1)).all()


def test_wide_decoder_reads_ff_past_the_end(emul):
    """A truncated stream: the reference's readByte() yields 0xFF after EOF
    (CABAC_BitstreamFile.cpp:153-158); the wide decoder must produce the same bins."""
    rng = np.random.default_rng(31)
    n_ops = 3000
    ops = _mixed_ops(rng, n_ops, 6, 0.3, 0.4)
    off = np.array([0, n_ops], dtype=np.uint64)
    ci = np.full(6, 1, dtype=np.uint8)
    s0, l0 = O.encode_ops(ops, off, ci, out_stride=2048, n_threads=1)
    for cut in (0, 1, 2, 3, 5, 100, int(l0[0]) - 1):
        data = s0[0, :cut].copy()
        want, _ = O.decode_ops(data if cut else np.zeros(1, np.uint8), np.array([0, cut], dtype=np.uint64), ops, off, ci, n_threads=1)
        got, _ = wide_decode(emul, data, [0, cut], ops, off, ci, pay_misalign=cut % 4)
        assert (got == want).all(), cut


def test_wide_carry_walk(emul):
    """The carry past a 0xFFFFFFFF pending word (needs 32 one-bits in a row at a word boundary, so
    no random stream reaches it): +1 ripples through the big-endian words already stored."""
    def be(words):
        return np.array(words, dtype=">u4").view(np.uint8).view("<u4").copy()
    row = be([0x12345678, 0x00FFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xAAAAAAAA])
    emul.emul_carry_walk(p(row, u32p), C.c_uint32(4), C.c_uint32(16))      # pending word index 4: words 0..3 are stored
    assert list(row.view(np.uint8).view(">u4")) == [0x12345678, 0x01000000, 0, 0, 0xAAAAAAAA]
    row = be([0xFFFFFFFF, 0xFFFFFFFF, 7])
    emul.emul_carry_walk(p(row, u32p), C.c_uint32(2), C.c_uint32(1))       # capacity 1 word: word 1 is not in memory
    assert list(row.view(np.uint8).view(">u4")) == [0, 0xFFFFFFFF, 7]


@pytest.mark.parametrize("max_run", [1, 2, 7, 16])
def test_wide_bypass_runs(emul, max_run):
    """encodeBinsEP / decodeBinsEP as ONE step of the wide formulation (encw_ep_run: low' = (low << n) +
    range * bits; decw_ep_run: long division of the window by range) give the bytes / bins of n single
    bypass steps -- the reference's own equivalence, here against the oracle."""
    rng = np.random.default_rng(31 + max_run)
    n_streams, n_ctx = 300, 5
    lens = rng.integers(0, 900, size=n_streams)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    code = rng.integers(0, n_ctx, size=n).astype(np.uint8)
    # long runs of bypass ops between context-coded stretches, bins of either skew
    run = np.repeat(rng.random(n // 8 + 1) < 0.55, 8)[:n]
    code[run] = O.OP8_EP
    bins = (rng.random(n) < rng.choice([0.1, 0.5, 0.9], size=n)).astype(np.uint8)
    ops = ((code << 1) | bins).astype(np.uint8)
    ci = rng.integers(0, 126, size=(n_streams, n_ctx)).astype(np.uint8)
    slab_ref, lens_ref = O.encode_ops(ops, off, ci, out_stride=1024)
    slab = np.zeros((n_streams, 1024), dtype=np.uint8)
    ln = np.zeros(n_streams, dtype=np.uint32)
    ovf = emul.emul_encode_ops_wide_runs(C.c_uint32(n_streams), p(off, u64p), p(ops, u8p), p(ci.reshape(-1), u8p),
                                         C.c_uint32(n_ctx), 1, p(slab, u8p), C.c_uint64(1024), p(ln, u32p), C.c_uint32(max_run))
    assert ovf == 0 and (ln == lens_ref).all()
    live = np.arange(1024)[None, :] < lens_ref[:, None]
    assert (slab[live] == slab_ref[live]).all()
    payload, boff = O.compact(slab_ref, lens_ref)
    pb = np.concatenate([payload, np.full(64, 0xA5, np.uint8)])
    out = np.zeros(n + 1, dtype=np.uint8)
    ok = np.zeros(n_streams, dtype=np.uint8)
    emul.emul_decode_ops_wide_runs(C.c_uint32(n_streams), p(np.ascontiguousarray(boff, dtype=np.uint64), u64p), p(pb, u8p),
                                   p(off, u64p), p(ops, u8p), p(ci.reshape(-1), u8p), C.c_uint32(n_ctx), 1,
                                   p(out, u8p), p(ok, u8p), C.c_uint32(max_run))
    assert ok.all() and (out[:n] == bins).all()
    # the tree decoder's form of a run: restoring division on the upper word of the window (decw_ep_bits)
    out[:] = 0
    ok[:] = 0
    emul.emul_decode_ops_wide_runs_loop(C.c_uint32(n_streams), p(np.ascontiguousarray(boff, dtype=np.uint64), u64p), p(pb, u8p),
                                        p(off, u64p), p(ops, u8p), p(ci.reshape(-1), u8p), C.c_uint32(n_ctx), 1,
                                        p(out, u8p), p(ok, u8p), C.c_uint32(max_run))
    assert ok.all() and (out[:n] == bins).all()


def test_out_of_range_codes_are_bypass_bins(emul):
    """Op format: a code >= n_ctx that is not the terminate code is coded as a BYPASS bin (include/isscabac.h) -- the same
    in every formulation and in both directions, so a malformed op array cannot make encoder and decoder disagree."""
    from test_lane_emulation import emul_decode, emul_encode
    rng = np.random.default_rng(41)
    n_ctx, n_streams, n_ops = 5, 40, 700
    off = (np.arange(n_streams + 1) * n_ops).astype(np.uint64)
    code = rng.integers(0, n_ctx, size=n_streams * n_ops).astype(np.uint8)
    bad = rng.random(len(code)) < 0.1
    code[bad] = rng.integers(n_ctx, 125, size=int(bad.sum()))        # not contexts of this call, not TRM, not EP
    ops = ((code << 1) | (rng.random(len(code)) < 0.4)).astype(np.uint8)
    as_ep = ops.copy()
    as_ep[bad] = (O.OP8_EP << 1) | (ops[bad] & 1)
    ci = rng.integers(0, 126, size=n_ctx).astype(np.uint8)
    s_ref, l_ref = O.encode_ops(as_ep, off, ci, out_stride=512)      # the same ops with the bad codes written as bypass ops
    s_o, l_o = O.encode_ops(ops, off, ci, out_stride=512)
    assert (l_o == l_ref).all() and (s_o == s_ref).all()
    s1, l1 = wide_encode(emul, ops, off, ci, 512, misalign=3)
    s2, l2 = emul_encode(emul, ops, off, ci, 512)
    live = np.arange(512)[None, :] < l_ref[:, None]
    assert (l1 == l_ref).all() and (s1[live] == s_ref[live]).all()
    assert (l2 == l_ref).all() and (s2[live] == s_ref[live]).all()
    payload, boff = O.compact(s_ref, l_ref)
    for bins, ok in (wide_decode(emul, payload, boff, ops, off, ci, misalign=3), emul_decode(emul, payload, boff, ops, off, ci),
                     O.decode_ops(payload, boff, ops, off, ci)):
        assert ok.all() and (bins == (ops & 1)).all()
