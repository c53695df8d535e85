"""GPU parity tests (run on the B200 box with -m gpu): the CUDA kernels, called through the
C ABI, against the golden vectors from the unmodified reference engine and against the oracle
on seeded random inputs.  Bar: bit-exact (all integer/byte work)."""
import json
import os

import numpy as np
import pytest

import oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def I():
    import isscabac_b200 as I
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return I


def same_rows(slab, s_ref, lens):
    """slab rows equal up to each stream's own length (bytes past it are unspecified)."""
    w = int(lens.max()) if len(lens) else 0
    live = np.arange(w)[None, :] < np.asarray(lens)[:, None]
    return bool((slab[:, :w][live] == s_ref[:, :w][live]).all())


def script_to_ops(script, wide=False):
    ep, trm = (O.OP16_EP, O.OP16_TRM) if wide else (O.OP8_EP, O.OP8_TRM)
    ops = []
    for k, a, b in script:
        if k == 0:
            ops.append((b << 1) | a)
        elif k == 1:
            ops.append((ep << 1) | a)
        elif k == 2:
            ops += [(ep << 1) | ((a >> (b - 1 - i)) & 1) for i in range(b)]
        else:
            ops.append((trm << 1) | a)
    return np.array(ops, dtype=np.uint16 if wide else np.uint8)


@pytest.fixture(params=["fused", "split", "lat"])
def enc_kernel(request, monkeypatch):
    """The op kernels have three formulations, picked by the number of tiles per SM: the wide kernels (one warp per 32
    streams), the two-warp encoder (a context warp + a coder warp per 32 streams) and the latency kernels (kernels_lat.cu:
    rows in the context slots, successor rows loaded ahead; encoder and decoder).  Run the test with each one forced."""
    monkeypatch.setenv("ISSCABAC_LAT", "1" if request.param == "lat" else "0")
    monkeypatch.setenv("ISSCABAC_ENC_SPLIT", "1" if request.param == "split" else "0")
    return request.param


def gpu_roundtrip(I, ops, off, ci, stride=None):
    enc = I.encode_ops(ops, np.asarray(off, dtype=np.int64), ci, slab_stride=stride)
    pay = I.compact(enc)
    bins, ok = I.decode_ops(pay, ops, np.asarray(off, dtype=np.int64), ci)
    torch.cuda.synchronize()
    enc.check_overflow()
    return (enc.slab.cpu().numpy(), enc.lengths.cpu().numpy().astype(np.uint32), pay.payload.cpu().numpy(),
            pay.byte_off.cpu().numpy().astype(np.uint64), bins.cpu().numpy(), ok.cpu().numpy())


@pytest.mark.parametrize("wide", [False, True])
def test_kats_batched(I, golden_dir, wide, enc_kernel):
    """All KATs (K0-K8 of SURVEY 4.1 + extras) as ONE ragged batch with per-stream context init."""
    with open(os.path.join(golden_dir, "kat.json")) as f:
        kat = json.load(f)
    names = sorted(kat)
    n_ctx = max(len(kat[k]["ctx"]) for k in names)
    streams = [script_to_ops(kat[k]["script"], wide) for k in names]
    off = np.zeros(len(names) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in streams], out=off[1:])
    ops = np.concatenate(streams) if off[-1] else np.zeros(0, dtype=streams[0].dtype)
    ci = np.ones((len(names), n_ctx), dtype=np.uint8)
    for i, k in enumerate(names):
        ci[i, :len(kat[k]["ctx"])] = kat[k]["ctx"]
    slab, lens, payload, boff, bins, ok = gpu_roundtrip(I, ops, off, ci, stride=64)
    for i, k in enumerate(names):
        assert bytes(slab[i, :lens[i]]).hex() == kat[k]["bytes"], k
        assert bytes(payload[int(boff[i]):int(boff[i + 1])]).hex() == kat[k]["bytes"], k
        if kat[k]["decoded"] is not None:
            assert ok[i] == 1, k
            assert (bins[int(off[i]):int(off[i + 1])] == (streams[i] & 1)).all(), k


@pytest.mark.parametrize("fname", ["random_ops.npz", "random_ops16.npz"])
def test_golden_random_ops(I, golden_dir, fname, enc_kernel):
    z = np.load(os.path.join(golden_dir, fname))
    ops, off = z["ops"], z["op_off"]
    for tag, ci in (("shared", z["ctx_shared"]), ("per", z["ctx_per"])):
        slab, lens, payload, boff, bins, ok = gpu_roundtrip(I, ops, off, ci, stride=512)
        assert (lens == z["lens_" + tag]).all()
        assert (payload[:len(z["payload_" + tag])] == z["payload_" + tag]).all()
        assert ok.all() and (bins == (ops & 1)).all()


def rand_ops(seed, n_streams, n_ops, n_ctx, p_ep, ragged=False):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, n_ops + 1, size=n_streams) if ragged else np.full(n_streams, n_ops)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    ctx = rng.integers(0, n_ctx, size=n)
    p1 = 0.20 + 0.10 * (np.arange(n_ctx) % 5)
    bins = (rng.random(n) < p1[ctx]).astype(np.uint8)
    code = ctx.astype(np.uint8)
    ep = rng.random(n) < p_ep
    code[ep] = O.OP8_EP
    bins[ep] = rng.integers(0, 2, size=int(ep.sum()))
    return ((code << 1) | bins).astype(np.uint8), off


@pytest.mark.parametrize("seed,n_streams,n_ops,n_ctx,p_ep,ragged", [
    (1, 30000, 64, 4, 0.25, False),     # short streams: finish() carry branch, ripple carries
    (2, 4096, 2048, 23, 0.25, False),   # config-3 mix
    (3, 3000, 3000, 23, 0.25, True),    # ragged lengths, unaligned stream starts
    (4, 1000, 4096, 3, 0.0, False),     # context only
    (5, 1000, 4096, 1, 1.0, False),     # bypass only
    (6, 129, 5000, 124, 0.1, True),     # many contexts (u8 op format limit region)
])
def test_random_vs_oracle(I, seed, n_streams, n_ops, n_ctx, p_ep, ragged, enc_kernel):
    ops, off = rand_ops(seed, n_streams, n_ops, n_ctx, p_ep, ragged)
    ci = np.random.default_rng(seed).integers(0, 126, size=n_ctx).astype(np.uint8)
    stride = ((n_ops // 4 + 80) + 15) & ~15
    slab, lens, payload, boff, bins, ok = gpu_roundtrip(I, ops, off, ci, stride=stride)
    s_ref, l_ref = O.encode_ops(ops, off, ci, out_stride=stride, n_threads=8)
    assert (lens == l_ref).all()
    assert same_rows(slab, s_ref, l_ref)
    p_ref, b_ref = O.compact(s_ref, l_ref)
    assert (boff == b_ref).all() and (payload[:len(p_ref)] == p_ref).all()
    assert ok.all() and (bins == (ops & 1)).all()


def test_many_contexts_global_path(I):
    """600 contexts: beyond the shared-memory budget -> context states live in global memory."""
    rng = np.random.default_rng(12)
    n_streams, n_ops, n_ctx = 300, 3000, 600
    off = (np.arange(n_streams + 1) * n_ops).astype(np.uint64)
    n = n_streams * n_ops
    code = rng.integers(0, n_ctx, size=n).astype(np.uint16)
    bins = (rng.random(n) < 0.2).astype(np.uint16)
    code[rng.random(n) < 0.1] = O.OP16_EP
    ops = ((code << 1) | bins).astype(np.uint16)
    ci = rng.integers(0, 126, size=(n_streams, n_ctx)).astype(np.uint8)
    slab, lens, payload, boff, b, ok = gpu_roundtrip(I, ops, off, ci, stride=1024)
    s_ref, l_ref = O.encode_ops(ops, off, ci, out_stride=1024, n_threads=8)
    assert (lens == l_ref).all() and same_rows(slab, s_ref, l_ref)
    assert ok.all() and (b == (ops & 1)).all()


def test_empty_and_tiny(I, enc_kernel):
    ci = np.array([1, 1], dtype=np.uint8)
    # zero streams
    enc = I.encode_ops(np.zeros(0, np.uint8), np.array([0], dtype=np.int64), ci, slab_stride=16)
    assert enc.lengths.numel() == 0
    # streams without ops encode to K0 = fe 80
    off = np.array([0, 0, 0, 1, 1], dtype=np.uint64)
    ops = np.array([1], dtype=np.uint8)
    slab, lens, payload, boff, bins, ok = gpu_roundtrip(I, ops, off, ci, stride=16)
    assert list(lens[[0, 1, 3]]) == [2, 2, 2]
    for i in (0, 1, 3):
        assert bytes(slab[i, :2]) == bytes.fromhex("fe80")
    assert ok.all()


def test_slab_overflow_is_reported(I):
    ops, off = rand_ops(7, 256, 4096, 1, 1.0)          # bypass: 1 bit/bin -> 514 bytes per stream
    enc = I.encode_ops(ops, off.astype(np.int64), np.array([1], dtype=np.uint8), slab_stride=256)
    torch.cuda.synchronize()
    assert int(enc.overflow[0].item()) & 1
    assert (enc.lengths.cpu().numpy() == 514).all()     # true lengths are still reported
    with pytest.raises(I.CabacError):
        enc.check_overflow()
    assert I.slab_stride_bound(4096) >= 514


def test_host_buffer_api(I):
    ops, off = rand_ops(8, 5000, 700, 23, 0.25, ragged=True)
    ci = np.full(23, 1, dtype=np.uint8)
    payload, boff = I.encode_ops_host(ops, off, ci)
    s_ref, l_ref = O.encode_ops(ops, off, ci, n_threads=8)
    p_ref, b_ref = O.compact(s_ref, l_ref)
    assert (boff == b_ref).all() and (payload == p_ref).all()
    bins, ok = I.decode_ops_host(payload, boff, ops, off, ci)
    assert ok.all() and (bins == (ops & 1)).all()
    # per-stream init + a sub-range of a larger array (offsets not starting at 0)
    cip = np.random.default_rng(3).integers(0, 126, size=(4999, 23)).astype(np.uint8)
    payload2, boff2 = I.encode_ops_host(ops, off[1:], cip)
    s2, l2 = O.encode_ops(ops, off[1:], cip, n_threads=8)
    p2, b2 = O.compact(s2, l2)
    assert (boff2 == b2).all() and (payload2 == p2).all()


def test_corrupt_stream_fails_finish_check(I):
    ops, off = rand_ops(9, 64, 500, 4, 0.2)
    ci = np.full(4, 1, dtype=np.uint8)
    enc = I.encode_ops(ops, off.astype(np.int64), ci, slab_stride=256)
    pay = I.compact(enc)
    pay.payload[int(pay.byte_off[1].item()) - 1] ^= 0x40      # damage the last byte of stream 0
    bins, ok = I.decode_ops(pay, ops, off.astype(np.int64), ci)
    okc = ok.cpu().numpy()
    assert okc[0] == 0 and okc[1:].all()


def test_full_length_streams_roundtrip(I, enc_kernel):
    """BASELINE size per stream (65,536 bins), fewer streams: size-independent properties --
    round trip, every finish() check, and byte parity of a sample against the oracle."""
    ops, off = rand_ops(10, 2048, 65536, 23, 0.25)
    ci = np.full(23, 1, dtype=np.uint8)
    slab, lens, payload, boff, bins, ok = gpu_roundtrip(I, ops, off, ci, stride=16448)
    assert ok.all() and (bins == (ops & 1)).all()
    s_ref, l_ref = O.encode_ops(ops[:64 * 65536], off[:65], ci, out_stride=16448, n_threads=8)
    assert (lens[:64] == l_ref).all() and same_rows(slab[:64], s_ref, l_ref)


def test_out_of_range_codes_are_bypass_bins(I, enc_kernel):
    """Codes >= n_ctx (other than terminate) are bypass bins in every kernel formulation: wide / two-warp / latency kernels
    (u8, shared-memory contexts), the general kernels (u16 ops, many contexts), n_ctx == 0 included."""
    rng = np.random.default_rng(43)
    for n_ctx, wide16 in ((5, False), (0, False), (23, True), (300, True)):
        n_streams, n_ops = 100, 500
        off = (np.arange(n_streams + 1) * n_ops).astype(np.int64)
        hi = 0x7FFD if wide16 else 125
        code = rng.integers(0, max(n_ctx, 1), size=n_streams * n_ops).astype(np.uint32)
        bad = rng.random(len(code)) < (0.1 if n_ctx else 1.0)
        code[bad] = rng.integers(n_ctx, hi, size=int(bad.sum()))
        bins = (rng.random(len(code)) < 0.4).astype(np.uint32)
        ops = ((code << 1) | bins).astype(np.uint16 if wide16 else np.uint8)
        ci = rng.integers(0, 126, size=max(n_ctx, 0)).astype(np.uint8)
        slab, lens, payload, boff, dbins, ok = gpu_roundtrip(I, ops, off, ci, stride=512)
        s_ref, l_ref = O.encode_ops(ops, off.astype(np.uint64), ci, out_stride=512, n_threads=4)
        assert (lens == l_ref).all() and same_rows(slab, s_ref, l_ref), (n_ctx, wide16)
        assert ok.all() and (dbins == bins).all(), (n_ctx, wide16)


@pytest.mark.parametrize("ragged", [False, True])
def test_host_buffer_api_packed_bins(I, ragged):
    """cabac_decode_ops_host_packed: the decoded bins one per BIT (little-endian within a byte) equal the one-per-byte
    result; ragged streams put the chunk boundaries of the pipeline off the byte grid (packed after the last chunk then)."""
    rng = np.random.default_rng(47)
    n_streams = 3000
    lens = rng.integers(0, 900, size=n_streams) if ragged else np.full(n_streams, 512)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    code = rng.integers(0, 7, size=n).astype(np.uint8)
    code[rng.random(n) < 0.2] = O.OP8_EP
    bins = (rng.random(n) < 0.35).astype(np.uint8)
    ops = ((code << 1) | bins).astype(np.uint8)
    ci = rng.integers(0, 126, size=7).astype(np.uint8)
    payload, boff = I.encode_ops_host(ops, off, ci)
    s_ref, l_ref = O.encode_ops(ops, off, ci, out_stride=512, n_threads=4)
    p_ref, b_ref = O.compact(s_ref, l_ref)
    assert (boff == b_ref).all() and (payload == p_ref).all()
    b8, ok8 = I.decode_ops_host(payload, boff, ops, off, ci)
    bp, okp = I.decode_ops_host(payload, boff, ops, off, ci, packed=True)
    assert ok8.all() and okp.all() and (b8 == bins).all()
    assert bp.size == (n + 7) // 8
    assert (np.unpackbits(bp, bitorder="little")[:n] == bins).all()
    if n % 8:    # padding bits of the last byte are zero
        assert (np.unpackbits(bp, bitorder="little")[n:] == 0).all()


def test_worst_case_window_growth_on_device(I, enc_kernel):
    """Every bin an LPS on a context parked at state 60..62 -- six shift bits per bin, the most the tables allow -- in FULL
    warps of equally long streams, i.e. through the lockstep 16-op blocks with their voted / unvoted emissions and top-ups:
    the encoder's early-emit guard and the decoder's look-ahead budget at their limits (the host emulation runs the same
    input through the per-lane schedule, tests/test_wide_emulation.py::test_wide_worst_case_growth).  A mix of worst-case
    and ordinary streams in one warp as well: the votes see both."""
    n_ctx, n_ops, n_streams = 120, 4000, 96
    rng = np.random.default_rng(61)
    ctx_seq = (np.arange(n_ops) % n_ctx).astype(np.uint8)
    off = (np.arange(n_streams + 1) * n_ops).astype(np.uint64)
    for state in (62, 61, 60):
        ci = np.zeros((n_streams, n_ctx), dtype=np.uint8)
        ops = np.zeros(n_streams * n_ops, dtype=np.uint8)
        for s in range(n_streams):
            hot = s < 64 or s % 3 == 0          # two warps of worst-case streams, a third warp with a few among ordinary ones
            ci[s] = ((state << 1) | 1) if hot else rng.integers(0, 126, size=n_ctx)
            bins = np.zeros(n_ops, dtype=np.uint8) if hot else (rng.random(n_ops) < 0.4).astype(np.uint8)
            ops[s * n_ops:(s + 1) * n_ops] = (ctx_seq << 1) | bins
        s_ref, l_ref = O.encode_ops(ops, off, ci, out_stride=8192, n_threads=8)
        slab, lens, payload, boff, bins_g, ok = gpu_roundtrip(I, ops, off, ci, stride=8192)
        assert (lens == l_ref).all(), state
        live = np.arange(8192)[None, :] < l_ref[:, None]
        assert (slab[live] == s_ref[live]).all(), state
        assert ok.all() and (bins_g == (ops & 1)).all(), state


@pytest.mark.parametrize("warps,ragged,partial", [(6, False, False), (6, True, False), (10, False, True), (10, True, False), (14, False, False)])
def test_encoder_handover_between_schedulers(I, monkeypatch, warps, ragged, partial):
    """k_encode_ops_wide_ho: with 4k + 2 tiles per SM the fourth warps of two schedulers hand their tile over, half-coded, to
    warps that slept on the other two (lane state through shared memory).  A job of exactly that geometry -- one CTA per SM,
    `warps` tiles each; ragged lengths; a last CTA whose giving warps have no or partial tiles -- against the oracle for a
    stream sample taken from every warp role, and against the plain kernel for ALL streams."""
    monkeypatch.setenv("ISSCABAC_LAT", "0")
    monkeypatch.setenv("ISSCABAC_ENC_SPLIT", "0")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n_streams = sms * warps * 32 - (warps * 32 + 40 if partial else 0)      # partial: the last CTA loses a warp's tiles and a bit
    n_ops = 1500
    assert I.lib().cabac_encode_ops_kernel(n_streams, 23) == b"k_encode_ops_wide_ho"
    ops, off = rand_ops(11 + warps, n_streams, n_ops, 23, 0.25, ragged)
    if ragged:      # keep every stream long enough for the hand-over to take place (>= 64 common blocks)
        lens = np.maximum(np.diff(off.astype(np.int64)), 1100)
        off = np.zeros(n_streams + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        ops = rand_ops(12 + warps, 1, int(off[-1]), 23, 0.25, False)[0]
    ci = np.random.default_rng(5).integers(0, 126, size=23).astype(np.uint8)
    stride = ((n_ops // 4 + 80) + 15) & ~15
    d_ops, d_off = torch.as_tensor(ops, device="cuda"), torch.as_tensor(off.astype(np.int64), device="cuda")
    enc = I.encode_ops(d_ops, d_off, ci, slab_stride=stride)
    monkeypatch.setenv("ISSCABAC_HANDOVER", "0")
    assert I.lib().cabac_encode_ops_kernel(n_streams, 23) == b"k_encode_ops_wide"
    plain = I.encode_ops(d_ops, d_off, ci, slab_stride=stride)
    torch.cuda.synchronize()
    enc.check_overflow()
    lens_g, lens_p = enc.lengths.cpu().numpy(), plain.lengths.cpu().numpy()
    assert (lens_g == lens_p).all()
    assert same_rows(enc.slab.cpu().numpy(), plain.slab.cpu().numpy(), lens_p)
    # the oracle on streams of every warp of the first, a middle and the last CTA (givers are the last two warps of a CTA)
    ids = np.unique(np.concatenate([np.arange(c * warps * 32, min((c + 1) * warps * 32, n_streams))[::7] for c in (0, sms // 2, sms - 1)]))
    sub_off = np.zeros(len(ids) + 1, dtype=np.uint64)
    np.cumsum([int(off[i + 1] - off[i]) for i in ids], out=sub_off[1:])
    sub_ops = np.concatenate([ops[int(off[i]):int(off[i + 1])] for i in ids])
    s_ref, l_ref = O.encode_ops(sub_ops, sub_off, ci, out_stride=stride, n_threads=8)
    assert (lens_g[ids].astype(np.uint32) == l_ref).all()
    assert same_rows(enc.slab[torch.as_tensor(ids, device="cuda")].cpu().numpy(), s_ref, l_ref)
    # ... and back: the hand-over decoder where it applies (more than 8 tiles per SM), bins and finish() flags of all streams
    monkeypatch.setenv("ISSCABAC_HANDOVER", "1")
    assert I.lib().cabac_decode_ops_kernel(n_streams, 23) == (b"k_decode_ops_wide_ho" if warps > 8 else b"k_decode_ops_wide")
    bins, ok = I.decode_ops(I.compact(enc), d_ops, d_off, ci)
    assert bool(ok.all().item()) and bool((bins == (d_ops & 1)).all().item())
