"""Container / wire format: payload + offset table + unit counts + context
bytes in one buffer; every stream cut out of it is byte-for-byte the file the reference would have
written, i.e. decodable by the UNMODIFIED reference engine (oracle/_ref).  CPU tests use the oracle
as the producer of the streams; the GPU test packs what the CUDA encoder produced."""
import zlib

import numpy as np
import pytest

import oracle as O


def _job(seed=3, n_streams=37, n_ctx=5):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 300, size=n_streams)
    lens[5] = 0                                      # an empty stream (start(); finish() only)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    code = rng.integers(0, n_ctx, size=n).astype(np.uint8)
    code[rng.random(n) < 0.2] = O.OP8_EP
    ops = ((code << 1) | (rng.random(n) < 0.3)).astype(np.uint8)
    ci = rng.integers(0, 126, size=(n_streams, n_ctx)).astype(np.uint8)
    return ops, off, ci


def test_pack_unpack_round_trip_and_reference_decodes_a_cut_out_stream():
    from isscabac_b200 import container as K
    ops, off, ci = _job()
    slab, lens = O.encode_ops(ops, off, ci, out_stride=256)
    payload, boff = O.compact(slab, lens)
    blob = K.pack(payload, boff, ci, unit_off=off)
    assert blob[:8].tobytes() == b"ISSCABAC" and blob.size % 8 == 0
    c = K.unpack(blob.tobytes())
    assert c.n_streams == len(lens) and c.cfg is None and not c.ctx_is_prob
    assert (c.byte_off == boff).all() and (c.unit_off == off).all()
    assert (c.payload == payload).all() and (c.ctx_init == ci).all()
    # any single stream, cut out, is a bitstream the reference engine decodes (decodeStart..decodeFinish)
    impl = "ref" if O.ref() is not None else "oracle"
    for s in (0, 5, 17, len(lens) - 1):
        one = c.stream(s)
        a, b = int(off[s]), int(off[s + 1])
        bins, ok = O.decode_ops(one, np.array([0, one.size], dtype=np.uint64), ops[a:b],
                                np.array([0, b - a], dtype=np.uint64), c.stream_ctx(s), impl=impl)
        assert ok.all() and (bins == (ops[a:b] & 1)).all()
        assert one.tobytes() == slab[s, :lens[s]].tobytes()


def test_symbol_level_header_and_shared_ctx():
    from isscabac_b200 import container as K
    import isscabac_b200 as I
    cfg = O.make_cfg(O.PROFILE_ISS, O.BIN_EG0, 8, 3, O.CM_COND0 | O.CM_COND1 | O.CM_CONDS0 | O.CM_CONDS1, 50)
    rng = np.random.default_rng(0)
    sym = rng.integers(0, 8, size=400).astype(np.uint8)
    soff = np.array([0, 100, 400], dtype=np.uint64)
    q = np.full(23, 128, dtype=np.uint8)                        # uint8 side info, p0 = 128/255
    st = O.ctx_from_p0(q / 255.0)
    slab, lens = O.encode_symbols(cfg, sym, soff, st, 1024)
    payload, boff = O.compact(slab, lens)
    pcfg = I.make_cfg(I.PROFILE_ISS, I.BIN_EG0, 8, 3, I.CM_COND0 | I.CM_COND1 | I.CM_CONDS0 | I.CM_CONDS1, rows=50)
    blob = K.pack(payload, boff, q, unit_off=soff, cfg=pcfg, sym_width=1, ctx_is_prob=True)
    c = K.unpack(blob)
    assert c.ctx_is_prob and c.sym_width == 1 and c.ctx_init.shape == (23,)
    assert (c.cfg.profile, c.cfg.method, c.cfg.Nq, c.cfg.Nlbp, c.cfg.types, c.cfg.rows) == \
        (pcfg.profile, pcfg.method, 8, 3, pcfg.types, 50)
    # decoder side: side info -> states (cabacDecode.m:13), streams -> symbols
    st2 = O.ctx_from_p0(c.ctx_init / 255.0)
    out, ok = O.decode_symbols(cfg, c.payload, c.byte_off, c.unit_off, st2)
    assert ok.all() and (out == sym).all()


def test_crc_is_zlib_crc32():
    from isscabac_b200 import container as K
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 8, 9, 1000, 4099):
        d = rng.integers(0, 256, size=n).astype(np.uint8)
        assert K.crc32(d) == zlib.crc32(d.tobytes())
        assert K.crc32(d[1:]) == zlib.crc32(d[1:].tobytes())      # unaligned start


def test_corruption_is_detected():
    from isscabac_b200 import container as K
    import isscabac_b200 as I
    ops, off, ci = _job(seed=4, n_streams=9)
    slab, lens = O.encode_ops(ops, off, ci, out_stride=256)
    payload, boff = O.compact(slab, lens)
    blob = K.pack(payload, boff, ci, unit_off=off)
    K.unpack(blob)
    for pos in (3, 17, 60, 130, 128 + 8 * 10 + 3, blob.size - 9):   # magic, n_streams, payload size, tables, payload
        bad = blob.copy()
        bad[pos] ^= 0x40
        with pytest.raises(I.CabacError):
            K.unpack(bad)
    with pytest.raises(I.CabacError):
        K.unpack(blob[: blob.size - 8])                            # truncated
    # a flipped payload byte passes when the payload CRC is not requested (tables still verified)
    bad = blob.copy()
    bad[blob.size - 9] ^= 1
    K.unpack(bad, verify_payload_crc=False)
    # writer-side validation
    with pytest.raises(I.CabacError):
        K.pack(payload, boff[::-1].copy(), ci)


def test_crafted_header_cannot_reach_outside_the_buffer():
    """A header whose sizes were chosen to wrap the 64-bit layout arithmetic, with its checksum recomputed (the header
    CRC is no defence): parse must refuse before any CRC pass or pointer derivation, with and without the payload CRC."""
    from isscabac_b200 import container as K
    import isscabac_b200 as I
    ops, off, ci = _job(seed=6, n_streams=9)
    slab, lens = O.encode_ops(ops, off, ci, out_stride=256)
    payload, boff = O.compact(slab, lens)
    blob = K.pack(payload, boff, ci, unit_off=off)

    def forge(payload_bytes=None, n_streams=None, total=None):
        bad = blob.copy()
        if payload_bytes is not None:
            bad[56:64] = np.frombuffer(np.uint64(payload_bytes).tobytes(), np.uint8)
        if n_streams is not None:
            bad[16:20] = np.frombuffer(np.uint32(n_streams).tobytes(), np.uint8)
        if total is not None:
            bad[96:104] = np.frombuffer(np.uint64(total).tobytes(), np.uint8)
        bad[112:116] = np.frombuffer(np.uint32(zlib.crc32(bad[:112].tobytes())).tobytes(), np.uint8)
        return bad

    off_payload = int(np.frombuffer(blob[88:96].tobytes(), np.uint64)[0])
    for verify in (True, False):
        # off_payload + payload_bytes wraps to a small total that would pass a "total <= n" test
        wrap = (1 << 64) - off_payload + 64
        for bad in (forge(payload_bytes=wrap, total=64), forge(payload_bytes=(1 << 64) - 8), forge(payload_bytes=blob.size),
                    forge(n_streams=0xFFFFFFFF), forge(n_streams=1 << 28)):
            with pytest.raises(I.CabacError):
                K.unpack(bad, verify_payload_crc=verify)
    # a unit table that does not start at 0 is refused as well (tables CRC recomputed)
    bad = blob.copy()
    uo = int(np.frombuffer(blob[72:80].tobytes(), np.uint64)[0])
    bad[uo] = 1
    bad[104:108] = np.frombuffer(np.uint32(zlib.crc32(bad[128:off_payload].tobytes())).tobytes(), np.uint8)
    bad[112:116] = np.frombuffer(np.uint32(zlib.crc32(bad[:112].tobytes())).tobytes(), np.uint8)
    with pytest.raises(I.CabacError):
        K.unpack(bad)
    # writer side: a payload shorter than the offset table claims
    with pytest.raises(ValueError):
        K.pack(payload[:-3], boff, ci)


def test_empty_container():
    from isscabac_b200 import container as K
    blob = K.pack(np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros(3, np.uint8))
    c = K.unpack(blob)
    assert c.n_streams == 0 and c.payload.size == 0 and c.ctx_init.shape == (3,)


@pytest.mark.gpu
def test_gpu_encode_pack_unpack_decode():
    import torch
    import isscabac_b200 as I
    from isscabac_b200 import container as K
    ops, off, ci = _job(seed=9, n_streams=300)
    enc = I.encode_ops(ops, off.astype(np.int64), ci, slab_stride=256)
    pay = I.compact(enc)
    torch.cuda.synchronize()
    blob = K.pack(pay.payload, pay.byte_off, ci, unit_off=off)
    c = K.unpack(blob)
    # the oracle (and through it the reference) decodes a cut-out stream of the GPU encoder
    s = 123
    a, b = int(off[s]), int(off[s + 1])
    one = c.stream(s)
    bins, ok = O.decode_ops(one, np.array([0, one.size], dtype=np.uint64), ops[a:b],
                            np.array([0, b - a], dtype=np.uint64), c.stream_ctx(s))
    assert ok.all() and (bins == (ops[a:b] & 1)).all()
    # and the GPU decoder reads the whole container back
    dbins, dok = I.decode_ops((c.payload, c.byte_off.astype(np.int64)), ops, c.unit_off.astype(np.int64), c.ctx_init)
    assert bool(dok.all().item()) and (dbins.cpu().numpy() == (ops & 1)).all()
