"""GPU parity tests for the symbol-level kernels (binarizer + context selection + coder)."""
import os

import numpy as np
import pytest

import oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ALLT = O.CM_COND0 | O.CM_COND1 | O.CM_CONDS0 | O.CM_CONDS1


@pytest.fixture(scope="module")
def I():
    import isscabac_b200 as I
    assert torch.cuda.is_available()
    return I


def test_golden_symbol_cases(I, golden_dir):
    """ops from the device binarizer == oracle ops; fused encode == bytes from the REFERENCE engine."""
    z = np.load(os.path.join(golden_dir, "symbols_refengine.npz"))
    names = sorted({k[:-4] for k in z.files if k.endswith("_sym")})
    for nm in names:
        prof, meth, Nq, Nlbp, types, rows = [int(x) for x in z[nm + "_cfg"]]
        cfg = I.make_cfg(prof, meth, Nq, Nlbp, types, rows)
        sym = z[nm + "_sym"].astype(np.uint32)
        off = np.array([0, len(sym)], dtype=np.int64)
        ops, op_off = I.binarize_symbols(cfg, sym, off)
        assert (ops.cpu().numpy() == z[nm + "_ops"]).all(), nm
        assert int(op_off[-1].item()) == len(z[nm + "_ops"])
        want = z[nm + "_bytes"]
        enc = I.encode_symbols(cfg, sym, off, z[nm + "_ctx"], slab_stride=len(want) + 64)
        assert int(enc.lengths[0].item()) == len(want), nm
        assert (enc.slab[0, :len(want)].cpu().numpy() == want).all(), nm
        pay = I.compact(enc)
        dec, ok = I.decode_symbols(cfg, pay, off, z[nm + "_ctx"])
        assert bool(ok.all().item()) and (dec.cpu().numpy().astype(np.uint32) == sym).all(), nm
        # two-kernel route (binarize -> ops encoder) gives the same bytes
        enc2 = I.encode_ops(ops, op_off, z[nm + "_ctx"], slab_stride=len(want) + 64)
        assert (enc2.slab[0, :len(want)].cpu().numpy() == want).all(), nm


@pytest.mark.parametrize("prof,meth,Nq,rows,dtype", [
    (O.PROFILE_DEMO, O.BIN_TU, 4, 0, np.uint8),
    (O.PROFILE_DEMO, O.BIN_EG0, 4, 0, np.uint8),
    (O.PROFILE_ISS, O.BIN_EG0, 8, 400, np.uint8),       # W matrices of ISS.m: 400 x 20
    (O.PROFILE_ISS, O.BIN_EG0, 8, 109, np.uint16),      # H matrices: 109 x 20
    (O.PROFILE_ISS, O.BIN_EG2, 64, 50, np.uint32),
    (O.PROFILE_ISS, O.BIN_TU, 8, 20, np.uint8),
    (O.PROFILE_FLAT, O.BIN_EG0, 16, 0, np.uint8),       # config C4 segments
    (O.PROFILE_FLAT_EPSUF, O.BIN_EG2, 256, 0, np.uint8),  # config C5 streams
    (O.PROFILE_FLAT, O.BIN_FL32, 0, 0, np.uint32),
    (O.PROFILE_FLAT, O.BIN_TR0, 16, 0, np.uint8),       # truncated Rice: binarize / encode only, as upstream
    (O.PROFILE_ISS, O.BIN_TR1, 16, 40, np.uint8),
    (O.PROFILE_DEMO, O.BIN_TR2, 32, 0, np.uint8),
])
def test_many_streams_vs_oracle(I, prof, meth, Nq, rows, dtype):
    rng = np.random.default_rng(21)
    n_streams = 700
    if rows:
        counts = np.full(n_streams, rows * 20)
        counts[:50] = rng.integers(0, rows * 3, size=50)     # ragged: partial last column
    else:
        counts = np.clip(np.round(rng.lognormal(np.log(256), 1.0, size=n_streams)), 1, 4000).astype(np.int64)
    counts[0] = 0
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    if meth == O.BIN_FL32:
        sym = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    else:
        sym = np.minimum(np.floor(rng.exponential(Nq / 6.0 + 0.5, size=n)), Nq - 1).astype(dtype)
    nctx = O.num_ctx(prof, 3)
    ci = rng.integers(0, 126, size=(n_streams, nctx)).astype(np.uint8)
    cfg = I.make_cfg(prof, meth, Nq, 3, ALLT, rows)
    ocfg = O.make_cfg(prof, meth, Nq, 3, ALLT, rows)
    stride = 16 * 1024 if meth != O.BIN_FL32 else 32 * 1024
    enc, bits = I.encode_symbols(cfg, sym, off.astype(np.int64), ci, slab_stride=stride, want_bits=True)
    s_ref, l_ref, bits_ref = O.encode_symbols(ocfg, sym, off, ci, stride, n_threads=8, want_bits=True)
    torch.cuda.synchronize()
    enc.check_overflow()
    assert (enc.lengths.cpu().numpy().astype(np.uint32) == l_ref).all()
    w = int(l_ref.max())
    live = np.arange(w)[None, :] < l_ref[:, None]   # bytes past a stream's length are unspecified
    assert (enc.slab[:, :w].cpu().numpy()[live] == s_ref[:, :w][live]).all()
    assert (bits.cpu().numpy().astype(np.uint32) == bits_ref).all()        # getNumBits() after every symbol
    # the fused encoders proper (no per-symbol bit trace): the ring formulation (binarizer ahead of the coder through a
    # per-lane ring of op bytes) and the per-bin state machine, each against the oracle
    for ring in ("1", "0"):
        os.environ["ISSCABAC_SYM_RING"] = ring
        try:
            encf = I.encode_symbols(cfg, sym, off.astype(np.int64), ci, slab_stride=stride)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("ISSCABAC_SYM_RING")
        encf.check_overflow()
        assert (encf.lengths.cpu().numpy().astype(np.uint32) == l_ref).all(), ("fused encoder lengths", ring)
        assert (encf.slab[:, :w].cpu().numpy()[live] == s_ref[:, :w][live]).all(), ("fused encoder bytes", ring)
    pay = I.compact(enc)
    tdt = {np.uint8: torch.uint8, np.uint16: torch.int16, np.uint32: torch.int32}[dtype]
    if meth >= O.BIN_TR0:   # the reference's decode loops have no case for truncated Rice: refused, not guessed
        with pytest.raises(I.CabacError):
            I.decode_symbols(cfg, pay, off.astype(np.int64), ci, sym_dtype=tdt)
    else:
        # both fused decoders: the code-tree walk (u8 symbols of small alphabets) and the closed-form state machine
        for tree in ("1", "0"):
            os.environ["ISSCABAC_SYM_TREE"] = tree
            try:
                dec, ok = I.decode_symbols(cfg, pay, off.astype(np.int64), ci, sym_dtype=tdt)
                torch.cuda.synchronize()
            finally:
                os.environ.pop("ISSCABAC_SYM_TREE")
            assert bool(ok.all().item()), ("finish flags", tree)
            assert (dec.cpu().numpy().view(dtype) == sym).all(), ("decoded symbols", tree)
    # symbol-parallel binarizer against the oracle's op stream
    want = np.concatenate([O.symbols_to_ops(ocfg, sym[int(off[s]):int(off[s + 1])]) for s in range(n_streams)])
    for bin8 in ("1", "0"):       # u8 symbols: table kernels (k_bin_count8 / k_bin_emit8) and the closed-form ones
        os.environ["ISSCABAC_BIN8"] = bin8
        try:
            ops, op_off = I.binarize_symbols(cfg, sym, off.astype(np.int64))
        finally:
            os.environ.pop("ISSCABAC_BIN8")
        assert (ops.cpu().numpy() == want).all(), bin8


def test_symbol_host_api(I):
    rng = np.random.default_rng(22)
    sym = rng.integers(0, 8, size=400 * 20).astype(np.uint8)
    off = np.array([0, len(sym)], dtype=np.uint64)
    ci = O.ctx_from_p0(np.full(23, 128 / 255.0))
    cfg = I.make_cfg(I.PROFILE_ISS, "DEC2EG0", 8, 3, ["cond0", "cond1", "conds0", "conds1"], rows=400)
    payload, boff, bits = I.encode_symbols_host(cfg, sym, off, ci, want_bits=True)
    s_ref, l_ref, b_ref = O.encode_symbols(O.make_cfg(O.PROFILE_ISS, O.BIN_EG0, 8, 3, ALLT, 400), sym, off, ci, 8192, want_bits=True)
    assert bytes(payload) == bytes(s_ref[0, :l_ref[0]]) and (bits == b_ref).all()
    dec, ok = I.decode_symbols_host(cfg, payload, boff, off, ci, dtype=np.uint8)
    assert ok.all() and (dec == sym).all()


@pytest.mark.parametrize("name,scale", [("c2", 0.25), ("c4", 1 / 16), ("c5", 1 / 16)])
def test_baseline_symbol_configs_properties(I, name, scale):
    """BASELINE.json configs 2, 4 and 5 at a reduced stream count (tools/bench_symbols.py runs them at full
    size): size-independent properties -- fused and two-pass (binarize -> op encoder) encoders agree on
    every stream length, decode(encode(x)) == x, every finish() check holds, compaction is dense."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    T = I.CM_COND0 | I.CM_COND1 | I.CM_CONDS0 | I.CM_CONDS1
    if name == "c2":
        n, rows = int(65520 * scale), 400
        u = torch.rand(n * rows, generator=g, device=dev)
        sym = torch.where(u < 0.7, torch.zeros_like(u), 1 + torch.floor(torch.log(torch.rand(u.shape, generator=g, device=dev)) / np.log(0.6))).clamp_(0, 7).to(torch.uint8)
        off = torch.arange(n + 1, dtype=torch.int64, device=dev) * rows
        cfg, nctx = I.make_cfg(I.PROFILE_ISS, I.BIN_EG0, 8, 3, T, rows=rows), 23
    elif name == "c4":
        n, per = int((1 << 20) * scale), 1024
        sym = torch.floor(torch.log(torch.rand(n * per, generator=g, device=dev)) / np.log(0.5)).clamp_(0, 15).to(torch.uint8)
        off = torch.arange(n + 1, dtype=torch.int64, device=dev) * per
        cfg, nctx = I.make_cfg(I.PROFILE_FLAT, I.BIN_EG0, 16, 3, 0), 8
    else:
        rng = np.random.default_rng(4)
        n = int((1 << 20) * scale)
        lens = np.clip(np.round(rng.lognormal(np.log(256), 1.0, size=n)), 1, 65536).astype(np.int64)
        lens[:3] = [0, 1, 65536]                                     # empty, minimal and maximal streams
        offn = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=offn[1:])
        sym = torch.floor(-6.0 * torch.log(torch.rand(int(offn[-1]), generator=g, device=dev))).clamp_(0, 255).to(torch.uint8)
        off = torch.as_tensor(offn, device=dev)
        cfg, nctx = I.make_cfg(I.PROFILE_FLAT_EPSUF, I.BIN_EG2, 256, 3, 0), 4
    ctx = torch.full((nctx,), 1, dtype=torch.uint8, device=dev)
    ops, op_off = I.binarize_symbols(cfg, sym, off)
    stride = (int((op_off[1:] - op_off[:-1]).max().item()) // 4 + 64 + 15) & ~15
    e1 = I.encode_ops(ops, op_off, ctx, slab_stride=stride)
    e2 = I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride)
    e1.check_overflow(); e2.check_overflow()
    assert bool((e1.lengths == e2.lengths).all().item())
    p1, p2 = I.compact(e1), I.compact(e2)
    assert bool((p1.byte_off == p2.byte_off).all().item()) and bool((p1.payload == p2.payload).all().item())
    assert int(p1.byte_off[-1].item()) == int(e1.lengths.to(torch.int64).sum().item())     # dense, no gaps
    dec, ok = I.decode_symbols(cfg, p2, off, ctx, sym_dtype=torch.uint8)
    assert bool(ok.all().item()) and bool((dec == sym).all().item())
    bins, ok2 = I.decode_ops(p1, ops, op_off, ctx)
    assert bool(ok2.all().item()) and bool((bins == (ops & 1)).all().item())
    # the CUDA path against the ORACLE on a stratified sample of the streams (runs of 8 spread over the whole range, first
    # and last streams included): bytes and lengths of the fused encoder, decoded symbols
    ocfg = O.make_cfg(cfg.profile, cfg.method, cfg.Nq, cfg.Nlbp, cfg.types, cfg.rows)
    offn_all = off.cpu().numpy()
    starts = np.unique(np.round(np.linspace(0, n - 8, 48)).astype(np.int64))
    ids = np.unique(np.concatenate([np.arange(st, st + 8) for st in starts] + [np.arange(0, 3)]))
    lens_s = offn_all[ids + 1] - offn_all[ids]
    off_s = np.zeros(len(ids) + 1, dtype=np.uint64)
    np.cumsum(lens_s, out=off_s[1:])
    gather = np.concatenate([np.arange(offn_all[i], offn_all[i + 1]) for i in ids])
    sym_s = sym[torch.as_tensor(gather, device=dev)].cpu().numpy().astype(np.uint32)
    s_ref, l_ref = O.encode_symbols(ocfg, sym_s, off_s, np.full(nctx, 1, np.uint8), stride, n_threads=8)
    idt = torch.as_tensor(ids, device=dev)
    assert (e2.lengths[idt].cpu().numpy().astype(np.uint32) == l_ref).all(), "stream lengths differ from the oracle"
    w = int(l_ref.max())
    live = np.arange(w)[None, :] < l_ref[:, None]
    assert (e2.slab[idt][:, :w].cpu().numpy()[live] == s_ref[:, :w][live]).all(), "fused encoder bytes differ from the oracle"
    assert (e1.slab[idt][:, :w].cpu().numpy()[live] == s_ref[:, :w][live]).all(), "two-pass encoder bytes differ from the oracle"
    pay_ref, boff_ref = O.compact(s_ref, l_ref)
    d_ref, ok_ref = O.decode_symbols(ocfg, pay_ref, boff_ref, off_s, np.full(nctx, 1, np.uint8), n_threads=8)
    assert ok_ref.all() and (d_ref == sym_s).all() and (dec[torch.as_tensor(gather, device=dev)].cpu().numpy() == d_ref).all()
    if name == "c5":   # the empty stream is the two bytes of start(); finish() (KAT K0)
        b = p1.payload[int(p1.byte_off[0].item()):int(p1.byte_off[1].item())].cpu().numpy()
        assert bytes(b) == bytes.fromhex("fe80")


@pytest.mark.parametrize("shift,short", [(0, 0), (1, 0), (7, 5), (13, 40000), (15, 1)])
def test_binarizer_unaligned_and_short_buffers(I, shift, short):
    """The op-parallel binarizer writes aligned 16-byte pieces: an op buffer at any byte alignment and one
    that is too small (filled up to its capacity, nothing written behind it) give the oracle's ops."""
    import ctypes as C
    from isscabac_b200 import engine as E
    rng = np.random.default_rng(23 + shift)
    n_streams = 300
    counts = rng.integers(0, 500, size=n_streams)
    off = np.zeros(n_streams + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    sym = np.minimum(np.floor(rng.exponential(3.0, size=n)), 15).astype(np.uint8)
    cfg = I.make_cfg(O.PROFILE_FLAT, O.BIN_EG0, 16, 3, 0, 0)
    ocfg = O.make_cfg(O.PROFILE_FLAT, O.BIN_EG0, 16, 3, 0, 0)
    want = np.concatenate([O.symbols_to_ops(ocfg, sym[int(off[s]):int(off[s + 1])]) for s in range(n_streams)])
    total = len(want)
    cap = total - short
    dev = torch.device("cuda")
    L = I.lib()
    sym_t, off_t = torch.as_tensor(sym, device=dev), torch.as_tensor(off, device=dev)
    scratch = torch.empty(int(L.cabac_binarize_scratch_bytes(C.c_uint64(n), C.c_uint32(n_streams))), dtype=torch.uint8, device=dev)
    op_off = torch.empty(n_streams + 1, dtype=torch.int64, device=dev)
    buf = torch.full((total + 64,), 0xAB, dtype=torch.uint8, device=dev)
    ops = buf[shift:]
    E.check(L.cabac_binarize_symbols(C.byref(cfg), C.c_uint32(n_streams), E.vp(off_t), E.vp(sym_t), 1, C.c_uint64(n),
                                     E.vp(op_off), E.vp(ops), C.c_uint64(cap), E.vp(scratch), E._stream_ptr()))
    torch.cuda.synchronize()
    got = buf.cpu().numpy()
    assert int(op_off[-1].item()) == total
    assert (got[shift:shift + cap] == want[:cap]).all()
    assert (got[:shift] == 0xAB).all() and (got[shift + cap:] == 0xAB).all()


@pytest.mark.parametrize("prof,meth,Nq,rows", [(O.PROFILE_FLAT_EPSUF, O.BIN_EG2, 256, 0), (O.PROFILE_FLAT, O.BIN_EG0, 16, 0),
                                               (O.PROFILE_ISS, O.BIN_EG0, 8, 50), (O.PROFILE_DEMO, O.BIN_TU, 4, 0)])
def test_corrupt_symbol_streams_are_decoded_without_harm(I, prof, meth, Nq, rows):
    """A payload of random bytes (and a good payload with bytes flipped) goes through both fused decoders: the call returns,
    every decoded symbol is a value of the alphabet, nothing outside the output is written, and the streams whose bytes were
    replaced fail their finish() check (CABAC_ArithmeticDecoder.cpp:73-85) almost always.  The tree decoder's lanes keep
    decoding past the end of their stream inside a step group: this is the test that what they touch stays inside the tables."""
    rng = np.random.default_rng(77)
    n_streams = 257
    counts = rng.integers(0, 700, size=n_streams)
    counts[:2] = [0, 1]
    off = np.zeros(n_streams + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    sym = np.minimum(np.floor(rng.exponential(Nq / 6.0 + 0.5, size=n)), Nq - 1).astype(np.uint8)
    nctx = O.num_ctx(prof, 3)
    ci = rng.integers(0, 126, size=nctx).astype(np.uint8)
    cfg = I.make_cfg(prof, meth, Nq, 3, ALLT, rows)
    enc = I.encode_symbols(cfg, sym, off, ci, slab_stride=4096)
    pay = I.compact(enc)
    enc.check_overflow()
    good = pay.payload.clone()
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    for what in ("random", "flipped"):
        if what == "random":
            pay.payload.copy_(torch.randint(0, 256, good.shape, generator=g, device="cuda", dtype=torch.uint8))
        else:
            pay.payload.copy_(good)
            idx = torch.randint(0, good.numel(), (good.numel() // 50 + 1,), generator=g, device="cuda")
            pay.payload[idx] ^= 0x5A
        for tree in ("1", "0"):
            os.environ["ISSCABAC_SYM_TREE"] = tree
            try:
                guard = torch.full((n + 64,), 0xEE, dtype=torch.uint8, device="cuda")
                dec, ok = I.decode_symbols(cfg, pay, off, ci, sym_dtype=torch.uint8)
                torch.cuda.synchronize()
            finally:
                os.environ.pop("ISSCABAC_SYM_TREE")
            d = dec.cpu().numpy()
            assert d.shape[0] == n
            if tree == "1":      # the tree has no leaf outside the alphabet (an escape restarts at the root with value 0);
                assert (d < Nq).all(), what      # the closed-form decoder returns whatever the bins spell, like the reference
            assert int(ok.sum().item()) < n_streams, (what, tree)       # not every stream can pass with its bytes replaced
            del guard
    pay.payload.copy_(good)
    dec, ok = I.decode_symbols(cfg, pay, off, ci, sym_dtype=torch.uint8)
    assert bool(ok.all().item()) and (dec.cpu().numpy() == sym).all()


@pytest.mark.parametrize("prof,meth,Nq", [(O.PROFILE_FLAT_EPSUF, O.BIN_EG2, 256), (O.PROFILE_FLAT, O.BIN_EG0, 16)])
def test_more_streams_than_lanes(I, prof, meth, Nq):
    """More streams than the device has lanes: the persistent kernels hand streams out longest first through a work counter, and
    the fused decoder runs as TWO launches (one CTA with an SM of its own for the 128 longest streams beside the main launch,
    launch_sym_tree).  Ragged streams, a few very long ones among many short ones: the two-launch form, the one-launch form
    (ISSCABAC_TREE_SOLO=0) and the closed-form decoder agree with the input; a stratified sample of the encoder's bytes
    against the oracle."""
    rng = np.random.default_rng(83)
    n_streams = 200_000
    counts = np.clip(np.round(rng.lognormal(np.log(12), 0.8, size=n_streams)), 0, 400).astype(np.int64)
    counts[rng.integers(0, n_streams, size=40)] = rng.integers(3000, 9000, size=40)      # the long tail
    counts[:3] = [0, 1, 9000]
    off = np.zeros(n_streams + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    sym = np.minimum(np.floor(rng.exponential(Nq / 20.0 + 1.5, size=n)), Nq - 1).astype(np.uint8)
    nctx = O.num_ctx(prof, 3)
    ci = rng.integers(0, 126, size=nctx).astype(np.uint8)
    cfg, ocfg = I.make_cfg(prof, meth, Nq, 3, 0, 0), O.make_cfg(prof, meth, Nq, 3, 0, 0)
    stride = 9000 * 2 + 64
    enc = I.encode_symbols(cfg, sym, off, ci, slab_stride=stride)
    pay = I.compact(enc)
    enc.check_overflow()
    ids = np.unique(np.concatenate([np.arange(0, 8), np.round(np.linspace(8, n_streams - 9, 40)).astype(np.int64), np.arange(n_streams - 8, n_streams),
                                    np.argsort(counts)[-6:]]))
    off_s = np.zeros(len(ids) + 1, dtype=np.uint64)
    np.cumsum(counts[ids], out=off_s[1:])
    sym_s = np.concatenate([sym[off[i]:off[i + 1]] for i in ids]).astype(np.uint32)
    s_ref, l_ref = O.encode_symbols(ocfg, sym_s, off_s, ci, stride, n_threads=8)
    idt = torch.as_tensor(ids, device="cuda")
    assert (enc.lengths[idt].cpu().numpy().astype(np.uint32) == l_ref).all()
    w = int(l_ref.max())
    live = np.arange(w)[None, :] < l_ref[:, None]
    assert (enc.slab[idt][:, :w].cpu().numpy()[live] == s_ref[:, :w][live]).all()
    for env_name, val in (("ISSCABAC_TREE_SOLO", "1"), ("ISSCABAC_TREE_SOLO", "0"), ("ISSCABAC_SYM_TREE", "0")):
        os.environ[env_name] = val
        try:
            for _ in range(2):      # twice: the work counters and the side stream are reused
                dec, ok = I.decode_symbols(cfg, pay, off, ci, sym_dtype=torch.uint8)
                torch.cuda.synchronize()
        finally:
            os.environ.pop(env_name)
        assert bool(ok.all().item()), (env_name, val)
        assert (dec.cpu().numpy() == sym).all(), (env_name, val)


@pytest.mark.parametrize("prof,meth,Nq,rows,shape", [
    (O.PROFILE_FLAT, O.BIN_EG0, 16, 0, "ragged"),
    (O.PROFILE_FLAT, O.BIN_EG0, 256, 0, "ragged"),          # strings of up to 17 ops: long table entries and the closed form
    (O.PROFILE_FLAT, O.BIN_EG0, 32, 0, "uniform"),          # half the strings have 9 ops (two appends)
    (O.PROFILE_FLAT_EPSUF, O.BIN_EG2, 256, 0, "ragged"),
    (O.PROFILE_ISS, O.BIN_EG0, 8, 400, "columns"),
    (O.PROFILE_ISS, O.BIN_EG0, 64, 0, "ragged"),            # values outside the neighbour table's domain
    (O.PROFILE_DEMO, O.BIN_TU, 12, 0, "ragged"),
    (O.PROFILE_FLAT, O.BIN_FL32, 256, 0, "ragged"),         # 32 ops per symbol: a tile's ops leave the stage in rounds
    (O.PROFILE_FLAT, O.BIN_TU, 300, 0, "ragged"),           # strings of up to 256 ops
    (O.PROFILE_FLAT, O.BIN_TR0 + 2, 16, 0, "ragged"),
    (O.PROFILE_FLAT, O.BIN_EG0, 16, 0, "tiny"),             # ~1,400 streams start in every tile
    (O.PROFILE_ISS, O.BIN_EG0, 16, 3, "tiny"),              # ... more than the shared-memory slice of stream offsets holds
    (O.PROFILE_DEMO, O.BIN_EG0, 16, 0, "tiny"),
])
def test_binarizer_u8_table_kernels(I, prof, meth, Nq, rows, shape):
    """k_bin_count8 / k_bin_emit8 (u8 symbols: bin counts by table, op strings appended word-wise) against the oracle's op
    stream and offsets: empty streams, streams that start on tile and thread boundaries, a symbol buffer at odd byte
    alignments (no vector loads), an input that ends inside a thread's run of 8, thousands of tiny and empty streams."""
    import ctypes as C
    from isscabac_b200 import engine as E
    rng = np.random.default_rng(prof * 1000 + meth * 31 + Nq)
    if shape == "columns":
        counts = np.full(160, rows, dtype=np.int64)
    elif shape == "tiny":
        counts = rng.integers(0, 4, size=9000)
        counts[4000:4600] = 0                                # hundreds of empty streams in a row, and at the very end
        counts[-300:] = 0
    else:
        counts = rng.integers(0, 900, size=220)
        counts[rng.integers(0, len(counts), size=30)] = 0
        counts[[0, 5, 6]] = [2048, 8, 2040]             # boundaries of tiles (2,048 symbols) and of 8-symbol runs
        counts[-2:] = 0
    n_streams = len(counts)
    off = np.zeros(n_streams + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    hi = min(Nq, 256)
    sym = (rng.integers(0, hi, size=n) if shape == "uniform" else np.minimum(np.floor(rng.exponential(hi / 5.0, size=n)), hi - 1)).astype(np.uint8)
    cfg = I.make_cfg(prof, meth, Nq, 3, 0x1b, rows)
    ocfg = O.make_cfg(prof, meth, Nq, 3, 0x1b, rows)
    per = [O.symbols_to_ops(ocfg, sym[int(off[s]):int(off[s + 1])]) for s in range(n_streams)]
    want = np.concatenate(per)
    want_off = np.zeros(n_streams + 1, dtype=np.int64)
    np.cumsum([len(x) for x in per], out=want_off[1:])
    dev = torch.device("cuda")
    L = I.lib()
    off_t = torch.as_tensor(off, device=dev)
    scratch = torch.empty(int(L.cabac_binarize_scratch_bytes(C.c_uint64(n), C.c_uint32(n_streams))), dtype=torch.uint8, device=dev)
    for sym_shift, op_shift in ((0, 0), (1, 3), (8, 15), (13, 8)):
        symbuf = torch.zeros(n + 32, dtype=torch.uint8, device=dev)
        sym_t = symbuf[sym_shift:sym_shift + n]
        sym_t.copy_(torch.as_tensor(sym, device=dev))
        op_off = torch.full((n_streams + 1,), -1, dtype=torch.int64, device=dev)
        E.check(L.cabac_binarize_symbols(C.byref(cfg), C.c_uint32(n_streams), E.vp(off_t), E.vp(sym_t), 1, C.c_uint64(n),
                                         E.vp(op_off), None, C.c_uint64(0), E.vp(scratch), E._stream_ptr()))
        assert (op_off.cpu().numpy() == want_off).all(), ("offsets-only call", sym_shift)
        buf = torch.full((len(want) + 64,), 0xAB, dtype=torch.uint8, device=dev)
        op_off.fill_(-1)
        E.check(L.cabac_binarize_symbols(C.byref(cfg), C.c_uint32(n_streams), E.vp(off_t), E.vp(sym_t), 1, C.c_uint64(n),
                                         E.vp(op_off), E.vp(buf[op_shift:]), C.c_uint64(len(want)), E.vp(scratch), E._stream_ptr()))
        torch.cuda.synchronize()
        got = buf.cpu().numpy()
        assert (op_off.cpu().numpy() == want_off).all(), ("offsets", sym_shift)
        bad = np.nonzero(got[op_shift:op_shift + len(want)] != want)[0]
        assert len(bad) == 0, ("ops", sym_shift, op_shift, int(bad[0]), len(bad))
        assert (got[:op_shift] == 0xAB).all() and (got[op_shift + len(want):] == 0xAB).all()


def test_binarizer_single_call_into_callers_buffer(I):
    """engine.binarize_symbols(..., ops=buffer): one stream-ordered call, ops and offsets equal the two-call form's; a
    buffer that is too small is filled up to its size and op_off[-1] still tells the true total."""
    rng = np.random.default_rng(77)
    counts = rng.integers(0, 3000, size=90)
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    sym = np.minimum(np.floor(rng.exponential(2.0, size=int(off[-1]))), 15).astype(np.uint8)
    cfg = I.make_cfg(O.PROFILE_FLAT, O.BIN_EG0, 16, 3, 0, 0)
    ops2, off2 = I.binarize_symbols(cfg, sym, off)
    total = int(off2[-1].item())
    dev = torch.device("cuda")
    buf = torch.full((total + 100,), 0xCD, dtype=torch.uint8, device=dev)
    ops1, off1 = I.binarize_symbols(cfg, sym, off, ops=buf)
    assert bool((off1 == off2).all().item()) and bool((ops1[:total] == ops2).all().item()) and bool((buf[total:] == 0xCD).all().item())
    small = torch.full((total - 1234,), 0xCD, dtype=torch.uint8, device=dev)
    guard = torch.full((4096,), 0xCD, dtype=torch.uint8, device=dev)
    ops3, off3 = I.binarize_symbols(cfg, sym, off, ops=small)
    assert int(off3[-1].item()) == total and bool((ops3 == ops2[:total - 1234]).all().item()) and bool((guard == 0xCD).all().item())


@pytest.mark.parametrize("meth", [O.BIN_TU, O.BIN_EG0])
def test_c1_demo_sequence_full_size(I, meth):
    """BASELINE configs[0] at its full size: the cabacDemo sequence (cabacDemo.m:26-37: 10,000 correlated half-normal samples,
    Nq = 4), demo context rule (3 contexts, p0 = 0.5), TU as the reference runs it and EG0 as BASELINE.json words it -- one
    stream through the fused kernels and through binarizer + op encoder, bytes equal to the oracle's, symbols decode back."""
    rng = np.random.default_rng(0)
    N, Nq = 10000, 4
    x = np.abs(rng.standard_normal(N))
    x[1:] += 0.8 * x[:-1]
    delta = np.quantile(x, 0.99) / Nq
    sym = np.minimum(np.floor(x / delta + 0.5), Nq - 1).astype(np.uint8)
    off = np.array([0, N], dtype=np.int64)
    ctx = O.ctx_from_p0(O.matlab_uint8(0.5 * np.ones(3) * 255) / 255.0)
    cfg, ocfg = I.make_cfg(O.PROFILE_DEMO, meth, Nq), O.make_cfg(O.PROFILE_DEMO, meth, Nq)
    s_ref, l_ref = O.encode_symbols(ocfg, sym.astype(np.uint32), off.astype(np.uint64), ctx, 8192)
    want = bytes(s_ref[0, :l_ref[0]])
    enc = I.encode_symbols(cfg, sym, off, ctx, slab_stride=8192)
    enc.check_overflow()
    assert int(enc.lengths[0].item()) == len(want) and bytes(enc.slab[0, :len(want)].cpu().numpy()) == want
    ops, op_off = I.binarize_symbols(cfg, sym, off)
    assert (ops.cpu().numpy() == O.symbols_to_ops(ocfg, sym.astype(np.uint32))).all()
    enc2 = I.encode_ops(ops, op_off, ctx, slab_stride=8192)
    assert int(enc2.lengths[0].item()) == len(want) and bytes(enc2.slab[0, :len(want)].cpu().numpy()) == want
    dec, ok = I.decode_symbols(cfg, I.compact(enc), off, ctx, sym_dtype=torch.uint8)
    assert bool(ok.all().item()) and (dec.cpu().numpy() == sym).all()
