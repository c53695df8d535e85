"""CPU: the product's per-lane coder logic (isscabac_b200/csrc/cabac_lane.cuh -- the code the
CUDA kernels execute) compiled for the host and checked against the golden vectors and the
oracle.  This validates the algorithmic restructuring (eager output + walk-back carries,
fused MPS/LPS step, closed-form binarizer, incremental symbol decoder) without a GPU; the
-m gpu tests check the kernels themselves."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emul", "lane_emul.cpp")
    so = os.path.join(HERE, "emul", "liblane_emul.so")
    hdrs = [os.path.join(HERE, "..", "isscabac_b200", "csrc", h) for h in ("cabac_lane.cuh", "cabac_wide.cuh", "cabac_spec.cuh", "bin_emit.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.emul_symbols_to_ops.restype = C.c_uint64
    L.emul_bin_emit8.restype = C.c_uint64
    return L


def p(a, t):
    return a.ctypes.data_as(t)


def emul_encode(L, ops, off, ci, stride, trace=False):
    ops = np.ascontiguousarray(ops)
    w = ops.dtype.itemsize
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = len(off) - 1
    ci = np.ascontiguousarray(ci, dtype=np.uint8)
    per = int(ci.ndim == 2)
    n_ctx = ci.shape[-1]
    slab = np.zeros((n, stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint32)
    tr = np.zeros(max(len(ops), 1), dtype=np.uint32) if trace else None
    ovf = L.emul_encode_ops(C.c_uint32(n), p(off, u64p), ops.ctypes.data_as(C.c_void_p), w,
                            p(ci.reshape(-1), u8p), C.c_uint32(n_ctx), per, p(slab, u8p), C.c_uint64(stride),
                            p(lens, u32p), p(tr, u32p) if trace else None)
    assert ovf == 0
    return (slab, lens, tr) if trace else (slab, lens)


def emul_decode(L, payload, boff, ops, off, ci):
    ops = np.ascontiguousarray(ops)
    w = ops.dtype.itemsize
    ci = np.ascontiguousarray(ci, dtype=np.uint8)
    per = int(ci.ndim == 2)
    n = len(off) - 1
    bins = np.zeros(max(len(ops), 1), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint8)
    payload = np.ascontiguousarray(payload) if len(payload) else np.zeros(1, np.uint8)
    L.emul_decode_ops(C.c_uint32(n), p(np.ascontiguousarray(boff, dtype=np.uint64), u64p), p(payload, u8p),
                      p(np.ascontiguousarray(off, dtype=np.uint64), u64p), ops.ctypes.data_as(C.c_void_p), w,
                      p(ci.reshape(-1), u8p), C.c_uint32(ci.shape[-1]), per, p(bins, u8p), p(ok, u8p))
    return bins[:len(ops)], ok


def script_to_ops(script):
    ops = []
    for k, a, b in script:
        if k == 0:
            ops.append((b << 1) | a)
        elif k == 1:
            ops.append((O.OP8_EP << 1) | a)
        elif k == 2:
            ops += [(O.OP8_EP << 1) | ((a >> (b - 1 - i)) & 1) for i in range(b)]   # a7: identical to single EP bins
        else:
            ops.append((O.OP8_TRM << 1) | a)
    return np.array(ops, dtype=np.uint8)


def test_lane_kats(emul, golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        kat = json.load(f)
    for name, k in kat.items():
        ops = script_to_ops(k["script"])
        off = np.array([0, len(ops)], dtype=np.uint64)
        slab, lens = emul_encode(emul, ops, off, np.array(k["ctx"], dtype=np.uint8), 64)
        assert bytes(slab[0, :lens[0]]).hex() == k["bytes"], name
        if k["decoded"] is not None:
            data = np.frombuffer(bytes.fromhex(k["bytes"]), dtype=np.uint8)
            bins, ok = emul_decode(emul, data, [0, len(data)], ops, off, np.array(k["ctx"], dtype=np.uint8))
            assert ok[0] == 1 and (bins == (ops & 1)).all(), name


@pytest.mark.parametrize("fname", ["random_ops.npz", "random_ops16.npz"])
def test_lane_random_golden(emul, golden_dir, fname):
    z = np.load(os.path.join(golden_dir, fname))
    ops, off = z["ops"], z["op_off"]
    for tag, ci in (("shared", z["ctx_shared"]), ("per", z["ctx_per"])):
        slab, lens = emul_encode(emul, ops, off, ci, 512)
        assert (lens == z["lens_" + tag]).all()
        payload, boff = O.compact(slab, lens)
        assert (payload == z["payload_" + tag]).all()
        bins, ok = emul_decode(emul, payload, boff, ops, off, ci)
        assert ok.all() and (bins == (ops & 1)).all()


def test_lane_vs_oracle_bulk_and_numbits(emul):
    # ~3M bins in short streams: hits writeOut carries, ripple carries and finish() carries
    rng = np.random.default_rng(9)
    n_streams, n_ops = 40000, 72
    n = n_streams * n_ops
    code = rng.integers(0, 4, size=n).astype(np.uint8)
    bins = (rng.random(n) < 0.15).astype(np.uint8)
    ep = rng.random(n) < 0.25
    code[ep] = O.OP8_EP
    ops = (code << 1) | bins
    off = (np.arange(n_streams + 1) * n_ops).astype(np.uint64)
    ci = np.full(4, 1, dtype=np.uint8)
    s1, l1, tr = emul_encode(emul, ops, off, ci, 64, trace=True)
    s0, l0 = O.encode_ops(ops, off, ci, out_stride=64, n_threads=8)
    assert (l0 == l1).all() and (s0 == s1).all()


def test_lane_numbits_trace(emul):
    # getNumBits() after every bin: lane bookkeeping vs the oracle's bit sink (CABAC_BitstreamFile.h:70)
    rng = np.random.default_rng(10)
    for trial in range(30):
        n = 600
        code = rng.integers(0, 3, size=n).astype(np.uint8)
        bins = (rng.random(n) < (0.5 if trial % 2 else 0.03)).astype(np.uint8)   # p=0.5 at state 0: many 0xFF leads
        ops = (code << 1) | bins
        off = np.array([0, n], dtype=np.uint64)
        ci = np.array([1, 1, 1], dtype=np.uint8)
        _, _, tr = emul_encode(emul, ops, off, ci, 1024, trace=True)
        # oracle trace through the symbol-less path: encode bin by bin
        import ctypes as CC
        L = O.lib()
        e = O._Enc()
        out = np.zeros(1024, dtype=np.uint8)
        cs = ci.copy()
        L.orc_enc_attach(CC.byref(e), p(out, u8p), CC.c_uint64(1024))
        L.orc_enc_start(CC.byref(e))
        want = []
        for o in ops:
            L.orc_enc_bin(CC.byref(e), int(o & 1), CC.byref(CC.c_uint8.from_buffer(cs, int(o >> 1))))
            want.append(L.orc_enc_num_bits(CC.byref(e)))
        assert list(tr[:n]) == want


def test_lane_symbol_path(emul, golden_dir):
    z = np.load(os.path.join(golden_dir, "symbols_refengine.npz"))
    names = sorted({k[:-4] for k in z.files if k.endswith("_sym")})
    for nm in names:
        cfgv = np.ascontiguousarray(z[nm + "_cfg"], dtype=np.int32)
        sym = np.ascontiguousarray(z[nm + "_sym"], dtype=np.uint32)
        ops = np.zeros(len(sym) * 80, dtype=np.uint8)
        k = emul.emul_symbols_to_ops(cfgv.ctypes.data_as(C.POINTER(C.c_int32)), p(sym, u32p), C.c_uint64(len(sym)), p(ops, u8p))
        assert k == len(z[nm + "_ops"]) and (ops[:k] == z[nm + "_ops"]).all(), nm
        data = np.ascontiguousarray(z[nm + "_bytes"])
        out = np.zeros(len(sym), dtype=np.uint32)
        ok = np.zeros(1, dtype=np.uint8)
        ci = np.ascontiguousarray(z[nm + "_ctx"])
        emul.emul_decode_symbols(cfgv.ctypes.data_as(C.POINTER(C.c_int32)), p(data, u8p), C.c_uint32(len(data)),
                                 C.c_uint64(len(sym)), p(ci, u8p), C.c_uint32(len(ci)), p(out, u32p), p(ok, u8p))
        assert ok[0] == 1 and (out == sym).all(), nm


def test_lane_symbol_wide_values(emul):
    # closed-form EG-k / FL32 / TU codes against the oracle's bin-by-bin binarizer, incl. 32-bit extremes
    vals = np.array(list(range(0, 70)) + [255, 256, 1023, 65535, 65536, 2**31 - 1, 2**31, 2**32 - 2, 2**32 - 1], dtype=np.uint32)
    for meth, Nq in ((O.BIN_EG0, 256), (O.BIN_EG1, 256), (O.BIN_EG2, 256), (O.BIN_FL32, 256), (O.BIN_TU, 80)):
        v = vals if meth != O.BIN_TU else vals[vals < 80]
        cfgv = np.array([O.PROFILE_FLAT, meth, Nq, 3, 0, 0], dtype=np.int32)
        ops = np.zeros(len(v) * 80, dtype=np.uint8)
        k = emul.emul_symbols_to_ops(cfgv.ctypes.data_as(C.POINTER(C.c_int32)), p(np.ascontiguousarray(v), u32p), C.c_uint64(len(v)), p(ops, u8p))
        want = O.symbols_to_ops(O.make_cfg(O.PROFILE_FLAT, meth, Nq, 3, 0, 0), v)
        assert k == len(want) and (ops[:k] == want).all(), meth


@pytest.mark.parametrize("k", [0, 1, 2])
def test_lane_truncated_rice(emul, k):
    """DEC2TR0 / TR1 / TR2 (cabacBinarizer.m:39-54, encode-only upstream): the closed-form code of cabac_lane.cuh against
    the oracle's bin-by-bin TRCode -- values below, at and above maxVal (the unfinished escape: suffix bits stay ones) --
    and, for k = 0, against the arithmetic of TRCode worked by hand: n_p = v + 1, n_s = 0 -> v ones and a zero."""
    Nq = 16
    v = np.array(list(range(0, 40)) + [100, 255], dtype=np.uint32)
    for prof in (O.PROFILE_FLAT, O.PROFILE_DEMO, O.PROFILE_ISS):
        cfgv = np.array([prof, O.BIN_TR0 + k, Nq, 3, 0x1b, 0], dtype=np.int32)
        ops = np.zeros(int(v.sum()) + 8 * len(v) + 64, dtype=np.uint8)
        n = emul.emul_symbols_to_ops(cfgv.ctypes.data_as(C.POINTER(C.c_int32)), p(v, u32p), C.c_uint64(len(v)), p(ops, u8p))
        want = O.symbols_to_ops(O.make_cfg(prof, O.BIN_TR0 + k, Nq, 3, 0x1b, 0), v)
        assert n == len(want) and (ops[:n] == want).all(), (k, prof)
    if k == 0:
        for x in (0, 1, 5, 15, 16):
            assert list(O.binarize(x, Nq, O.BIN_TR0)) == [1] * x + [0]
    else:
        assert list(O.binarize(5, Nq, O.BIN_TR0 + k)) == [1] * (5 >> k) + [0] + [(5 >> (k - 1 - i)) & 1 for i in range(k)]
        assert list(O.binarize(15, Nq, O.BIN_TR0 + k)) == [1] * (15 >> k) + [0] + [1] * k        # v >= maxVal: escape TODO upstream


@pytest.mark.parametrize("prof,meth,Nq,rows", [
    (O.PROFILE_FLAT, O.BIN_EG0, 32, 0),          # C4's shape: strings of 1..9 ops (the long ones in pieces of 7)
    (O.PROFILE_FLAT, O.BIN_EG0, 256, 0),         # up to 17 ops: closed form, 7 ops per append
    (O.PROFILE_FLAT_EPSUF, O.BIN_EG2, 64, 0),    # C5's shape
    (O.PROFILE_ISS, O.BIN_EG0, 20, 7),           # neighbour-conditioned table, rows of 7
    (O.PROFILE_ISS, O.BIN_EG0, 64, 0),           # values outside the table's domain (31)
    (O.PROFILE_DEMO, O.BIN_TU, 12, 0),
    (O.PROFILE_DEMO, O.BIN_EG0, 16, 0),
    (O.PROFILE_FLAT, O.BIN_FL32, 256, 0),        # 32 ops per symbol: no table at all, tiles larger than the stage
    (O.PROFILE_FLAT, O.BIN_TR0 + 1, 16, 0),
    (O.PROFILE_FLAT, O.BIN_TU, 300, 0),          # strings of up to 256 ops
])
def test_bin_emit8_tile_logic(emul, prof, meth, Nq, rows):
    """The u8 binarizer's emit phase (isscabac_b200/csrc/bin_emit.cuh: op strings appended word-wise to a stage, tail
    bytes after the barrier) run for whole tiles on the host, against the oracle's op stream: every op-array alignment,
    a stage small enough to force the rounds of an over-full tile, three orders of the threads inside a phase, inputs
    that end inside a thread's run of 8 and inside a tile."""
    rng = np.random.default_rng(prof * 100 + meth * 7 + Nq)
    cfgv = np.array([prof, meth, Nq, 3, 0x1b, rows], dtype=np.int32)
    cfg = O.make_cfg(prof, meth, Nq, 3, 0x1b, rows)
    hi = min(Nq, 256)
    cases = [(1, 0, 16384, 0), (5, 3, 16384, 1), (8, 15, 16384, 2), (2048, 1, 16384, 0), (2049, 7, 16384, 1), (4999, 5, 16384, 2),
             (4999, 9, 256, 0), (6000, 2, 64, 1), (2048 * 3, 0, 1024, 2)]
    for n, skew, stage, order in cases:
        skewed = rng.random() < 0.5
        sym = (np.minimum(rng.geometric(0.45, n) - 1, hi - 1) if skewed else rng.integers(0, hi, n)).astype(np.uint8)
        want = O.symbols_to_ops(cfg, sym.astype(np.uint32))
        got = np.full(len(want) + 64, 0xAA, dtype=np.uint8)
        k = emul.emul_bin_emit8(cfgv.ctypes.data_as(C.POINTER(C.c_int32)), p(sym, u8p), C.c_uint64(n), p(got, u8p),
                                C.c_uint32(skew), C.c_uint32(stage), order)
        assert k == len(want), (n, skew, stage, order)
        assert (got[:k] == want).all(), (n, skew, stage, order, int(np.argmax(got[:k] != want)))
        assert (got[k:] == 0xAA).all()
