"""TEST INFRASTRUCTURE ONLY.

ctypes front-end to the parity oracle:

* ``liboracle_cabac.so``  -- the C restatement of the reference CABAC path
  (``cabac_oracle.c``; every function cites the reference file:line it follows);
* ``_ref/libref_cabac.so`` -- the UNMODIFIED reference engine behind our batch
  driver (``ref_driver.cpp``), built only where ``/root/reference`` exists and
  shipped prebuilt to the GPU box;
* ``_ref/libref_mex.so``  -- the UNMODIFIED reference ``mexFunction`` compiled
  against the stub ``mexstub/mex.h``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs may import this package.  Nothing under
``isscabac_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference/CABAC"

# binarization methods / profiles / cm-type mask (mirror cabac_oracle.h)
BIN_TU, BIN_EG0, BIN_EG1, BIN_EG2, BIN_FL32, BIN_TR0, BIN_TR1, BIN_TR2 = range(8)
PROFILE_DEMO, PROFILE_ISS, PROFILE_FLAT, PROFILE_FLAT_EPSUF = range(4)
CM_COND0, CM_COND1, CM_CONDBINLFT, CM_CONDS0, CM_CONDS1 = 1, 2, 4, 8, 16
OP8_TRM, OP8_EP = 125, 126
OP16_TRM, OP16_EP = 0x7FFD, 0x7FFE

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> None:
    """Compile the C restatement and, when the reference tree is present, oracle/_ref."""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle_cabac.so"))
    if os.path.isdir(REF_DIR):
        for f in ("libref_cabac.so", "libref_mex.so", "libref_mex_trace.so"):
            need = need or not os.path.exists(os.path.join(HERE, "_ref", f))
    if need or force:
        subprocess.run(["make", "-C", HERE, "all"] + (["-B"] if force else []),
                       check=True, capture_output=True)


class _Enc(C.Structure):
    _fields_ = [("low", C.c_uint32), ("range", C.c_uint32), ("bits_left", C.c_int32),
                ("buffered_byte", C.c_uint32), ("num_buffered", C.c_int32),
                ("bins_coded", C.c_uint64), ("out", _u8p), ("cap", C.c_uint64),
                ("n_out", C.c_uint64), ("held", C.c_uint32), ("n_held", C.c_uint32),
                ("bits_written", C.c_uint64)]


class _Dec(C.Structure):
    _fields_ = [("range", C.c_uint32), ("value", C.c_uint32), ("bits_needed", C.c_int32),
                ("inp", _u8p), ("len", C.c_uint64), ("pos", C.c_uint64), ("last_byte", C.c_uint32)]


class SymCfg(C.Structure):
    _fields_ = [("profile", C.c_int), ("method", C.c_int), ("Nq", C.c_uint32),
                ("Nlbp", C.c_int), ("types", C.c_uint), ("rows", C.c_uint32)]


_LIB = None
_REF = None
_MEX = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle_cabac.so"))
        L.orc_enc_num_bits.restype = C.c_uint64
        L.orc_symbols_to_ops.restype = C.c_uint64
        L.orc_ctx_from_p0.restype = C.c_uint8
        L.orc_ctx_from_p0.argtypes = [C.c_double]
        L.orc_matlab_uint8.restype = C.c_uint8
        L.orc_matlab_uint8.argtypes = [C.c_double]
        L.orc_ctx_next_mps.restype = C.c_uint8
        L.orc_ctx_next_lps.restype = C.c_uint8
        L.orc_dec_bins_ep.restype = C.c_uint32
        L.orc_debinarize.restype = C.c_uint32
        _LIB = L
    return _LIB


def ref():
    """The unmodified reference engine (or None when oracle/_ref was never built)."""
    global _REF
    if _REF is None:
        p = os.path.join(HERE, "_ref", "libref_cabac.so")
        if not os.path.exists(p) and os.path.isdir(REF_DIR):
            build()
        _REF = C.CDLL(p) if os.path.exists(p) else False
    return _REF or None


def ref_mex():
    global _MEX
    if _MEX is None:
        p = os.path.join(HERE, "_ref", "libref_mex.so")
        if not os.path.exists(p) and os.path.isdir(REF_DIR):
            build()
        _MEX = C.CDLL(p) if os.path.exists(p) else False
    return _MEX or None


_MEXT = None


def ref_mex_trace():
    """The unmodified reference mexFunction built with -D_WIN32, i.e. with the reference's
    RWTH_TRACE_CABAC_STATES code (getEncoderStats / getDecoderStats) compiled in."""
    global _MEXT
    if _MEXT is None:
        p = os.path.join(HERE, "_ref", "libref_mex_trace.so")
        if not os.path.exists(p) and os.path.isdir(REF_DIR):
            build()
        _MEXT = C.CDLL(p) if os.path.exists(p) else False
    return _MEXT or None


def tmpdir() -> str:
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()


def _p(a, t):
    return a.ctypes.data_as(t)


def _ops_arr(ops):
    ops = np.ascontiguousarray(ops)
    if ops.dtype == np.uint8:
        return ops, 1
    if ops.dtype == np.uint16:
        return ops, 2
    raise TypeError("ops must be uint8 or uint16")


def _prep_ctx(ctx_init, n_streams):
    ci = np.ascontiguousarray(ctx_init, dtype=np.uint8)
    if ci.ndim == 2:
        assert ci.shape[0] == n_streams
        return ci.reshape(-1), ci.shape[1], 1
    return ci, ci.shape[0], 0


def encode_ops(ops, op_off, ctx_init, out_stride=None, n_threads=1, impl="oracle"):
    """-> (slab[n_streams, out_stride] u8, lengths u32).  impl: 'oracle' | 'ref'."""
    ops, w = _ops_arr(ops)
    op_off = np.ascontiguousarray(op_off, dtype=np.uint64)
    n = len(op_off) - 1
    ci, n_ctx, per = _prep_ctx(ctx_init, n)
    if out_stride is None:
        longest = int((op_off[1:] - op_off[:-1]).max()) if n else 0
        out_stride = longest + 16  # >1 byte per op never happens in practice; checked below
    slab = np.zeros((n, out_stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint32)
    if impl == "ref":
        r = ref()
        assert r is not None, "oracle/_ref not built"
        rc = r.ref_encode_ops(C.c_uint32(n), _p(op_off, _u64p), ops.ctypes.data_as(C.c_void_p), w,
                              _p(ci, _u8p), C.c_uint32(n_ctx), per, _p(slab, _u8p),
                              C.c_uint64(out_stride), _p(lens, _u32p), n_threads, tmpdir().encode())
    else:
        rc = lib().orc_encode_ops(C.c_uint32(n), _p(op_off, _u64p), ops.ctypes.data_as(C.c_void_p), w,
                                  _p(ci, _u8p), C.c_uint32(n_ctx), per, _p(slab, _u8p),
                                  C.c_uint64(out_stride), _p(lens, _u32p), n_threads)
    assert rc == 0
    assert n == 0 or int(lens.max()) <= out_stride, "slab too small"
    return slab, lens


def compact(slab, lens):
    """-> (payload u8, byte_off u64[n+1]) : streams back to back."""
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    payload = np.zeros(int(off[-1]), dtype=np.uint8)
    for s, l in enumerate(lens):
        payload[int(off[s]):int(off[s + 1])] = slab[s, :l]
    return payload, off


def decode_ops(payload, byte_off, ops, op_off, ctx_init, n_threads=1, impl="oracle"):
    """-> (bins u8[n_ops], finish_ok u8[n_streams])."""
    ops, w = _ops_arr(ops)
    op_off = np.ascontiguousarray(op_off, dtype=np.uint64)
    byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    n = len(op_off) - 1
    ci, n_ctx, per = _prep_ctx(ctx_init, n)
    bins = np.zeros(len(ops), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint8)
    if payload.size == 0:
        payload = np.zeros(1, dtype=np.uint8)
    if impl == "ref":
        r = ref()
        assert r is not None, "oracle/_ref not built"
        rc = r.ref_decode_ops(C.c_uint32(n), _p(byte_off, _u64p), _p(payload, _u8p), _p(op_off, _u64p),
                              ops.ctypes.data_as(C.c_void_p), w, _p(ci, _u8p), C.c_uint32(n_ctx), per,
                              _p(bins, _u8p), _p(ok, _u8p), n_threads, tmpdir().encode())
    else:
        rc = lib().orc_decode_ops(C.c_uint32(n), _p(byte_off, _u64p), _p(payload, _u8p), _p(op_off, _u64p),
                                  ops.ctypes.data_as(C.c_void_p), w, _p(ci, _u8p), C.c_uint32(n_ctx), per,
                                  _p(bins, _u8p), _p(ok, _u8p), n_threads)
    assert rc == 0
    return bins, ok


# --------------------------------------------------------------------------
# single-stream "engine scripts": entries (kind, a, b)
#   0 encodeBin(bin=a, ctx=b)   1 encodeBinEP(a)   2 encodeBinsEP(a, n=b)   3 encodeBinTrm(a)
# --------------------------------------------------------------------------
def encode_script(script, ctx_init, impl="oracle", cap=1 << 16):
    """-> (bytes, final ctx states)"""
    ci = np.ascontiguousarray(ctx_init, dtype=np.uint8).copy()
    out = np.zeros(cap, dtype=np.uint8)
    if impl == "ref":
        r = ref()
        assert r is not None
        sc = np.ascontiguousarray(np.asarray(script, dtype=np.uint32).reshape(-1, 3))
        fin = np.zeros(max(len(ci), 1), dtype=np.uint8)
        n = r.ref_encode_script(_p(sc, _u32p), C.c_uint32(len(sc)), _p(ci, _u8p), C.c_uint32(len(ci)),
                                _p(out, _u8p), C.c_uint32(cap), _p(fin, _u8p), tmpdir().encode())
        assert n >= 0
        return bytes(out[:n]), fin[:len(ci)].copy()
    L = lib()
    e = _Enc()
    L.orc_enc_attach(C.byref(e), _p(out, _u8p), C.c_uint64(cap))
    L.orc_enc_start(C.byref(e))
    for k, a, b in script:
        if k == 0:
            L.orc_enc_bin(C.byref(e), int(a), C.byref(C.c_uint8.from_buffer(ci, int(b))))
        elif k == 1:
            L.orc_enc_ep(C.byref(e), int(a))
        elif k == 2:
            L.orc_enc_bins_ep(C.byref(e), C.c_uint32(int(a)), int(b))
        else:
            L.orc_enc_trm(C.byref(e), int(a))
    L.orc_enc_finish(C.byref(e))
    return bytes(out[:e.n_out]), ci


def decode_script(script, ctx_init, data: bytes, impl="oracle"):
    """script kinds: 0 decodeBin(ctx=b), 1 decodeBinEP, 2 decodeBinsEP(n=b), 3 decodeBinTrm -> list of values"""
    ci = np.ascontiguousarray(ctx_init, dtype=np.uint8).copy()
    buf = np.frombuffer(data, dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
    if impl == "ref":
        r = ref()
        assert r is not None
        sc = np.ascontiguousarray(np.asarray(script, dtype=np.uint32).reshape(-1, 3))
        res = np.zeros(len(sc), dtype=np.uint32)
        rc = r.ref_decode_script(_p(sc, _u32p), C.c_uint32(len(sc)), _p(ci, _u8p), C.c_uint32(len(ci)),
                                 _p(buf, _u8p), C.c_uint32(len(data)), _p(res, _u32p), tmpdir().encode())
        assert rc == 0
        return [int(x) for x in res]
    L = lib()
    d = _Dec()
    L.orc_dec_start(C.byref(d), _p(buf, _u8p), C.c_uint64(len(data)))
    res = []
    for k, a, b in script:
        if k == 0:
            res.append(L.orc_dec_bin(C.byref(d), C.byref(C.c_uint8.from_buffer(ci, int(b)))))
        elif k == 1:
            res.append(L.orc_dec_ep(C.byref(d)))
        elif k == 2:
            res.append(L.orc_dec_bins_ep(C.byref(d), int(b)))
        else:
            res.append(L.orc_dec_trm(C.byref(d)))
    return [int(x) for x in res]


# --------------------------------------------------------------------------
# context init
# --------------------------------------------------------------------------
def ctx_from_p0(p0) -> np.ndarray:
    L = lib()
    return np.array([L.orc_ctx_from_p0(float(p)) for p in np.asarray(p0, dtype=np.float64).reshape(-1)],
                    dtype=np.uint8)


def matlab_uint8(x) -> np.ndarray:
    L = lib()
    return np.array([L.orc_matlab_uint8(float(v)) for v in np.asarray(x, dtype=np.float64).reshape(-1)],
                    dtype=np.uint8)


# --------------------------------------------------------------------------
# binarizer & friends
# --------------------------------------------------------------------------
def binarize(v, Nq, method):
    b = np.zeros(4200, dtype=np.uint8)
    n = lib().orc_binarize(C.c_uint32(int(v)), C.c_uint32(int(Nq)), int(method), _p(b, _u8p))
    assert n >= 0
    return b[:n].copy()


def debinarize(bins, Nq, method) -> int:
    b = np.ascontiguousarray(bins, dtype=np.uint8)
    return int(lib().orc_debinarize(_p(b, _u8p), len(b), C.c_uint32(int(Nq)), int(method)))


def select_ctx(profile, n, g, up, Nlbp=3, types=0) -> int:
    g = np.ascontiguousarray(g, dtype=np.uint8)
    up = np.ascontiguousarray(up, dtype=np.uint8)
    gp = g if g.size else np.zeros(1, dtype=np.uint8)
    upp = up if up.size else np.zeros(1, dtype=np.uint8)
    return int(lib().orc_select_ctx(int(profile), int(n), _p(gp, _u8p), _p(upp, _u8p), int(up.size),
                                    int(Nlbp), C.c_uint(int(types))))


def num_ctx(profile, Nlbp=3) -> int:
    return int(lib().orc_profile_num_ctx(int(profile), int(Nlbp)))


def make_cfg(profile, method, Nq, Nlbp=3, types=0, rows=0) -> SymCfg:
    return SymCfg(int(profile), int(method), int(Nq), int(Nlbp), int(types), int(rows))


def symbols_to_ops(cfg: SymCfg, symbols) -> np.ndarray:
    sym = np.ascontiguousarray(symbols, dtype=np.uint32)
    L = lib()
    n = L.orc_symbols_to_ops(C.byref(cfg), _p(sym, _u32p), C.c_uint64(len(sym)), None, C.c_uint64(0))
    ops = np.zeros(max(int(n), 1), dtype=np.uint8)
    L.orc_symbols_to_ops(C.byref(cfg), _p(sym, _u32p), C.c_uint64(len(sym)), _p(ops, _u8p), C.c_uint64(int(n)))
    return ops[:int(n)]


def encode_symbols(cfg: SymCfg, symbols, sym_off, ctx_init, out_stride, n_threads=1, want_bits=False):
    sym = np.ascontiguousarray(symbols, dtype=np.uint32)
    sym_off = np.ascontiguousarray(sym_off, dtype=np.uint64)
    n = len(sym_off) - 1
    ci, n_ctx, per = _prep_ctx(ctx_init, n)
    slab = np.zeros((n, out_stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint32)
    bits = np.zeros(max(len(sym), 1), dtype=np.uint32) if want_bits else None
    rc = lib().orc_encode_symbols(C.byref(cfg), C.c_uint32(n), _p(sym_off, _u64p), _p(sym, _u32p),
                                  _p(ci, _u8p), C.c_uint32(n_ctx), per, _p(slab, _u8p),
                                  C.c_uint64(out_stride), _p(lens, _u32p),
                                  _p(bits, _u32p) if want_bits else None, n_threads)
    assert rc == 0 and (n == 0 or int(lens.max()) <= out_stride)
    if want_bits:
        return slab, lens, bits[:len(sym)]
    return slab, lens


def decode_symbols(cfg: SymCfg, payload, byte_off, sym_off, ctx_init, n_threads=1):
    sym_off = np.ascontiguousarray(sym_off, dtype=np.uint64)
    byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    if payload.size == 0:
        payload = np.zeros(1, dtype=np.uint8)
    n = len(sym_off) - 1
    ci, n_ctx, per = _prep_ctx(ctx_init, n)
    out = np.zeros(max(int(sym_off[-1]), 1), dtype=np.uint32)
    ok = np.zeros(n, dtype=np.uint8)
    rc = lib().orc_decode_symbols(C.byref(cfg), C.c_uint32(n), _p(byte_off, _u64p), _p(payload, _u8p),
                                  _p(sym_off, _u64p), _p(ci, _u8p), C.c_uint32(n_ctx), per,
                                  _p(out, _u32p), _p(ok, _u8p), n_threads)
    assert rc == 0
    return out[:int(sym_off[-1])], ok


def iss_ctx_init(G, Nq, method, Nlbp=3, types=CM_COND0 | CM_COND1 | CM_CONDS0 | CM_CONDS1):
    """G: 2-D (rows x cols) integer matrix -> p(0) per context (7*Nlbp+2 doubles)."""
    G = np.asarray(G)
    rows, cols = G.shape
    flat = np.ascontiguousarray(G.T.reshape(-1), dtype=np.uint32)  # column-major
    p0 = np.zeros(7 * Nlbp + 2, dtype=np.float64)
    lib().orc_iss_ctx_init(_p(flat, _u32p), C.c_uint32(rows), C.c_uint32(cols), C.c_uint32(int(Nq)),
                           int(method), int(Nlbp), C.c_uint(int(types)), p0.ctypes.data_as(C.POINTER(C.c_double)))
    return p0


# --------------------------------------------------------------------------
# the xorshift64 op recipe of SURVEY.md 4.1 (K7/K8 and bulk random tests)
# --------------------------------------------------------------------------
def xorshift_ops(k: int, n_ops: int = 64) -> np.ndarray:
    M = (1 << 64) - 1
    s = (0x9E3779B97F4A7C15 * (k + 1)) & M
    ops = np.zeros(n_ops, dtype=np.uint8)
    for i in range(n_ops):
        s ^= (s << 13) & M
        s ^= s >> 7
        s ^= (s << 17) & M
        r = (s >> 11) & 0xFFFFFFFF
        if (r & 3) == 0:
            ops[i] = (OP8_EP << 1) | ((r >> 2) & 1)
        else:
            ops[i] = (((r >> 2) & 3) << 1) | (1 if ((r >> 4) % 100) < 15 else 0)
    return ops


# --------------------------------------------------------------------------
# the unmodified reference mexFunction (stub mex.h)
# --------------------------------------------------------------------------
class _MexArg(C.Structure):
    _fields_ = [("is_char", C.c_int), ("s", C.c_char_p), ("d", C.POINTER(C.c_double)),
                ("m", C.c_int), ("n", C.c_int)]


def mex_call(nlhs, *args, trace_build=False):
    """Drive the reference mexFunction: args are str or array-likes of doubles.
    -> (rc, outputs list[float], error text)  rc 1 = mexErrMsgTxt raised."""
    m = ref_mex_trace() if trace_build else ref_mex()
    assert m is not None, "oracle/_ref/libref_mex*.so not built"
    keep = []
    arr = (_MexArg * max(len(args), 1))()
    for i, a in enumerate(args):
        if isinstance(a, str):
            b = a.encode()
            keep.append(b)
            arr[i] = _MexArg(1, b, None, 1, len(b))
        else:
            d = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            shape = d.shape if d.ndim == 2 else (1, d.size)
            d = np.asfortranarray(d.reshape(shape)).reshape(-1, order="F").copy()
            keep.append(d)
            arr[i] = _MexArg(0, None, d.ctypes.data_as(C.POINTER(C.c_double)), shape[0], shape[1])
    out = (C.c_double * 16)()
    out_n = C.c_int(0)
    err = C.create_string_buffer(512)
    rc = m.refmex_call(int(nlhs), out, 16, C.byref(out_n), len(args), arr, err, 512)
    return rc, [out[i] for i in range(out_n.value)], err.value.decode(errors="replace")


def mex_stats(handle: float, ctx_idx: int, decoder: bool = False, cap_steps: int = 1 << 20):
    """getEncoderStats / getDecoderStats of the trace build -> (steps uint8 [M, 5], trans uint32 [128, 128]
    indexed [state_p, state_a] in the reference's memory order)."""
    m = ref_mex_trace()
    assert m is not None
    m.refmex_stats.restype = C.c_long
    steps = np.zeros((cap_steps, 5), dtype=np.uint8)
    trans = np.zeros((128, 128), dtype=np.uint32)
    err = C.create_string_buffer(512)
    n = m.refmex_stats(int(bool(decoder)), C.c_double(handle), int(ctx_idx), _p(steps, _u8p), C.c_long(cap_steps),
                       _p(trans, _u32p), err, 512)
    if n < 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return steps[:n].copy(), trans


# pure-Python restatement of the trace members (ContextModel.cpp:97-134, SimpleCABACMex.cpp:231-241,
# 318-327): state sequence of ONE context given the bins coded with it.  decoder=True mirrors the
# reference decoder's log, which records the bin variable BEFORE decodeBin fills it in (always 0).
def trace_context(init_byte: int, bins, decoder: bool = False):
    L = lib()
    st = int(init_byte) & 127
    steps = np.zeros((len(bins), 5), dtype=np.uint8)
    trans = np.zeros((128, 128), dtype=np.uint32)
    hist = np.zeros(128, dtype=np.uint64)
    ts = lambda b: (b >> 1) + 64 if (b & 1) else 63 - (b >> 1)
    for i, b in enumerate(bins):
        nxt = int(L.orc_ctx_next_mps(st)) if (int(b) & 1) == (st & 1) else int(L.orc_ctx_next_lps(st))
        # addCabacStep gets the raw state numbers and logs the TRACE states (ContextModel.cpp:128-132)
        steps[i] = (0 if decoder else int(b), ts(st), st & 1, ts(nxt), nxt & 1)
        hist[ts(st)] += 1
        trans[ts(st), ts(nxt)] += 1
        st = nxt
    return steps, trans, hist


def ref_prob_to_state(p0) -> np.ndarray:
    m = ref_mex()
    assert m is not None
    p = np.ascontiguousarray(p0, dtype=np.float64).reshape(-1)
    out = np.zeros(len(p), dtype=np.uint8)
    for i in range(0, len(p), 900):
        chunk = np.ascontiguousarray(p[i:i + 900])
        o = np.zeros(len(chunk), dtype=np.uint8)
        m.refmex_prob_to_state(chunk.ctypes.data_as(C.POINTER(C.c_double)), len(chunk), _p(o, _u8p))
        out[i:i + 900] = o
    return out
