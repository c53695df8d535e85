"""ISS application coders on the GPU: mirrors of ISS/+coder/cabacEncode.m and cabacDecode.m
(same names, argument meaning and results), with the per-bin MATLAB loops replaced by the fused
device kernels:

  cabacEncode(G, Nq, param)            -> nbits, ctxInit0      (cabacEncode.m:1; writes param['fn'])
  cabacDecode(Nq, param, ctxInit, siz) -> G                    (cabacDecode.m:1; reads  param['fn'])

`param` is a dict with the fields of ISS.m:46-57: binMethod ('DEC2EG0'), cmTypes (list of
'cond0'/'cond1'/'condbinlft'/'conds0'/'conds1'), Nlbp, equalProb, fn, DEMO.  One matrix = one
stream, column-major (k outer, d inner), one shared context set -- exactly what the reference
writes.  encode_matrices / decode_matrices batch many matrices (or, with per_column=True, one
stream per column with the matrix's statistics, the north-star partitioning) into one launch.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as E


def _cfg(Nq, param, rows):
    method = param.get("binMethod", "DEC2EG0")
    if int(Nq) <= 2:
        # cabacEncode.m:16 / cabacDecode.m:62: with two levels nothing is binarised, every symbol IS its one
        # bin -- which is what the truncated-unary code of two levels produces (0 -> '0', 1 -> '1', no
        # terminating zero at Nq-1, cabacBinarizer.m:31-32), whatever binMethod says
        method, Nq = "DEC2TU", 2
    return E.make_cfg(E.PROFILE_ISS, method, int(Nq), int(param.get("Nlbp", 3)),
                      list(param.get("cmTypes", ["cond0", "cond1", "conds0", "conds1"])), rows=int(rows))


def encode_matrices(Gs, Nq, param, per_column: bool = False, want_heat: bool = False):
    """Gs: list of 2-D integer arrays with the same number of rows.  -> dict(payload, byte_off, ctxInit0
    [n_mat, 7N+2] uint8, ctx_state, sym_off, cfg, (heat: bits spent per symbol))."""
    rows = int(Gs[0].shape[0])
    assert all(int(g.shape[0]) == rows for g in Gs), "matrices of one batch share their row count"
    cfg = _cfg(Nq, param, rows)
    cols = [int(g.shape[1]) for g in Gs]
    sym = np.concatenate([np.asarray(g).T.reshape(-1) for g in Gs]).astype(np.uint32)   # column-major, cabacEncode.m:45-46
    if per_column:
        counts = np.full(sum(cols), rows, dtype=np.int64)
        group_of_stream = np.repeat(np.arange(len(Gs)), cols)
    else:
        counts = np.array([rows * c for c in cols], dtype=np.int64)
        group_of_stream = np.arange(len(Gs))
    sym_off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=sym_off[1:])
    d_sym = torch.as_tensor(sym.view(np.int32), device="cuda")
    d_off = torch.as_tensor(sym_off, device="cuda")
    # statistics per matrix (cabacEncode.m:23), whatever the stream partitioning
    mat_off = np.zeros(len(Gs) + 1, dtype=np.int64)
    np.cumsum([rows * c for c in cols], out=mat_off[1:])
    cnt = E.iss_ctx_stats(cfg, d_sym, torch.as_tensor(mat_off, device="cuda"), 1)
    # counters -> uint8 side information -> state bytes on the device (no host round trip in front of the encoder)
    _, d_q, d_st = E.iss_ctx_from_counters_device(cfg, cnt, bool(param.get("equalProb", False)))
    ctx_rows = d_st[torch.as_tensor(group_of_stream, device="cuda")]   # per-stream context init
    bins_bound = int(counts.max()) * (67 if cfg.method != E.BIN_TU else max(int(Nq), 2))
    stride = E.slab_stride_bound(bins_bound)
    res = E.encode_symbols(cfg, d_sym, d_off, ctx_rows, slab_stride=stride, want_bits=want_heat)
    enc, bits = res if want_heat else (res, None)
    pay = E.compact(enc)
    torch.cuda.synchronize()
    enc.check_overflow()
    q, st = d_q.cpu().numpy(), d_st.cpu().numpy()
    out = dict(payload=pay.payload, byte_off=pay.byte_off, ctxInit0=q, ctx_state=st, sym_off=sym_off, cfg=cfg,
               group_of_stream=group_of_stream)
    if want_heat:
        b = bits.cpu().numpy().astype(np.int64)
        prev = np.concatenate([[0], b[:-1]])
        prev[sym_off[:-1][counts > 0]] = 0                  # getNumBits() restarts with every stream here
        out["heat"] = b - prev                              # H(d,k) of cabacEncode.m:49,67
    return out


def decode_matrices(enc, shapes, per_column: bool = False):
    """Inverse of encode_matrices: -> list of 2-D arrays."""
    cfg = enc["cfg"]
    ctx_rows = enc["ctx_state"][enc["group_of_stream"]]
    sym, ok = E.decode_symbols(cfg, E.Payload(enc["payload"], enc["byte_off"]), enc["sym_off"], ctx_rows)
    if not bool(ok.all().item()):
        raise E.CabacError(-8, "bitstream not terminated properly")
    flat = sym.cpu().numpy().view(np.uint32)
    out, a = [], 0
    for (r, c) in shapes:
        out.append(flat[a:a + r * c].reshape(c, r).T.copy())
        a += r * c
    return out


def to_container(enc, sym_width: int = 4) -> np.ndarray:
    """encode_matrices result -> one buffer (container.py): payload, offset table, symbols per stream and
    the uint8 side information ctxInit0 of every stream (what the reference ships in a .mat file,
    ISS/ISS.m:197-201); any stream of it is a bitstream the reference decodes unchanged."""
    from . import container as K
    q_rows = np.ascontiguousarray(enc["ctxInit0"][enc["group_of_stream"]], dtype=np.uint8)
    return K.pack(enc["payload"], enc["byte_off"], q_rows, unit_off=enc["sym_off"], cfg=enc["cfg"],
                  sym_width=sym_width, ctx_is_prob=True)


def from_container(blob):
    """Inverse of to_container: -> the dict decode_matrices takes (context states re-derived from the
    uint8 side information exactly like cabacDecode.m:13 does)."""
    from . import container as K
    c = K.unpack(blob)
    q = c.ctx_init if c.ctx_init.ndim == 2 else c.ctx_init[None, :].repeat(max(c.n_streams, 1), 0)
    st = E.ctx_from_prob(q.astype(np.float64).reshape(-1) / 255.0).reshape(q.shape)
    return dict(payload=torch.as_tensor(c.payload, device="cuda") if c.payload.size else torch.zeros(1, dtype=torch.uint8, device="cuda"),
                byte_off=torch.as_tensor(c.byte_off.astype(np.int64), device="cuda"),
                ctxInit0=q, ctx_state=st, sym_off=c.unit_off.astype(np.int64), cfg=c.cfg,
                group_of_stream=np.arange(c.n_streams))


def cabacEncode(G, Nq=2, param=None):
    """ISS/+coder/cabacEncode.m: -> (nbits, ctxInit0).  The bitstream goes to param['fn']."""
    param = dict(param or {})
    G = np.asarray(G)
    enc = encode_matrices([G], Nq, param, want_heat=bool(param.get("DEMO", 0)))
    data = enc["payload"].cpu().numpy().tobytes()
    with open(param["fn"], "wb") as f:
        f.write(data)
    nbits = 8 * len(data)                                   # dir(fn).bytes*8, cabacEncode.m:100-101
    if param.get("DEMO", 0):
        param["_H"] = enc["heat"].reshape(G.shape[1], G.shape[0]).T
    return nbits, enc["ctxInit0"][0]


def cabacDecode(Nq, param, ctxInit, siz):
    """ISS/+coder/cabacDecode.m: ctxInit is the uint8 side information of cabacEncode (ctxInit0)."""
    param = dict(param or {})
    rows, cols = int(siz[0]), int(siz[1])
    cfg = _cfg(Nq, param, rows)
    p = np.asarray(ctxInit, dtype=np.float64) / 255.0       # cabacDecode.m:13
    st = E.ctx_from_prob(p)
    data = np.fromfile(param["fn"], dtype=np.uint8)
    boff = np.array([0, len(data)], dtype=np.int64)
    sym, ok = E.decode_symbols(cfg, (data, boff), np.array([0, rows * cols], dtype=np.int64), st)
    if not bool(ok.all().item()):
        raise E.CabacError(-8, "bitstream not terminated properly")   # Decoder::finish() asserts
    return sym.cpu().numpy().view(np.uint32).reshape(cols, rows).T.copy()
