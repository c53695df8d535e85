"""Device-tensor front end of the batch C ABI (include/isscabac.h).

torch is used for what the task brief assigns to it: device memory, streams and (in
multi_gpu.py) torch.distributed.  All coding work happens in libisscabac.so's CUDA kernels;
there is no CPU path -- on a machine without a GPU these functions raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from ._lib import CabacError, SymCfg, check, lib, vp

BIN_TU, BIN_EG0, BIN_EG1, BIN_EG2, BIN_FL32, BIN_TR0, BIN_TR1, BIN_TR2 = range(8)
PROFILE_DEMO, PROFILE_ISS, PROFILE_FLAT, PROFILE_FLAT_EPSUF = range(4)
CM_COND0, CM_COND1, CM_CONDBINLFT, CM_CONDS0, CM_CONDS1 = 1, 2, 4, 8, 16
OP8_TRM, OP8_EP, OP16_TRM, OP16_EP = 125, 126, 0x7FFD, 0x7FFE
METHODS = {"DEC2TU": BIN_TU, "DEC2EG0": BIN_EG0, "DEC2EG1": BIN_EG1, "DEC2EG2": BIN_EG2, "DEC2FL32": BIN_FL32,
           "DEC2TR0": BIN_TR0, "DEC2TR1": BIN_TR1, "DEC2TR2": BIN_TR2}   # truncated Rice: binarize / encode only (as upstream)
CM_TYPES = {"cond0": CM_COND0, "cond1": CM_COND1, "condbinlft": CM_CONDBINLFT, "conds0": CM_CONDS0, "conds1": CM_CONDS1}


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise CabacError(-2, "no CUDA device: isscabac_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(x, dtype, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=dtype)
    else:
        t = torch.as_tensor(np.ascontiguousarray(x), device=device).to(dtype)
    return t.contiguous()


def _ops_tensor(ops, device) -> tuple[torch.Tensor, int]:
    if isinstance(ops, torch.Tensor):
        t = ops.to(device).contiguous()
    else:
        a = np.ascontiguousarray(ops)
        if a.dtype == np.uint16:  # torch has limited uint16 support: ship the raw bytes
            return torch.as_tensor(a.view(np.uint8), device=device), 2
        t = torch.as_tensor(a, device=device)
    if t.dtype == torch.uint8:
        return t, 1
    if t.dtype in (torch.int16, torch.uint16):
        return t.view(torch.uint8), 2
    raise TypeError("ops must be uint8 (u8 op format) or uint16/int16 (u16 op format)")


def _ctx_tensor(ctx_init, n_streams, device) -> tuple[torch.Tensor, int, int]:
    t = _dev(ctx_init, torch.uint8, device)
    if t.dim() == 2:
        if t.shape[0] != n_streams:
            raise ValueError("per-stream ctx_init must have one row per stream")
        return t, int(t.shape[1]), 1
    return t, int(t.numel()), 0


def ctx_from_prob(p0) -> np.ndarray:
    """p(0) per context -> state bytes (CABAC_ContextModelsInit.cpp:124-148)."""
    p = np.ascontiguousarray(p0, dtype=np.float64).reshape(-1)
    out = np.zeros(p.size, dtype=np.uint8)
    check(lib().cabac_ctx_from_prob(vp(p), C.c_uint32(p.size), vp(out)))
    return out


def ctx_from_state(triples) -> np.ndarray:
    """[ctxIdx mps state] triples (3 x N column-major or N x 3 row-major flattened) -> state bytes."""
    t = np.ascontiguousarray(triples, dtype=np.float64).reshape(-1)
    n = t.size // 3
    out = np.zeros(n, dtype=np.uint8)
    check(lib().cabac_ctx_from_state(vp(t), C.c_uint32(n), vp(out)))
    return out


def profile_num_ctx(profile: int, Nlbp: int = 3) -> int:
    return int(lib().cabac_profile_num_ctx(int(profile), int(Nlbp)))


def slab_stride_bound(max_ops: int) -> int:
    return int(lib().cabac_slab_stride_bound(C.c_uint64(int(max_ops))))


@dataclass
class Encoded:
    """Result of a batch encode: per-stream slab rows + lengths (device tensors)."""
    slab: torch.Tensor        # u8 [n_streams, stride]
    lengths: torch.Tensor     # int32 [n_streams] (u32 values)
    overflow: torch.Tensor    # int32 [4]; bit 0 set = some stream exceeded the stride

    def check_overflow(self):
        if int(self.overflow[0].item()) & 1:
            raise CabacError(-3, "a stream exceeded slab_stride; retry with a larger stride")


def encode_ops(ops, op_off, ctx_init, slab_stride: int | None = None, *, out: Encoded | None = None) -> Encoded:
    """start(); ops...; finish() for every stream s = ops[op_off[s]:op_off[s+1]] (device-resident)."""
    dev = _require_cuda()
    ops_t, width = _ops_tensor(ops, dev)
    off_t = _dev(op_off, torch.int64, dev)
    n = off_t.numel() - 1
    ctx_t, n_ctx, per = _ctx_tensor(ctx_init, n, dev)
    if out is None:
        if slab_stride is None:
            longest = int((off_t[1:] - off_t[:-1]).max().item()) if n else 0
            slab_stride = (longest // 4 + 64 + 15) & ~15
        slab_stride = (int(slab_stride) + 15) & ~15
        out = Encoded(torch.empty((n, slab_stride), dtype=torch.uint8, device=dev),
                      torch.empty(n, dtype=torch.int32, device=dev),
                      torch.zeros(4, dtype=torch.int32, device=dev))
    check(lib().cabac_encode_ops(C.c_uint32(n), vp(off_t), vp(ops_t), width, vp(ctx_t), C.c_uint32(n_ctx), per,
                                 vp(out.slab), C.c_uint64(out.slab.shape[1]), vp(out.lengths), vp(out.overflow),
                                 _stream_ptr()))
    return out


@dataclass
class Payload:
    """One contiguous bitstream + offset table: stream s = payload[byte_off[s]:byte_off[s+1]]."""
    payload: torch.Tensor     # u8 [total]
    byte_off: torch.Tensor    # int64 [n_streams+1]


def compact(enc: Encoded, payload_cap: int | None = None, *, payload: torch.Tensor | None = None,
            byte_off: torch.Tensor | None = None, scratch: torch.Tensor | None = None) -> Payload:
    """Device-wide exclusive scan over the lengths + copy into one contiguous bitstream."""
    dev = _require_cuda()
    n = enc.lengths.numel()
    L = lib()
    if scratch is None:
        scratch = torch.empty(int(L.cabac_compact_scratch_bytes(C.c_uint32(n))), dtype=torch.uint8, device=dev)
    if byte_off is None:
        byte_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    if payload is None:
        if payload_cap is None:
            # offsets first: the total sizes the payload exactly (one host sync)
            check(L.cabac_compact(C.c_uint32(n), vp(enc.slab), C.c_uint64(enc.slab.shape[1]), vp(enc.lengths),
                                  None, C.c_uint64(0), vp(byte_off), vp(scratch), vp(enc.overflow), _stream_ptr()))
            payload_cap = int(byte_off[-1].item())
        # the decoders read the payload in aligned 32-bit words (include/isscabac.h): the allocation covers the word that
        # holds the last byte, the tensor handed out is the exact size
        payload = torch.empty((max(int(payload_cap), 1) + 3) & ~3, dtype=torch.uint8, device=dev)[:max(int(payload_cap), 1)]
    check(L.cabac_compact(C.c_uint32(n), vp(enc.slab), C.c_uint64(enc.slab.shape[1]), vp(enc.lengths),
                          vp(payload), C.c_uint64(payload.numel()), vp(byte_off), vp(scratch), vp(enc.overflow),
                          _stream_ptr()))
    return Payload(payload, byte_off)


def decode_ops(payload: Payload | tuple, ops, op_off, ctx_init, *, bins: torch.Tensor | None = None,
               finish_ok: torch.Tensor | None = None):
    """-> (bins u8[n_ops], finish_ok u8[n_streams]); bit 0 of every op is ignored."""
    dev = _require_cuda()
    if isinstance(payload, Payload):
        pay, boff = payload.payload, payload.byte_off
    else:
        pay, boff = _dev(payload[0], torch.uint8, dev), _dev(payload[1], torch.int64, dev)
    ops_t, width = _ops_tensor(ops, dev)
    off_t = _dev(op_off, torch.int64, dev)
    n = off_t.numel() - 1
    ctx_t, n_ctx, per = _ctx_tensor(ctx_init, n, dev)
    n_ops = ops_t.numel() // width
    if bins is None:
        bins = torch.empty(max(n_ops, 1), dtype=torch.uint8, device=dev)
    if finish_ok is None:
        finish_ok = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    if pay.numel() == 0:
        pay = torch.zeros(1, dtype=torch.uint8, device=dev)
    check(lib().cabac_decode_ops(C.c_uint32(n), vp(boff), vp(pay), vp(off_t), vp(ops_t), width, vp(ctx_t),
                                 C.c_uint32(n_ctx), per, vp(bins), vp(finish_ok), _stream_ptr()))
    return bins[:n_ops], finish_ok[:n]


# ------------------------------------------------------------------------------------
# symbol level
# ------------------------------------------------------------------------------------
def make_cfg(profile, method, Nq, Nlbp=3, types=0, rows=0) -> SymCfg:
    if isinstance(method, str):
        method = METHODS[method]
    if not isinstance(types, int):
        m = 0
        for t in types:
            m |= CM_TYPES[t]
        types = m
    return SymCfg(int(profile), int(method), int(Nq), int(Nlbp), int(types), int(rows))


def _sym_tensor(symbols, dev) -> tuple[torch.Tensor, int]:
    if isinstance(symbols, torch.Tensor):
        t = symbols.to(dev).contiguous()
        if t.dtype == torch.uint8:
            return t, 1
        if t.dtype in (torch.int16, torch.uint16):
            return t, 2
        if t.dtype in (torch.int32, torch.uint32):
            return t, 4
        return t.to(torch.int32), 4
    a = np.ascontiguousarray(symbols)
    if a.dtype == np.uint8:
        return torch.as_tensor(a, device=dev), 1
    if a.dtype == np.uint16:
        return torch.as_tensor(a.view(np.int16), device=dev), 2
    return torch.as_tensor(a.astype(np.uint32).view(np.int32), device=dev), 4


def binarize_symbols(cfg: SymCfg, symbols, sym_off, ops=None, op_off=None, scratch=None):
    """Vectorised binarizer + context selector: symbols -> (ops u8, op_off int64[n_streams+1]).
    Without `ops`: the two-call API (sizing call, one host read of the total, then the ops into a buffer of that size).
    With `ops` (a device u8 buffer whose size the caller knows to be enough, e.g. from an earlier sizing call): ONE
    stream-ordered call, no host read; no op is written past the buffer and op_off[-1] tells the true total."""
    dev = _require_cuda()
    sym_t, width = _sym_tensor(symbols, dev)
    off_t = _dev(sym_off, torch.int64, dev)
    n = off_t.numel() - 1
    n_sym = sym_t.numel()
    L = lib()
    if scratch is None:
        scratch = torch.empty(int(L.cabac_binarize_scratch_bytes(C.c_uint64(n_sym), C.c_uint32(n))), dtype=torch.uint8, device=dev)
    if op_off is None:
        op_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    if ops is not None:
        check(L.cabac_binarize_symbols(C.byref(cfg), C.c_uint32(n), vp(off_t), vp(sym_t), width, C.c_uint64(n_sym),
                                       vp(op_off), vp(ops), C.c_uint64(ops.numel()), vp(scratch), _stream_ptr()))
        return ops, op_off
    check(L.cabac_binarize_symbols(C.byref(cfg), C.c_uint32(n), vp(off_t), vp(sym_t), width, C.c_uint64(n_sym),
                                   vp(op_off), None, C.c_uint64(0), vp(scratch), _stream_ptr()))
    total = int(op_off[-1].item())
    ops = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    check(L.cabac_binarize_symbols(C.byref(cfg), C.c_uint32(n), vp(off_t), vp(sym_t), width, C.c_uint64(n_sym),
                                   vp(op_off), vp(ops), C.c_uint64(total), vp(scratch), _stream_ptr()))
    return ops[:total], op_off


def encode_symbols(cfg: SymCfg, symbols, sym_off, ctx_init, slab_stride: int, want_bits: bool = False):
    """Fused binarize + context select + encode.  -> Encoded (and bits_after_symbol if asked)."""
    dev = _require_cuda()
    sym_t, width = _sym_tensor(symbols, dev)
    off_t = _dev(sym_off, torch.int64, dev)
    n = off_t.numel() - 1
    ctx_t, n_ctx, per = _ctx_tensor(ctx_init, n, dev)
    slab_stride = (int(slab_stride) + 15) & ~15
    out = Encoded(torch.empty((n, slab_stride), dtype=torch.uint8, device=dev),
                  torch.empty(n, dtype=torch.int32, device=dev), torch.zeros(4, dtype=torch.int32, device=dev))
    bits = torch.empty(max(sym_t.numel(), 1), dtype=torch.int32, device=dev) if want_bits else None
    check(lib().cabac_encode_symbols(C.byref(cfg), C.c_uint32(n), vp(off_t), vp(sym_t), width, vp(ctx_t),
                                     C.c_uint32(n_ctx), per, vp(out.slab), C.c_uint64(slab_stride), vp(out.lengths),
                                     vp(bits), vp(out.overflow), _stream_ptr()))
    return (out, bits[:sym_t.numel()]) if want_bits else out


def decode_symbols(cfg: SymCfg, payload: Payload | tuple, sym_off, ctx_init, sym_dtype=torch.int32):
    """-> (symbols, finish_ok)."""
    dev = _require_cuda()
    if isinstance(payload, Payload):
        pay, boff = payload.payload, payload.byte_off
    else:
        pay, boff = _dev(payload[0], torch.uint8, dev), _dev(payload[1], torch.int64, dev)
    off_t = _dev(sym_off, torch.int64, dev)
    n = off_t.numel() - 1
    ctx_t, n_ctx, per = _ctx_tensor(ctx_init, n, dev)
    n_sym = int(off_t[-1].item()) if n else 0
    width = {torch.uint8: 1, torch.int16: 2, torch.int32: 4}[sym_dtype]
    out = torch.empty(max(n_sym, 1), dtype=sym_dtype, device=dev)
    ok = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    if pay.numel() == 0:
        pay = torch.zeros(1, dtype=torch.uint8, device=dev)
    check(lib().cabac_decode_symbols(C.byref(cfg), C.c_uint32(n), vp(boff), vp(pay), vp(off_t), vp(ctx_t),
                                     C.c_uint32(n_ctx), per, vp(out), width, vp(ok), _stream_ptr()))
    return out[:n_sym], ok[:n]


# ------------------------------------------------------------------------------------
# host-buffer API (numpy in / numpy out; every copy happens inside the C call)
# ------------------------------------------------------------------------------------
def _np_ops(ops):
    a = np.ascontiguousarray(ops)
    if a.dtype == np.uint8:
        return a, 1
    if a.dtype == np.uint16:
        return a, 2
    raise TypeError("ops must be uint8 or uint16")


def _np_ctx(ctx_init, n):
    c = np.ascontiguousarray(ctx_init, dtype=np.uint8)
    if c.ndim == 2:
        if c.shape[0] != n:
            raise ValueError("per-stream ctx_init must have one row per stream")
        return c, c.shape[1], 1
    return c, c.size, 0


def encode_ops_host(ops, op_off, ctx_init, payload_out: np.ndarray | None = None, byte_off_out: np.ndarray | None = None):
    """-> (payload u8, byte_off u64[n+1]) with host buffers; see cabac_encode_ops_host."""
    _require_cuda()
    a, w = _np_ops(ops)
    off = np.ascontiguousarray(op_off, dtype=np.uint64)
    n = off.size - 1
    c, n_ctx, per = _np_ctx(ctx_init, n)
    if payload_out is None:
        payload_out = np.empty(int(off[-1] - off[0]) // 4 + 64 * max(n, 1) + 64, dtype=np.uint8)
    if byte_off_out is None:
        byte_off_out = np.empty(n + 1, dtype=np.uint64)
    check(lib().cabac_encode_ops_host(C.c_uint32(n), vp(off), vp(a), w, vp(c), C.c_uint32(n_ctx), per,
                                      vp(payload_out), C.c_uint64(payload_out.size), vp(byte_off_out)))
    return payload_out[:int(byte_off_out[-1])], byte_off_out


def decode_ops_host(payload, byte_off, ops, op_off, ctx_init, bins_out: np.ndarray | None = None, packed: bool = False):
    """-> (bins, finish_ok).  packed=True: bins come back bit-packed (bit i & 7 of byte i >> 3 = bin of op i,
    np.unpackbits(..., bitorder="little") restores one per byte) -- cabac_decode_ops_host_packed."""
    _require_cuda()
    a, w = _np_ops(ops)
    off = np.ascontiguousarray(op_off, dtype=np.uint64)
    boff = np.ascontiguousarray(byte_off, dtype=np.uint64)
    pay = np.ascontiguousarray(payload, dtype=np.uint8)
    n = off.size - 1
    c, n_ctx, per = _np_ctx(ctx_init, n)
    n_out = (a.size + 7) // 8 if packed else a.size
    if bins_out is None:
        bins_out = np.empty(max(n_out, 1), dtype=np.uint8)
    ok = np.zeros(max(n, 1), dtype=np.uint8)
    if pay.size == 0:
        pay = np.zeros(1, dtype=np.uint8)
    fn = lib().cabac_decode_ops_host_packed if packed else lib().cabac_decode_ops_host
    check(fn(C.c_uint32(n), vp(boff), vp(pay), vp(off), vp(a), w, vp(c), C.c_uint32(n_ctx), per, vp(bins_out), vp(ok)))
    return bins_out[:n_out], ok[:n]


def encode_symbols_host(cfg: SymCfg, symbols, sym_off, ctx_init, want_bits: bool = False, payload_out: np.ndarray | None = None,
                        byte_off_out: np.ndarray | None = None):
    """-> (payload u8, byte_off u64[n+1] (, bits_after_symbol)) with host buffers; see cabac_encode_symbols_host.
    payload_out / byte_off_out: caller-owned (e.g. pinned) result buffers."""
    _require_cuda()
    s = np.ascontiguousarray(symbols)
    if s.dtype not in (np.uint8, np.uint16, np.uint32):
        s = s.astype(np.uint32)
    off = np.ascontiguousarray(sym_off, dtype=np.uint64)
    n = off.size - 1
    c, n_ctx, per = _np_ctx(ctx_init, n)
    if cfg.method >= BIN_TR0:
        k = cfg.method - BIN_TR0
        per_sym = ((max(int(cfg.Nq), 2) - 1) >> k) + 1 + k if cfg.Nq else (1 << (8 * s.dtype.itemsize)) + k
    else:
        per_sym = 67 if cfg.method != BIN_TU else max(int(cfg.Nq), 2)
    payload = payload_out if payload_out is not None else np.empty(int(s.size) * per_sym // 8 + 16 * max(n, 1) + 64, dtype=np.uint8)
    boff = byte_off_out if byte_off_out is not None else np.empty(n + 1, dtype=np.uint64)
    bits = np.empty(max(s.size, 1), dtype=np.uint32) if want_bits else None
    check(lib().cabac_encode_symbols_host(C.byref(cfg), C.c_uint32(n), vp(off), vp(s), s.dtype.itemsize, vp(c),
                                          C.c_uint32(n_ctx), per, vp(payload), C.c_uint64(payload.size), vp(boff),
                                          vp(bits)))
    res = (payload[:int(boff[-1])], boff)
    return res + (bits[:s.size],) if want_bits else res


def decode_symbols_host(cfg: SymCfg, payload, byte_off, sym_off, ctx_init, dtype=np.uint32, out: np.ndarray | None = None):
    _require_cuda()
    off = np.ascontiguousarray(sym_off, dtype=np.uint64)
    boff = np.ascontiguousarray(byte_off, dtype=np.uint64)
    pay = np.ascontiguousarray(payload, dtype=np.uint8)
    if pay.size == 0:
        pay = np.zeros(1, dtype=np.uint8)
    n = off.size - 1
    c, n_ctx, per = _np_ctx(ctx_init, n)
    if out is None:
        out = np.empty(max(int(off[-1]), 1), dtype=dtype)
    ok = np.zeros(max(n, 1), dtype=np.uint8)
    check(lib().cabac_decode_symbols_host(C.byref(cfg), C.c_uint32(n), vp(boff), vp(pay), vp(off), vp(c),
                                          C.c_uint32(n_ctx), per, vp(out), out.dtype.itemsize, vp(ok)))
    return out[:int(off[-1])], ok[:n]


# ------------------------------------------------------------------------------------
# ISS context-init statistics (ISS/+coder/cabacInitContextModel.m)
# ------------------------------------------------------------------------------------
def iss_ctx_stats(cfg: SymCfg, symbols, sym_off, streams_per_group: int = 1) -> torch.Tensor:
    """Per-group counters (int64 [n_groups, K]) of the context-init statistics, reduced on the device."""
    dev = _require_cuda()
    sym_t, width = _sym_tensor(symbols, dev)
    off_t = _dev(sym_off, torch.int64, dev)
    n = off_t.numel() - 1
    L = lib()
    K = int(L.cabac_iss_num_counters(int(cfg.Nlbp)))
    if K < 0:
        raise CabacError(K, "Nlbp out of range")
    groups = (n + streams_per_group - 1) // max(streams_per_group, 1)
    cnt = torch.empty((max(groups, 1), K), dtype=torch.int64, device=dev)
    check(L.cabac_iss_ctx_stats(C.byref(cfg), C.c_uint32(n), vp(off_t), vp(sym_t), width, C.c_uint64(sym_t.numel()),
                                C.c_uint32(int(streams_per_group)), vp(cnt), _stream_ptr()))
    return cnt[:groups]


def iss_ctx_from_counters(cfg: SymCfg, counters, equal_prob: bool = False):
    """-> (p0 float64 [g, 7N+2], ctxInit0 uint8 [g, 7N+2] (the side information), state bytes uint8 [g, 7N+2])."""
    c = np.ascontiguousarray(counters.cpu().numpy() if isinstance(counters, torch.Tensor) else counters, dtype=np.uint64)
    g = c.shape[0]
    nctx = 7 * int(cfg.Nlbp) + 2
    p0 = np.zeros((g, nctx), dtype=np.float64)
    q = np.zeros((g, nctx), dtype=np.uint8)
    st = np.zeros((g, nctx), dtype=np.uint8)
    check(lib().cabac_iss_ctx_from_counters(C.byref(cfg), vp(c), C.c_uint32(g), int(bool(equal_prob)), vp(p0), vp(q), vp(st)))
    return p0, q, st


def iss_ctx_from_counters_device(cfg: SymCfg, counters: torch.Tensor, equal_prob: bool = False, want_p0: bool = False):
    """Device-side twin of iss_ctx_from_counters: -> (p0 or None, ctxInit0 uint8 [g, 7N+2], state bytes uint8 [g, 7N+2]) as
    device tensors, stream-ordered, no host synchronisation."""
    dev = _require_cuda()
    cnt = _dev(counters, torch.int64, dev)
    g = cnt.shape[0]
    nctx = 7 * int(cfg.Nlbp) + 2
    q = torch.empty((max(g, 1), nctx), dtype=torch.uint8, device=dev)
    st = torch.empty((max(g, 1), nctx), dtype=torch.uint8, device=dev)
    p0 = torch.empty((max(g, 1), nctx), dtype=torch.float64, device=dev) if want_p0 else None
    check(lib().cabac_iss_ctx_from_counters_device(C.byref(cfg), vp(cnt), C.c_uint32(g), int(bool(equal_prob)), vp(p0), vp(q), vp(st),
                                                   _stream_ptr()))
    return (p0[:g] if want_p0 else None), q[:g], st[:g]


# ------------------------------------------------------------------------------------
# statistics outputs (ctxHist / ctxCost of cabacEncode.m:40-65, trace members of ContextModel.cpp:97-134)
# ------------------------------------------------------------------------------------
@dataclass
class CtxTrace:
    state_hist: torch.Tensor             # int64 [groups, n_ctx, 128] visits per trace state before each update
    trans: torch.Tensor | None           # int32 [groups, n_ctx, 128, 128] (u32 values) [before][after]
    cost_bits: torch.Tensor | None       # int64 [groups, n_ctx + 1]; last slot = bypass + terminate bins
    step_states: torch.Tensor | None     # u8 [n_ops, 2] state byte before / after (0xFF for non-context ops)
    final_ctx: torch.Tensor | None       # u8 [n_streams, n_ctx]

    @property
    def usage(self) -> torch.Tensor:
        """bins coded per context (ctxHist of cabacEncode.m:40,61)"""
        return self.state_hist.sum(-1)


def trace_state(state_byte):
    """Trace-state index of a context state byte (ContextModel.cpp:99-101)."""
    b = np.asarray(state_byte).astype(np.int64)
    return np.where(b & 1, (b >> 1) + 64, 63 - (b >> 1))


def ctx_trace_ops(ops, op_off, ctx_init, streams_per_group: int = 1, want_trans: bool = False, want_cost: bool = False,
                  want_steps: bool = False, want_final: bool = False) -> CtxTrace:
    """One pass over the op arrays on the device; see cabac_ctx_trace_ops in include/isscabac.h."""
    dev = _require_cuda()
    ops_t, width = _ops_tensor(ops, dev)
    off_t = _dev(op_off, torch.int64, dev)
    n = off_t.numel() - 1
    ctx_t, n_ctx, per = _ctx_tensor(ctx_init, n, dev)
    spg = max(int(streams_per_group), 1)
    groups = max((n + spg - 1) // spg, 1)
    n_ops = ops_t.numel() // width
    hist = torch.empty((groups, n_ctx, 128), dtype=torch.int64, device=dev)
    trans = torch.empty((groups, n_ctx, 128, 128), dtype=torch.int32, device=dev) if want_trans else None
    cost = torch.empty((groups, n_ctx + 1), dtype=torch.int64, device=dev) if want_cost else None
    steps = torch.empty((max(n_ops, 1), 2), dtype=torch.uint8, device=dev) if want_steps else None
    final = torch.empty((max(n, 1), n_ctx), dtype=torch.uint8, device=dev) if want_final else None
    check(lib().cabac_ctx_trace_ops(C.c_uint32(n), vp(off_t), vp(ops_t), width, vp(ctx_t), C.c_uint32(n_ctx), per,
                                    C.c_uint32(spg), vp(hist), vp(trans), vp(cost), vp(steps), vp(final), _stream_ptr()))
    return CtxTrace(hist, trans, cost, steps[:n_ops] if want_steps else None, final[:n] if want_final else None)
