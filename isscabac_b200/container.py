"""Wire format for many CABAC streams: payload + u64 offset table + per-stream unit counts +
context-init bytes in one buffer (include/isscabac.h, `cabac_container_*`; layout in
csrc/container.cpp).  Host-side framing only -- no coding work happens here.

What it replaces in the reference: one bitstream FILE per stream (SimpleCABACMex.cpp:195 writes
it, :288 reads it) and the context initialisation handed to the decoder as uint8 side information
through a .mat file (ISS/ISS.m:197-201, ISS/+coder/cabacEncode.m:30, cabacDecode.m:13).  The bytes
of every stream are left exactly as the reference encoder would have written them, so
`Container.stream(s)` / `Container.write_stream_file(s, fn)` give something the reference's
decodeStart ... decodeFinish reads unchanged.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import SymCfg, check, lib, u8p, u64p


class ContainerView(C.Structure):
    """isscabac_container_view"""
    _fields_ = [("n_streams", C.c_uint32), ("n_ctx", C.c_uint32), ("per_stream_init", C.c_int32),
                ("ctx_is_prob", C.c_int32), ("has_cfg", C.c_int32), ("sym_width", C.c_int32), ("cfg", SymCfg),
                ("payload_bytes", C.c_uint64), ("byte_off", u64p), ("unit_off", u64p), ("ctx_init", u8p),
                ("payload", u8p)]


def _host(x, dtype):
    if x is None:
        return None
    if hasattr(x, "detach"):   # torch tensor, any device
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def _ptr(a, typ):
    return C.cast(C.c_void_p(a.ctypes.data if a is not None and a.size else 0), typ)


@dataclass
class Container:
    byte_off: np.ndarray              # uint64 [n+1]
    payload: np.ndarray               # uint8 [payload_bytes]
    ctx_init: np.ndarray              # uint8 [n_ctx] or [n, n_ctx]
    unit_off: np.ndarray | None       # uint64 [n+1] symbols (or ops) per stream, exclusive offsets
    cfg: SymCfg | None
    sym_width: int
    ctx_is_prob: bool

    @property
    def n_streams(self) -> int:
        return int(self.byte_off.size - 1)

    def stream(self, s: int) -> np.ndarray:
        """Bytes of stream s: the file the reference would have written for it."""
        return self.payload[int(self.byte_off[s]):int(self.byte_off[s + 1])]

    def stream_ctx(self, s: int) -> np.ndarray:
        return self.ctx_init[s] if self.ctx_init.ndim == 2 else self.ctx_init

    def write_stream_file(self, s: int, fn: str) -> None:
        self.stream(s).tofile(fn)


def pack(payload, byte_off, ctx_init, unit_off=None, cfg: SymCfg | None = None, sym_width: int = 0,
         ctx_is_prob: bool = False) -> np.ndarray:
    """-> uint8 array holding the container.  Inputs may be numpy arrays or torch tensors (a device
    payload is copied to the host here)."""
    boff = _host(byte_off, np.uint64)
    n = boff.size - 1
    total = int(boff[-1]) if n >= 0 else 0
    pay = _host(payload, np.uint8).reshape(-1)
    if pay.size < total:
        raise ValueError(f"payload holds {pay.size} bytes but byte_off[-1] = {total}")
    pay = pay[:total]
    units = _host(unit_off, np.uint64)
    if units is not None and units.size != boff.size:
        raise ValueError("unit_off must have one entry per byte_off entry")
    ctx = _host(ctx_init, np.uint8)
    v = ContainerView()
    v.n_streams = n
    v.n_ctx = int(ctx.shape[-1]) if ctx.ndim else 0
    v.per_stream_init = 1 if ctx.ndim == 2 else 0
    if ctx.ndim == 2 and ctx.shape[0] != n:
        raise ValueError("per-stream ctx_init must have one row per stream")
    v.ctx_is_prob = int(bool(ctx_is_prob))
    v.has_cfg = 0 if cfg is None else 1
    v.sym_width = int(sym_width)
    if cfg is not None:
        v.cfg = cfg
    v.payload_bytes = total
    v.byte_off = _ptr(boff, u64p)
    v.unit_off = _ptr(units, u64p)
    v.ctx_init = _ptr(ctx, u8p)
    v.payload = _ptr(pay, u8p)
    L = lib()
    L.cabac_container_size.restype = C.c_uint64
    size = int(L.cabac_container_size(C.byref(v)))
    out = np.zeros(max(size, 8) // 8 + 1, dtype=np.uint64).view(np.uint8)[:max(size, 8)]   # 8-byte aligned
    written = C.c_uint64(0)
    check(L.cabac_container_write(C.byref(v), _ptr(out, u8p), C.c_uint64(out.size), C.byref(written)))
    return out[:int(written.value)]


def unpack(buf, verify_payload_crc: bool = True) -> Container:
    """Parse + validate a container.  The returned arrays are copies (independent of `buf`)."""
    raw = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1)
    al = np.zeros(raw.size // 8 + 1, dtype=np.uint64).view(np.uint8)[:raw.size]
    al[:] = raw
    v = ContainerView()
    check(lib().cabac_container_parse(_ptr(al, u8p), C.c_uint64(al.size), int(bool(verify_payload_crc)), C.byref(v)))
    n = int(v.n_streams)
    boff = np.ctypeslib.as_array(v.byte_off, shape=(n + 1,)).copy()
    units = np.ctypeslib.as_array(v.unit_off, shape=(n + 1,)).copy() if v.unit_off else None
    nb = int(v.payload_bytes)
    pay = np.ctypeslib.as_array(v.payload, shape=(nb,)).copy() if nb else np.zeros(0, np.uint8)
    n_ctx = int(v.n_ctx)
    if n_ctx and (n or not v.per_stream_init):
        shape = (n, n_ctx) if v.per_stream_init else (n_ctx,)
        ctx = np.ctypeslib.as_array(v.ctx_init, shape=shape).copy()
    else:
        ctx = np.zeros((n, n_ctx) if v.per_stream_init else (n_ctx,), np.uint8)
    cfg = None
    if v.has_cfg:
        cfg = SymCfg(v.cfg.profile, v.cfg.method, v.cfg.Nq, v.cfg.Nlbp, v.cfg.types, v.cfg.rows)
    return Container(boff, pay, ctx, units, cfg, int(v.sym_width), bool(v.ctx_is_prob))


def crc32(data) -> int:
    a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8))
    L = lib()
    L.cabac_crc32.restype = C.c_uint32
    return int(L.cabac_crc32(_ptr(a, u8p), C.c_uint64(a.size)))
