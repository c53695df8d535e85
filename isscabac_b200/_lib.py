"""ctypes binding of libisscabac.so (the C ABI declared in include/isscabac.h).

There is no fallback of any kind: if the shared library is missing it is built with
nvcc, and if that fails or no CUDA device is present the compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class CabacError(RuntimeError):
    def __init__(self, code: int, detail: str):
        self.code = code
        super().__init__(f"isscabac error {code}: {detail}")


class SymCfg(C.Structure):
    """isscabac_symcfg"""
    _fields_ = [("profile", C.c_int32), ("method", C.c_int32), ("Nq", C.c_uint32),
                ("Nlbp", C.c_int32), ("types", C.c_uint32), ("rows", C.c_uint32)]


class MxArg(C.Structure):
    """isscabac_mxarg"""
    _fields_ = [("is_char", C.c_int32), ("s", C.c_char_p), ("d", f64p), ("m", C.c_int32), ("n", C.c_int32)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.environ.get("ISSCABAC_LIB")      # a tuning variant built by build.build_variant()
        if path:
            if not os.path.exists(path):
                raise FileNotFoundError(f"ISSCABAC_LIB={path} does not exist")
        else:
            path = _build.LIB
            if _build.needs_build():
                path = _build.build()
        L = C.CDLL(path)
        L.isscabac_strerror.restype = C.c_char_p
        L.isscabac_last_error.restype = C.c_char_p
        L.cabac_slab_stride_bound.restype = C.c_uint64
        L.cabac_slab_stride_bound.argtypes = [C.c_uint64]
        L.cabac_compact_scratch_bytes.restype = C.c_size_t
        L.cabac_compact_scratch_bytes.argtypes = [C.c_uint32]
        if hasattr(L, "cabac_encode_ops_kernel"):
            L.cabac_encode_ops_kernel.restype = C.c_char_p
            L.cabac_encode_ops_kernel.argtypes = [C.c_uint32, C.c_uint32]
        if hasattr(L, "cabac_decode_ops_kernel"):
            L.cabac_decode_ops_kernel.restype = C.c_char_p
            L.cabac_decode_ops_kernel.argtypes = [C.c_uint32, C.c_uint32]
        if hasattr(L, "cabac_binarize_scratch_bytes"):
            L.cabac_binarize_scratch_bytes.restype = C.c_size_t
            L.cabac_binarize_scratch_bytes.argtypes = [C.c_uint64, C.c_uint32]
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    if rc != 0:
        L = lib()
        detail = L.isscabac_last_error().decode(errors="replace") or L.isscabac_strerror(rc).decode()
        raise CabacError(rc, detail)


def vp(x) -> C.c_void_p:
    """torch tensor / numpy array / int / None -> void*"""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)
