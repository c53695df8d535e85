"""Python mirrors of the reference's MATLAB layer, same names and argument meaning, so the
parity tests read like the reference's own scripts:

  SimpleCABACMex(cmd, ...)   CABAC/SimpleCABACMex.cpp:100-472 command protocol (via simplecabac_dispatch)
  cabacWrapper               CABAC/cabacWrapper.m:15-77
  cabacBinarizer / cabacDebinarizer / cabacDecodeSymbolFinished
                             CABAC/cabacBinarizer.m, cabacDebinarizer.m, cabacDecodeSymbolFinished.m

Everything computes on the GPU through libisscabac.so (the binarizer helpers launch the
device binarizer on a one-symbol batch); MATLAB's 1-based context ids stay 1-based at the
coder.py level and 0-based here, exactly as in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import engine as E
from ._lib import MxArg, f64p, lib


class MexError(RuntimeError):
    """mexErrMsgTxt: the MEX call was aborted with this message."""


def SimpleCABACMex(cmd, *args, nargout=None):
    """One mexFunction call.  Numeric arguments are converted to MATLAB doubles (column-major).
    nargout defaults to 1 for the commands that return something."""
    if nargout is None:
        nargout = 1 if cmd in ("initByProb", "initByState", "getNumBits", "decodeBin") else 0
    allargs = (cmd,) + args
    arr = (MxArg * max(len(allargs), 1))()
    keep = []
    for i, a in enumerate(allargs):
        if isinstance(a, str):
            b = a.encode()
            keep.append(b)
            arr[i] = MxArg(1, b, None, 1, len(b))
        else:
            d = np.asarray(a, dtype=np.float64)
            shape = d.shape if d.ndim == 2 else (1, d.size)
            d = np.ascontiguousarray(d.reshape(shape).reshape(-1, order="F"))
            keep.append(d)
            arr[i] = MxArg(0, None, d.ctypes.data_as(f64p), shape[0], shape[1])
    out = (C.c_double * 4)()
    out_n = C.c_int(0)
    err = C.create_string_buffer(512)
    rc = lib().simplecabac_dispatch(int(nargout), out, 4, C.byref(out_n), len(allargs), arr, err, 512)
    if rc != 0:
        raise MexError(err.value.decode(errors="replace"))
    if out_n.value:
        return out[0]
    return None


def SimpleCABACMexStats(cmd, handle, ctxID, nargout=2):
    """[trace, stats] = SimpleCABACMex('getEncoderStats' | 'getDecoderStats', handle, ctxID)
    (SimpleCABACMex.cpp:356-466, Windows builds of the reference).  -> (trace uint8 [5, M] as MATLAB
    sees it, stats uint32 [128, 128] in MATLAB's column-major view: stats[a, p] = transitions p -> a)."""
    allargs = (cmd, handle, ctxID)
    arr = (MxArg * 3)()
    keep = []
    for i, a in enumerate(allargs):
        if isinstance(a, str):
            b = a.encode()
            keep.append(b)
            arr[i] = MxArg(1, b, None, 1, len(b))
        else:
            d = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
            keep.append(d)
            arr[i] = MxArg(0, None, d.ctypes.data_as(f64p), 1, d.size)
    err = C.create_string_buffer(512)
    n = C.c_uint64(0)
    trans = np.zeros(128 * 128, dtype=np.uint32)
    L = lib()
    rc = L.simplecabac_dispatch_stats(int(nargout), 3, arr, None, C.c_uint64(0), C.byref(n), None, err, 512)
    if rc != 0:
        raise MexError(err.value.decode(errors="replace"))
    steps = np.zeros(max(int(n.value), 1) * 5, dtype=np.uint8)
    rc = L.simplecabac_dispatch_stats(int(nargout), 3, arr, steps.ctypes.data_as(C.c_void_p), C.c_uint64(int(n.value)),
                                      C.byref(n), trans.ctypes.data_as(C.c_void_p), err, 512)
    if rc != 0:
        raise MexError(err.value.decode(errors="replace"))
    m = int(n.value)
    # memory order of the reference: trace[entry*5 + k] in a 5 x M matrix, stats[p*128 + a] in a 128 x 128 one
    return steps[:5 * m].reshape(m, 5).T.copy(), trans.reshape(128, 128).T.copy()


class cabacWrapper:
    """CABAC/cabacWrapper.m -- thin 1:1 forwarding to the MEX commands."""

    def __init__(self, cm, fn, initByProb=1):
        if not isinstance(fn, str):
            raise ValueError("bitStreamName should be string")
        self.bitStreamName = fn
        self.contextModelInitOptions = np.asarray(cm, dtype=np.float64)
        self.cabac_handle = None
        self.init(initByProb)

    def init(self, initByProb=1):
        cmd = "initByProb" if initByProb else "initByState"
        self.cabac_handle = SimpleCABACMex(cmd, self.bitStreamName, self.contextModelInitOptions)

    def encodeStart(self):
        SimpleCABACMex("encodeStart", self.cabac_handle)

    def encodeBin(self, binValue, ctxID):
        if binValue == 1 or binValue == 0:
            SimpleCABACMex("encodeBin", self.cabac_handle, binValue, ctxID)
        else:
            raise ValueError("bin value to be encoded should either be 1 or 0")

    def encodeFinish(self):
        SimpleCABACMex("encodeFinish", self.cabac_handle)

    def decodeStart(self):
        SimpleCABACMex("decodeStart", self.cabac_handle)

    def decodeBin(self, ctxID):
        return int(SimpleCABACMex("decodeBin", self.cabac_handle, ctxID))

    def decodeFinish(self):
        SimpleCABACMex("decodeFinish", self.cabac_handle)

    def getNumBits(self):
        return int(SimpleCABACMex("getNumBits", self.cabac_handle))

    def setTrace(self, on=1):
        """Not in the reference: its Windows builds always trace, its Linux builds never do."""
        SimpleCABACMex("setTrace", self.cabac_handle, on)

    def getEncoderStats(self, ctxID):
        """cabacWrapper.m:69-71 -> (trace, stats)"""
        return SimpleCABACMexStats("getEncoderStats", self.cabac_handle, ctxID)

    def getDecoderStats(self, ctxID):
        """cabacWrapper.m:72-74 -> (trace, stats)"""
        return SimpleCABACMexStats("getDecoderStats", self.cabac_handle, ctxID)

    def close(self):
        """Not in the reference (its instances leak by design): releases the GPU-side state."""
        if self.cabac_handle is not None:
            SimpleCABACMex("destroy", self.cabac_handle)
            self.cabac_handle = None


def _method(method):
    return E.METHODS[method] if isinstance(method, str) else int(method)


def cabacBinarizer(v, Nq, method):
    """cabacBinarizer.m: integer v -> bin string (list of 0/1), computed by the device binarizer."""
    cfg = E.make_cfg(E.PROFILE_FLAT, _method(method), Nq, 3, 0, 0)
    ops, _ = E.binarize_symbols(cfg, np.array([int(v)], dtype=np.uint32), np.array([0, 1], dtype=np.int64))
    return [int(x) & 1 for x in ops.cpu().numpy()]


def cabacDebinarizer(c, Nq, method):
    """cabacDebinarizer.m: bin string -> integer (closed forms of rTUCode / rEGCode / rFLCode)."""
    m = _method(method)
    c = [int(x) for x in c]
    first0 = c.index(0) + 1 if 0 in c else 0
    if m == E.BIN_TU:
        return first0 - 1 if first0 else Nq - 1
    if m == E.BIN_FL32:
        return int("".join(map(str, c)), 2)
    k = m - E.BIN_EG0
    xi = c[first0:]
    return (1 << k) * ((1 << (first0 - 1)) - 1) + (int("".join(map(str, xi)), 2) if xi else 0)


def cabacDecodeSymbolFinished(g, n, Nq, binMethod, n_p, n_s):
    """cabacDecodeSymbolFinished.m:10-32 -> (isFinished, n, n_p, n_s); n is 1-based."""
    m = _method(binMethod)
    fin = False
    if m == E.BIN_TU:
        fin = g[n - 1] == 0 or n == Nq - 1
    elif m in (E.BIN_EG0, E.BIN_EG1, E.BIN_EG2):
        k = m - E.BIN_EG0
        if n_s == -1:
            if g[n - 1] == 0:
                n_p = n_p + n
                n_s = k + n_p - 1
                if n_s == 0:
                    fin = True
        elif n_s == 1:
            fin = True
        else:
            n_s -= 1
    return fin, n, n_p, n_s
