"""Mirror of the reference's quantiser in front of the coder (ISS/quantizeWrapper.m:1-88, ISS/quantize.m)
over a batch of matrices on the device: `quantizeWrapper(x, qParam)` -> (group indices - 1 = the coder's
symbols, centroids).  See csrc/quantize.cu.  No CPU fallback: without a CUDA device the call fails."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, vp
from .engine import _require_cuda, _stream_ptr

QUANT_UNIFORM, QUANT_LLOYD, QUANT_FIXED = 0, 1, 2


class QuantCfg(C.Structure):
    """isscabac_quantcfg"""
    _fields_ = [("N", C.c_int32), ("mode", C.c_int32), ("deadzone_quant", C.c_double), ("q_lo", C.c_double),
                ("q_hi", C.c_double), ("tol", C.c_double), ("max_iter", C.c_int32), ("reserved", C.c_int32)]


def make_quant_cfg(N=8, GMM=1, deadzoneQuant=0.7, quantileprob=(0.0, 1.0), fixedCentroids=None,
                   tol=float(np.finfo(np.float64).eps), nIter=100) -> QuantCfg:
    """qParam of quantizeWrapper.m with ISS.m:42-46's defaults (N = 8, GMM = 1, deadzoneQuant = 0.7)."""
    mode = QUANT_FIXED if fixedCentroids is not None else (QUANT_LLOYD if GMM else QUANT_UNIFORM)
    dz = -1.0 if deadzoneQuant is None else float(deadzoneQuant)
    return QuantCfg(int(N), mode, dz, float(quantileprob[0]), float(quantileprob[1]), float(tol), int(nIter), 0)


def quantize_matrices(xs, cfg: QuantCfg, fixed_centroids=None, want_iters: bool = False):
    """xs: list of matrices (numpy or torch, any shape; flattened column-major like x(:)) or one
    concatenated 1-D double tensor with `elem_off` given as a (tensor, offsets) pair.
    -> (groups u8 [total elements], centroids f64 [n_matrices, N] (, iterations))."""
    dev = _require_cuda()
    if isinstance(xs, tuple):
        x_t, off = xs
        x_t = torch.as_tensor(x_t, dtype=torch.float64, device=dev).contiguous()
        off_t = torch.as_tensor(off, dtype=torch.int64, device=dev).contiguous()
        sizes = (off_t[1:] - off_t[:-1])
        max_elems = int(sizes.max().item()) if sizes.numel() else 0
    else:
        flat = [np.asarray(m.detach().cpu().numpy() if hasattr(m, "detach") else m, dtype=np.float64).ravel(order="F") for m in xs]
        off = np.zeros(len(flat) + 1, dtype=np.int64)
        np.cumsum([f.size for f in flat], out=off[1:])
        x_t = torch.as_tensor(np.concatenate(flat) if flat else np.zeros(0), dtype=torch.float64, device=dev)
        off_t = torch.as_tensor(off, device=dev)
        max_elems = int(max([f.size for f in flat], default=0))
    n_mat = off_t.numel() - 1
    L = lib()
    L.cabac_quantize_scratch_bytes.restype = C.c_size_t
    L.cabac_quantize_scratch_bytes.argtypes = [C.c_uint32, C.c_uint64]
    scratch = torch.empty(int(L.cabac_quantize_scratch_bytes(n_mat, max_elems)), dtype=torch.uint8, device=dev)
    groups = torch.empty(max(x_t.numel(), 1), dtype=torch.uint8, device=dev)
    cent = torch.empty((max(n_mat, 1), cfg.N), dtype=torch.float64, device=dev)
    iters = torch.zeros(max(n_mat, 1), dtype=torch.int32, device=dev) if want_iters else None
    fixed_t = None
    if fixed_centroids is not None:
        fixed_t = torch.as_tensor(np.asarray(fixed_centroids, dtype=np.float64), device=dev).contiguous()
        if fixed_t.numel() != cfg.N:
            raise ValueError("fixed_centroids must hold N values")
    check(L.cabac_quantize_matrices(C.byref(cfg), C.c_uint32(n_mat), vp(off_t), C.c_uint64(max_elems), vp(x_t), vp(fixed_t),
                                    vp(groups), vp(cent), vp(iters), vp(scratch), _stream_ptr()))
    out = (groups[:x_t.numel()], cent[:n_mat])
    return out + (iters[:n_mat],) if want_iters else out


def quantizeWrapper(x, qParam: dict | None = None):
    """[xq, misc] = quantizeWrapper(x, qParam) for one matrix: returns (xq, misc) with misc = {centroids, group
    (1-based like MATLAB), delta} (quantizeWrapper.m:74-87)."""
    q = dict(qParam or {})
    cfg = make_quant_cfg(q.get("N", 8), q.get("GMM", 1), q.get("deadzoneQuant", None), q.get("quantileprob", (0.0, 1.0)),
                         q.get("fixedCentroids", None))
    x = np.asarray(x, dtype=np.float64)
    g, c = quantize_matrices([x], cfg, q.get("fixedCentroids", None))
    group = g.cpu().numpy().astype(np.int64).reshape(x.shape, order="F") + 1
    cent = c[0].cpu().numpy()
    return cent[group - 1], {"centroids": cent, "group": group, "delta": np.diff(cent)}
