"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's quantiser in front of the coder
(SURVEY.md 8(f) rank 4): ISS/quantizeWrapper.m and ISS/quantize.m, in the reference's own
formulation (masks over the unsorted data, means of the selected elements), double precision.

PARITY UNPINNED by reference outputs: MATLAB / Octave are absent here, the reference holds no vectors
for this step, and MATLAB's `mean`/`sum` summation order is not specified -- centroids can differ from
MATLAB's in the last bits whatever is done.  The restatement is pinned by the hand-computed cases and
the fixed points of Lloyd's iteration in tests/test_quantize_oracle.py, and by review against the cited
lines.  The device path (isscabac_b200/csrc/quantize.cu: sorted data + prefix sums) is compared with
this file at the tolerance stated in tests/test_gpu_quantize.py.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

MODE_UNIFORM, MODE_LLOYD, MODE_FIXED = 0, 1, 2
EPS = float(np.finfo(np.float64).eps)


def matlab_quantile(x: np.ndarray, p) -> np.ndarray:
    """quantile(x, p) of a vector without NaNs as MATLAB's prctile computes it: sample i (1-based, sorted)
    is the (i - 0.5)/n quantile, linear interpolation in between, the extremes outside
    (quantizeWrapper.m:23, quantize.m:61)."""
    xs = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = xs.size
    out = []
    for pp in np.atleast_1d(np.asarray(p, dtype=np.float64)):
        r = pp * n
        k = int(np.floor(r + 0.5))
        kp1 = k + 1
        r = r - k
        k = max(k, 1)
        kp1 = min(kp1, n)
        k = min(k, n)
        out.append((0.5 + r) * xs[kp1 - 1] + (0.5 - r) * xs[k - 1])
    return np.asarray(out)


def linspace(a: float, b: float, n: int) -> np.ndarray:
    """MATLAB linspace: a + (0:n-1)*(b-a)/(n-1) with the last point set to b; n == 1 gives b."""
    if n == 1:
        return np.array([b], dtype=np.float64)
    c = a + np.arange(n, dtype=np.float64) * (b - a) / (n - 1)
    c[-1] = b
    return c


def quantize(x: np.ndarray, centroids=None, N: int = 16, border=(0.0, 1.0)):
    """quantize.m:59-84 (RATE control, QUANTILE borders): returns (xbar, centroids, group[1-based])."""
    x = np.asarray(x, dtype=np.float64).ravel()
    if centroids is None or len(centroids) == 0:
        mn, mx = matlab_quantile(x, border)                 # quantize.m:61-66
        centroids = linspace(mn, mx, N)                     # :75
    centroids = np.sort(np.asarray(centroids, dtype=np.float64))   # :79
    mids = (centroids[:-1] + centroids[1:]) / 2
    # [~,group] = histc(x,[-inf,mids,inf]) (:81): bin k holds edges(k) <= x < edges(k+1)
    group = np.searchsorted(mids, x, side="right") + 1
    return centroids[group - 1], centroids, group


def quantize_lloyd(x: np.ndarray, N: int, init_centroids=None, tol: float = EPS, n_iter: int = 100):
    """quantizeLloyd, quantizeWrapper.m:91-176 (without its own dead-zone option, which the wrapper never uses)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    if init_centroids is None or len(init_centroids) == 0:
        _, centroids, group = quantize(x, None, N, (0.0, 1.0))      # :124
    else:
        _, centroids, group = quantize(x, init_centroids)           # :127
    edges = _edges(x, centroids)                                    # :135
    old = centroids.copy()
    iters = 0
    for _ in range(n_iter):                                         # :141
        iters += 1
        centroids = (edges[:-1] + edges[1:]) / 2                    # :143
        for g in np.unique(group):                                  # :146-151
            centroids[g - 1] = x[group == g].mean()
        edges = _edges(x, centroids)                                # :156
        _, centroids, group = quantize(x, centroids)                # :159-161
        if np.mean((centroids - old) ** 2) < tol:                   # :167
            break
        old = centroids.copy()
    return centroids[group - 1], centroids, group, iters


def _edges(x, c):
    return np.concatenate([[min(x.min(), c.min())], (c[1:] + c[:-1]) / 2, [max(x.max(), c.max())]])


def quantize_wrapper(x: np.ndarray, N: int = 8, mode: int = MODE_LLOYD, deadzone_quant=0.7,
                     quantileprob=(0.0, 1.0), fixed_centroids=None, tol: float = EPS, n_iter: int = 100):
    """quantizeWrapper.m:1-88.  Returns (group - 1 as the coder's symbols, shaped like x; centroids; iterations)."""
    shape = np.shape(x)
    x = np.asarray(x, dtype=np.float64).ravel(order="F")
    fixed = None if fixed_centroids is None else np.asarray(fixed_centroids, dtype=np.float64).copy()
    maskdz = np.zeros(x.size, dtype=bool)
    had_dz = False
    xin = x
    if deadzone_quant is not None:                                   # :22-36
        thr = matlab_quantile(x, deadzone_quant)[0]
        maskdz = x < thr
        if maskdz.sum() != 0:
            had_dz = True
            small = x[maskdz]
            x = x[~maskdz]
            N = N - 1
            if fixed is not None:
                fixed = fixed[1:]
    iters = 0
    if mode == MODE_FIXED:                                           # :39-40
        _, centroids, group = quantize(x, fixed)
    elif mode == MODE_LLOYD:                                         # :42-43
        _, centroids, group, iters = quantize_lloyd(x, N, None, tol, n_iter)
    else:                                                            # :46-47
        _, centroids, group = quantize(x, None, N, quantileprob)
    if had_dz:                                                       # :54-72
        centroids = np.concatenate([[small.mean()], centroids])
        g = np.zeros(xin.size, dtype=np.int64)
        g[~maskdz] = group + 1
        g[maskdz] = 1
        group = g
    return (group - 1).reshape(shape, order="F").astype(np.uint8), centroids, iters
