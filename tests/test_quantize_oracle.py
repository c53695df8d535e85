"""CPU: the numpy restatement of the reference quantiser (oracle/quantize_oracle.py; quantizeWrapper.m,
quantize.m) against hand-computed cases and the properties Lloyd's iteration must have.  The reference
holds no vectors for this step and MATLAB is absent: parity with MATLAB output is unpinned (see the
oracle's header); these tests pin the restatement to the cited lines."""
import json
import os

import numpy as np

from oracle import quantize_oracle as Q


def test_matlab_quantile_definition():
    x = np.array([4.0, 1.0, 3.0, 2.0])            # sorted 1 2 3 4: sample i is the (i - 0.5)/4 quantile
    assert Q.matlab_quantile(x, [0.0, 0.125, 0.25, 0.375, 0.5, 0.875, 1.0]).tolist() == [1.0, 1.0, 1.5, 2.0, 2.5, 4.0, 4.0]
    assert Q.matlab_quantile(np.arange(1.0, 11.0), 0.7)[0] == 7.5


def test_quantize_uniform_and_fixed():
    x = np.array([0.0, 0.24, 0.25, 0.5, 0.74, 0.76, 1.0])
    xb, c, g = Q.quantize(x, None, 3, (0.0, 1.0))   # centroids 0, .5, 1; edges -inf .25 .75 inf; edges(k) <= x < edges(k+1)
    assert c.tolist() == [0.0, 0.5, 1.0]
    assert g.tolist() == [1, 1, 2, 2, 2, 3, 3]
    assert xb.tolist() == [0.0, 0.0, 0.5, 0.5, 0.5, 1.0, 1.0]
    _, c2, g2 = Q.quantize(x, [1.0, 0.0], 0)        # given centroids are sorted first (quantize.m:79)
    assert c2.tolist() == [0.0, 1.0] and g2.tolist() == [1, 1, 1, 2, 2, 2, 2]


def test_lloyd_two_clusters_by_hand():
    # start: linspace(0, 11, 2) = [0, 11], edge 5.5 -> groups {0,1,2} / {9,10,11}; means 1 and 10; next edge 5.5:
    # nothing moves, the second iteration finds the same centroids and stops
    x = np.array([0.0, 1.0, 2.0, 9.0, 10.0, 11.0])
    xb, c, g, it = Q.quantize_lloyd(x, 2)
    assert c.tolist() == [1.0, 10.0] and g.tolist() == [1, 1, 1, 2, 2, 2] and it == 2
    assert xb.tolist() == [1.0, 1.0, 1.0, 10.0, 10.0, 10.0]


def test_lloyd_is_a_fixed_point_and_lowers_the_error():
    rng = np.random.default_rng(5)
    x = np.log(rng.gamma(0.6, 1.0, size=4000) + 1e-5)
    xb0, c0, g0 = Q.quantize(x, None, 7, (0.0, 1.0))
    xb, c, g, it = Q.quantize_lloyd(x, 7)
    assert np.mean((x - xb) ** 2) < np.mean((x - xb0) ** 2)
    assert (np.diff(c) > 0).all() and 1 <= it <= 100
    # centroid condition: exact at a fixed point; the reference stops after nIter = 100 (quantizeWrapper.m:103)
    # wherever it is, so the returned centroids are the means of the grouping one step earlier
    for k in np.unique(g):
        assert abs(c[k - 1] - x[g == k].mean()) < (1e-9 if it < 100 else 1e-2)
    _, c_long, g_long, it_long = Q.quantize_lloyd(x, 7, n_iter=5000)
    assert it_long < 5000
    for k in np.unique(g_long):
        assert abs(c_long[k - 1] - x[g_long == k].mean()) < 1e-7
    mids = (c[1:] + c[:-1]) / 2                      # nearest-neighbour condition
    assert (g == np.searchsorted(mids, x, side="right") + 1).all()


def test_wrapper_dead_zone():
    rng = np.random.default_rng(6)
    x = np.log(rng.gamma(0.6, 1.0, size=(400, 20)) + 1e-5)
    sym, cent, it = Q.quantize_wrapper(x, N=8, mode=Q.MODE_LLOYD, deadzone_quant=0.7)
    assert sym.shape == x.shape and sym.dtype == np.uint8 and sym.max() == 7 and len(cent) == 8
    thr = Q.matlab_quantile(x, 0.7)[0]
    assert ((sym == 0) == (x < thr)).all()           # quantizeWrapper.m:23-24,66
    assert abs((sym == 0).mean() - 0.7) < 1e-3       # ISS.m:44: the lowest 70 % share one symbol
    assert cent[0] == x[x < thr].mean()              # :57
    # without a dead zone all N centroids come from Lloyd's iteration
    sym2, cent2, _ = Q.quantize_wrapper(x, N=8, mode=Q.MODE_LLOYD, deadzone_quant=None)
    assert len(cent2) == 8 and sym2.max() == 7
    # fixed centroids: the first one is dropped for the dead zone and replaced by its mean (:33,57)
    fc = np.linspace(x.min(), x.max(), 8)
    sym3, cent3, _ = Q.quantize_wrapper(x, N=8, mode=Q.MODE_FIXED, deadzone_quant=0.7, fixed_centroids=fc)
    assert (cent3[1:] == fc[1:]).all() and cent3[0] == cent[0]
    # uniform between quantiles
    sym4, cent4, _ = Q.quantize_wrapper(x, N=8, mode=Q.MODE_UNIFORM, deadzone_quant=None, quantileprob=(0.1, 0.9))
    lo, hi = Q.matlab_quantile(x, [0.1, 0.9])
    assert cent4[0] == lo and cent4[-1] == hi and np.allclose(np.diff(cent4), (hi - lo) / 7)


def test_handworked_deadzone_cases(golden_dir):
    """tests/golden/quantize_handworked.json: a 12-element matrix taken through quantizeWrapper.m BY HAND (the file shows the
    arithmetic line by line; integer data, so every operation is exact or a single correctly rounded division) in all
    three modes with a dead zone.  The restatement must reproduce it to the bit."""
    with open(os.path.join(golden_dir, "quantize_handworked.json")) as f:
        fx = json.load(f)
    x = np.array(fx["x_rows"], dtype=np.float64)
    assert x.ravel(order="F").tolist() == fx["x_flattened_column_major"]
    assert Q.matlab_quantile(x, fx["deadzoneQuant"])[0] == -1.5
    for c in fx["cases"]:
        mode = {"lloyd": Q.MODE_LLOYD, "uniform": Q.MODE_UNIFORM, "fixed": Q.MODE_FIXED}[c["mode"]]
        sym, cent, it = Q.quantize_wrapper(x, N=c["N"], mode=mode, deadzone_quant=fx["deadzoneQuant"],
                                           quantileprob=tuple(c.get("quantileprob", (0.0, 1.0))), fixed_centroids=c.get("fixedCentroids"))
        assert cent.tolist() == c["centroids"], c["name"]
        assert sym.ravel(order="F").tolist() == c["group_minus_1_flattened"], c["name"]
        if "iterations" in c:
            assert it == c["iterations"], c["name"]
