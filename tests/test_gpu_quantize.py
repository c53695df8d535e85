"""GPU: the device quantiser (csrc/quantize.cu through isscabac_b200/quantizer.py) against the numpy
restatement of quantizeWrapper.m / quantize.m (oracle/quantize_oracle.py).

Tolerance (floating point, stated here as the task asks): the device sums the SORTED data with prefix sums,
the restatement sums in storage order like the reference, so centroids are compared at 1e-10 relative
(observed: ~1e-15) and group indices must be identical except for elements closer than 1e-9 to a decision
threshold (observed: none)."""
import json
import os

import numpy as np
import pytest

from oracle import quantize_oracle as Q

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def QZ():
    import isscabac_b200  # noqa: F401
    from isscabac_b200 import quantizer
    assert torch.cuda.is_available()
    return quantizer


def ntf_like(rng, rows, cols):
    """log(theta + eps) of gamma-distributed NTF factors (ISS.m:104-105, trans())"""
    return np.log(rng.gamma(0.6, 1.0, size=(rows, cols)) + 1e-5)


def compare(x, sym, cent, ref_sym, ref_cent, thresholds):
    assert np.allclose(cent, ref_cent, rtol=1e-10, atol=1e-12)
    diff = sym != ref_sym
    if diff.any():     # only elements that sit on a threshold (within rounding) may differ
        d = np.min(np.abs(x[diff][:, None] - np.asarray(thresholds)[None, :]), axis=1)
        assert (d < 1e-9).all()


@pytest.mark.parametrize("mode,dz", [(Q.MODE_LLOYD, 0.7), (Q.MODE_LLOYD, None), (Q.MODE_UNIFORM, 0.7), (Q.MODE_UNIFORM, None),
                                     (Q.MODE_FIXED, 0.7), (Q.MODE_FIXED, None)])
def test_batch_vs_oracle(QZ, mode, dz):
    rng = np.random.default_rng(41)
    shapes = [(400, 20), (109, 20), (1, 1), (3, 2), (37, 5), (400, 20), (109, 20), (64, 128)]   # the last one: 8192 elements
    mats = [ntf_like(rng, r, c) for r, c in shapes]
    mats[2][:] = 1.25                       # a constant matrix: no dead zone, all centroids collapse
    N = 8
    fixed = np.linspace(-9.0, 2.0, N) if mode == Q.MODE_FIXED else None
    cfg = QZ.make_quant_cfg(N=N, GMM=int(mode == Q.MODE_LLOYD), deadzoneQuant=dz, quantileprob=(0.05, 0.95), fixedCentroids=fixed)
    g, c, it = QZ.quantize_matrices(mats, cfg, fixed, want_iters=True)
    torch.cuda.synchronize()
    g, c, it = g.cpu().numpy(), c.cpu().numpy(), it.cpu().numpy()
    pos = 0
    for k, x in enumerate(mats):
        ref_sym, ref_cent, ref_it = Q.quantize_wrapper(x, N=N, mode=mode, deadzone_quant=dz, quantileprob=(0.05, 0.95),
                                                       fixed_centroids=fixed)
        sym = g[pos:pos + x.size].reshape(x.shape, order="F")
        pos += x.size
        thr = [] if dz is None else [Q.matlab_quantile(x, dz)[0]]
        has_dz = dz is not None and (x < thr[0]).any()
        rc = ref_cent[1:] if has_dz else ref_cent
        compare(x, sym, c[k][:len(ref_cent)], ref_sym, ref_cent, thr + list((rc[1:] + rc[:-1]) / 2))
        if mode == Q.MODE_LLOYD:
            assert it[k] == ref_it, (k, it[k], ref_it)


def test_large_matrix_uses_global_scratch(QZ):
    rng = np.random.default_rng(42)
    mats = [ntf_like(rng, 1300, 20), ntf_like(rng, 400, 20)]        # 26,000 elements: H of a one-minute signal
    cfg = QZ.make_quant_cfg(N=8, GMM=1, deadzoneQuant=0.7)
    g, c = QZ.quantize_matrices(mats, cfg)
    g, c = g.cpu().numpy(), c.cpu().numpy()
    pos = 0
    for k, x in enumerate(mats):
        ref_sym, ref_cent, _ = Q.quantize_wrapper(x, N=8, mode=Q.MODE_LLOYD, deadzone_quant=0.7)
        sym = g[pos:pos + x.size].reshape(x.shape, order="F")
        pos += x.size
        rc = ref_cent[1:]
        compare(x, sym, c[k], ref_sym, ref_cent, [Q.matlab_quantile(x, 0.7)[0]] + list((rc[1:] + rc[:-1]) / 2))


def test_quantize_wrapper_mirror_and_coder_round_trip(QZ, tmp_path):
    """quantizeWrapper -> cabacEncode -> cabacDecode (ISS.m:104-110,187-206): the symbols the quantiser
    produces go through the coder and come back."""
    from isscabac_b200 import coder
    rng = np.random.default_rng(43)
    x = ntf_like(rng, 400, 20)
    xq, misc = QZ.quantizeWrapper(x, {"N": 8, "GMM": 1, "deadzoneQuant": 0.7})
    ref_sym, ref_cent, _ = Q.quantize_wrapper(x, N=8, mode=Q.MODE_LLOYD, deadzone_quant=0.7)
    assert (misc["group"] - 1 == ref_sym).all() and np.allclose(misc["centroids"], ref_cent, rtol=1e-10)
    assert np.allclose(xq, ref_cent[ref_sym], rtol=1e-10)
    G = misc["group"] - 1
    param = {"binMethod": "DEC2EG0", "cmTypes": ["cond0", "cond1", "conds0", "conds1"], "Nlbp": 3, "equalProb": 0,
             "fn": str(tmp_path / "W.bin")}
    nbits, ctx_init = coder.cabacEncode(G, 8, param)
    assert nbits > 0
    back = coder.cabacDecode(8, param, ctx_init, G.shape)
    assert (np.asarray(back) == G).all()


def test_handworked_deadzone_cases_on_device(QZ, golden_dir):
    """The fixture worked by hand from quantizeWrapper.m (tests/golden/quantize_handworked.json, arithmetic shown there):
    integer data make every sum exact, so the device result must equal it BIT FOR BIT -- centroids, groups and the
    Lloyd iteration count -- in all three modes with a dead zone.  No tolerance."""
    with open(os.path.join(golden_dir, "quantize_handworked.json")) as f:
        fx = json.load(f)
    x = np.array(fx["x_rows"], dtype=np.float64)
    for c in fx["cases"]:
        fixed = c.get("fixedCentroids")
        cfg = QZ.make_quant_cfg(N=c["N"], GMM=int(c["mode"] == "lloyd"), deadzoneQuant=fx["deadzoneQuant"],
                                quantileprob=tuple(c.get("quantileprob", (0.0, 1.0))), fixedCentroids=fixed)
        g, cent, it = QZ.quantize_matrices([x, x], cfg, fixed, want_iters=True)     # twice: batch entries are independent
        torch.cuda.synchronize()
        for k in range(2):
            assert cent[k].cpu().numpy().tolist() == c["centroids"], c["name"]
            assert g.cpu().numpy()[12 * k: 12 * k + 12].tolist() == c["group_minus_1_flattened"], c["name"]
            if "iterations" in c:
                assert int(it[k].item()) == c["iterations"], c["name"]
