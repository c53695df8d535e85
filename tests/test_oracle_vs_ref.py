"""CPU, build container only: the C restatement against the UNMODIFIED reference engine
(oracle/_ref) on a few million random bins -- enough to hit the rare carry branches
(ripple carries ~2.5 per M bins, finish() carry branch 0.6% of streams)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (no /root/reference here)")


def _streams(seed, n_streams, n_ops, n_ctx, p_ep):
    rng = np.random.default_rng(seed)
    n = n_streams * n_ops
    ctx = rng.integers(0, n_ctx, size=n)
    p1 = 0.20 + 0.10 * (np.arange(n_ctx) % 5)
    bins = (rng.random(n) < p1[ctx]).astype(np.uint8)
    code = ctx.astype(np.uint8)
    ep = rng.random(n) < p_ep
    code[ep] = O.OP8_EP
    bins[ep] = rng.integers(0, 2, size=int(ep.sum()))
    ops = ((code << 1) | bins).astype(np.uint8)
    off = (np.arange(n_streams + 1) * n_ops).astype(np.uint64)
    return ops, off


@pytest.mark.parametrize("seed,n_streams,n_ops,n_ctx,p_ep", [
    (1, 20000, 64, 4, 0.25),      # the K7/K8 regime: short streams, many finish() carries
    (2, 512, 4096, 23, 0.25),     # config-3 mix
    (3, 256, 4096, 3, 0.0),       # context-only, skewed
    (4, 256, 4096, 1, 1.0),       # bypass-only
])
def test_oracle_equals_reference(seed, n_streams, n_ops, n_ctx, p_ep):
    ops, off = _streams(seed, n_streams, n_ops, n_ctx, p_ep)
    ci = np.full(n_ctx, 1, dtype=np.uint8)
    s_ref, l_ref = O.encode_ops(ops, off, ci, impl="ref", n_threads=8)
    s_orc, l_orc = O.encode_ops(ops, off, ci, impl="oracle", n_threads=8)
    assert (l_ref == l_orc).all()
    assert (s_ref == s_orc).all()
    payload, boff = O.compact(s_ref, l_ref)
    b_ref, ok_ref = O.decode_ops(payload, boff, ops, off, ci, impl="ref", n_threads=8)
    b_orc, ok_orc = O.decode_ops(payload, boff, ops, off, ci, impl="oracle", n_threads=8)
    assert ok_ref.all() and ok_orc.all()
    assert (b_ref == b_orc).all() and (b_orc == (ops & 1)).all()


def test_prob_mapping_dense():
    p = np.linspace(0.0, 1.0, 20001)
    assert (O.ctx_from_p0(p) == O.ref_prob_to_state(p)).all()
