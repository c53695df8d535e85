"""GPU: the multi-GPU entry points of the C ABI (include/isscabac.h: cabac_multi_gpu_*) with a communicator of ONE rank --
ncclCommInitRank, the length exchange + device scan, both payload assemblies (grouped NCCL broadcasts / compaction fused
with the stores into the symmetric buffer) and the barrier all run their real code paths on a single GPU; the result must be
the oracle's container.  The N = 2 and N = 8 runs of the same checks are tests/gpu_tools/nccl_check.py (profiles/r2_nccl_check_n*.jsonl)."""
import numpy as np
import pytest

import oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _job(seed=13, n_streams=1500, n_ctx=9):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 700, size=n_streams)
    off = np.zeros(n_streams + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    code = rng.integers(0, n_ctx, size=n).astype(np.uint8)
    code[rng.random(n) < 0.25] = O.OP8_EP
    ops = ((code << 1) | (rng.random(n) < 0.3)).astype(np.uint8)
    ci = rng.integers(0, 126, size=n_ctx).astype(np.uint8)
    return ops, off, ci


@pytest.mark.parametrize("fused", [False, True])
def test_single_rank_communicator(fused):
    import isscabac_b200 as I
    from isscabac_b200 import multi_gpu as MG
    assert torch.cuda.is_available()
    ops, off, ci = _job()
    n = len(off) - 1
    first = np.array([0, n], dtype=np.uint32)
    pay, table, full = MG.encode_ops_sharded(ops, off, ci, assemble=True, slab_stride=512, first=first, fused_p2p=fused)
    torch.cuda.synchronize()
    slab, lens = O.encode_ops(ops, off.astype(np.uint64), ci, out_stride=512, n_threads=4)
    payload, boff = O.compact(slab, lens)
    assert (table.lengths.cpu().numpy().astype(np.uint32) == lens).all()
    assert (table.byte_off.cpu().numpy().astype(np.uint64) == boff).all()
    assert table.stream_counts == [n] and table.rank_bytes == [len(payload)] and table.rank_base == [0]
    assert (full.cpu().numpy() == payload).all()
    bins, ok = MG.decode_ops_sharded(full, table, ops, off, ci)
    assert bool(ok.all().item()) and (bins.cpu().numpy() == (ops & 1)).all()
    # the raw entry points: info, barrier, argument checks
    mg = MG.default_handle()
    assert (mg.rank, mg.world) == (0, 1)
    mg.barrier()
    with pytest.raises(I.CabacError):
        mg.gather_table(np.array([1, n], dtype=np.uint32), table.lengths)      # a partition must start at stream 0


def test_stride_overflow_is_retried_not_ignored():
    """encode_ops_sharded used the heuristic stride and never looked at the overflow flag.  Contexts parked at
    state 62, each hit once by an LPS, grow 6 bits per bin -- past the 2 bits per op of the heuristic stride."""
    from isscabac_b200 import multi_gpu as MG
    n_streams, n_ops, n_ctx = 64, 990, 999                                    # every op on a fresh context (u16 op format)
    ops = np.tile(((np.arange(n_ops) << 1) | 0).astype(np.uint16), n_streams)  # bin 0 = LPS for mps = 1
    off = (np.arange(n_streams + 1) * n_ops).astype(np.int64)
    ci = np.full(n_ctx, (62 << 1) | 1, dtype=np.uint8)
    pay, table, _ = MG.encode_ops_sharded(ops, off, ci)                       # default (heuristic) stride
    torch.cuda.synchronize()
    slab, lens = O.encode_ops(ops, off.astype(np.uint64), ci, out_stride=4096, n_threads=4)
    payload, boff = O.compact(slab, lens)
    assert int(lens.max()) > n_ops // 4 + 64                                  # the heuristic stride really is too small
    assert (table.byte_off.cpu().numpy().astype(np.uint64) == boff).all()
    assert (pay.payload.cpu().numpy()[:len(payload)] == payload).all()
