"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path (isscabac_b200/multi_gpu.py) --
uneven shards, global offset table, payload assembly -- with the per-rank bitstreams produced by
the oracle standing in for the GPU coder.  The assembled container must equal what one process
produces for all streams."""
import os
import socket

import numpy as np
import pytest

import oracle as O

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_job(seed=5, n_streams=203):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 400, size=n_streams)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    code = rng.integers(0, 7, size=n).astype(np.uint8)
    code[rng.random(n) < 0.2] = O.OP8_EP
    ops = ((code << 1) | (rng.random(n) < 0.3)).astype(np.uint8)
    ci = rng.integers(0, 126, size=7).astype(np.uint8)
    return ops, off, ci


def _worker(rank, world, port, balanced, q):
    import torch.distributed as dist
    from isscabac_b200 import multi_gpu as MG
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops, off, ci = _make_job()
        n = len(off) - 1
        a, b = MG.balanced_ranges(off, world, granule=8)[rank] if balanced else MG.shard_range(n, rank, world)
        lo, hi = int(off[a]), int(off[b])
        loc_off = (off[a:b + 1] - off[a]).astype(np.uint64)
        slab, lens = O.encode_ops(ops[lo:hi], loc_off, ci, out_stride=256, n_threads=1)   # stands in for the GPU coder
        payload, _ = O.compact(slab, lens)
        table = MG.gather_table(torch.from_numpy(lens.astype(np.int64)).to(torch.int32))
        full = MG.assemble_payload(torch.from_numpy(payload.copy()), table)
        sa, sb, b0, b1 = MG.local_slice(table, rank)
        assert (sa, sb) == (a, b) and b1 - b0 == len(payload)
        # this rank decodes its own streams out of the replicated payload
        boff = (table.byte_off[sa:sb + 1] - b0).numpy().astype(np.uint64)
        bins, ok = O.decode_ops(full[b0:b1].numpy(), boff, ops[lo:hi], loc_off, ci, n_threads=1)
        assert ok.all() and (bins == (ops[lo:hi] & 1)).all()
        if rank == 0:
            q.put((table.byte_off.numpy().copy(), full.numpy().copy(), table.stream_counts, table.rank_bytes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balanced", [False, True])
def test_two_rank_container_equals_single_process(balanced):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, balanced, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    byte_off, full, counts, rank_bytes = q.get(timeout=120)
    for p_ in procs:
        p_.join(timeout=120)
        assert p_.exitcode == 0
    ops, off, ci = _make_job()
    slab, lens = O.encode_ops(ops, off, ci, out_stride=256, n_threads=2)
    payload, boff = O.compact(slab, lens)
    assert sum(counts) == len(off) - 1 and sum(rank_bytes) == len(payload)
    assert (byte_off.astype(np.uint64) == boff).all()
    assert (full == payload).all()


def test_partitions():
    from isscabac_b200 import multi_gpu as MG
    assert [MG.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [MG.shard_range(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    # skewed work: 1000 light streams then 24 heavy ones
    w = np.concatenate([np.full(1000, 10), np.full(24, 10000)])
    off = np.concatenate([[0], np.cumsum(w)])
    r = MG.balanced_ranges(off, 4, granule=8)
    assert r[0][0] == 0 and r[-1][1] == 1024 and all(r[i][1] == r[i + 1][0] for i in range(3))
    work = [int(off[b] - off[a]) for a, b in r]
    assert max(work) <= 1.5 * (sum(work) / 4)
