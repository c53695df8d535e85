"""N-GPU check of the sharded path over NCCL (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tests/gpu_tools/nccl_check.py

Every rank encodes its contiguous range of streams on its own GPU (isscabac_b200.multi_gpu),
the ranks all-gather the lengths, assemble ONE payload, and every rank decodes its own streams
back out of the assembled payload.  Rank 0 compares the assembled container with the oracle's
single-process result (byte-identical payload and offset table).  Equal-count and work-balanced
partitions.  Prints one JSON line per partitioning on rank 0.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import isscabac_b200 as I  # noqa: E402
from isscabac_b200 import multi_gpu as MG  # noqa: E402


def make_job(seed=11, n_streams=5003, max_len=3000, n_ctx=23):
    rng = np.random.default_rng(seed)
    lens = np.minimum(np.round(rng.lognormal(np.log(300), 1.0, size=n_streams)), max_len).astype(np.int64)
    off = np.zeros(n_streams + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    code = rng.integers(0, n_ctx, size=n).astype(np.uint8)
    code[rng.random(n) < 0.25] = I.OP8_EP
    ops = ((code << 1) | (rng.random(n) < 0.3)).astype(np.uint8)
    ci = rng.integers(0, 126, size=n_ctx).astype(np.uint8)
    return ops, off, ci


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops, off, ci = make_job()
    n = len(off) - 1
    for balanced, fused in ((False, False), (True, False), (False, True), (True, True)):
        parts = MG.balanced_ranges(off, world) if balanced else [MG.shard_range(n, r, world) for r in range(world)]
        a, b = parts[rank]
        first = np.array([p[0] for p in parts] + [n], dtype=np.uint32)     # the partition is known to every rank
        lo, hi = int(off[a]), int(off[b])
        loc_off = off[a:b + 1] - off[a]
        # through the C ABI (cabac_multi_gpu_*): NCCL all-gather-v of the lengths + device scan, then either grouped NCCL
        # broadcasts of the payloads or the compaction kernel storing straight into every rank's buffer (fused)
        pay, table, full = MG.encode_ops_sharded(ops[lo:hi], loc_off, ci, assemble=True, slab_stride=1024,
                                                 first=first if balanced else None, fused_p2p=fused)
        bins, ok = MG.decode_ops_sharded(full, table, ops[lo:hi], loc_off, ci)
        torch.cuda.synchronize()
        good = bool(ok.all().item()) and bool((bins.cpu().numpy() == (ops[lo:hi] & 1)).all())
        flag = torch.tensor([1 if good else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            import oracle as O
            slab, lens = O.encode_ops(ops, off.astype(np.uint64), ci, out_stride=1024)
            payload, boff = O.compact(slab, lens)
            same = bool((full.cpu().numpy() == payload).all()) and bool(
                (table.byte_off.cpu().numpy().astype(np.uint64) == boff).all())
            print(json.dumps({"check": "nccl_sharded_container", "world": world, "balanced": balanced, "assembly": "compaction fused with peer stores (cabac_multi_gpu_compact_p2p)" if fused else "grouped ncclBroadcast (cabac_multi_gpu_assemble)",
                              "streams": n, "payload_bytes": int(len(payload)), "stream_counts": table.stream_counts,
                              "round_trip_all_ranks": bool(flag.item()), "byte_identical_to_oracle": same}))
            assert same and bool(flag.item())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
