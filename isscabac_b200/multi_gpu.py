"""Stream sharding over the GPUs of one box (one process per GPU, torch.distributed).

Streams are independent (separate start()...finish() lifetimes, separate context sets), so the
coding loop never communicates: every rank codes a contiguous range of stream ids with the
single-GPU entry points of engine.py.  What ranks exchange is the metadata that turns N local
bitstreams into one: per-stream byte lengths (all-gather) -> global offset table, and -- only
when one contiguous payload is wanted on every rank -- the payload bytes themselves
(variable-size all-gather, done as one broadcast per rank into its slice of the output).
Works on any backend: NCCL over NVLink on the GPUs, gloo on CPU tensors in the tests.

Reference context: the reference codes one matrix = one stream = one file
(ISS/+coder/cabacEncode.m:34-37,97); sharding many such streams is the data parallelism
SURVEY.md 2.2 identifies, there is no reference code to mirror.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_streams: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous equal-count range of stream ids for `rank` (the first n % world ranks get one more)."""
    q, r = divmod(int(n_streams), int(world))
    a = rank * q + min(rank, r)
    return a, a + q + (1 if rank < r else 0)


def balanced_ranges(work_off, world: int, granule: int = 32) -> list[tuple[int, int]]:
    """Contiguous ranges with (nearly) equal WORK: work_off is the exclusive prefix sum of the
    per-stream work (bins or symbols), e.g. the op_off / sym_off table.  Boundaries are rounded to
    `granule` streams (one warp tile).  For skewed stream lengths."""
    off = np.asarray(work_off, dtype=np.int64)
    n = off.size - 1
    total = int(off[-1] - off[0])
    cuts = [0]
    for k in range(1, world):
        s = int(np.searchsorted(off, off[0] + total * k // world, side="left"))
        s = min(n, (s + granule // 2) // granule * granule)
        cuts.append(max(s, cuts[-1]))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


@dataclass
class GlobalTable:
    """Offset table of the whole job: stream g (global id) = payload[byte_off[g]:byte_off[g+1]]."""
    lengths: torch.Tensor        # int32 [n_total] (u32 values), global stream order
    byte_off: torch.Tensor       # int64 [n_total + 1]
    stream_counts: list[int]     # streams per rank
    rank_bytes: list[int]        # payload bytes per rank
    rank_base: list[int]         # first byte of every rank's slice in the assembled payload


def _world(group=None) -> tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def gather_table(local_lengths: torch.Tensor, group=None) -> GlobalTable:
    """All-gather the per-stream byte lengths of every rank (ranks may hold different numbers of
    streams) and build the global offset table on every rank.  Two small collectives: the counts
    (8 B per rank) and the lengths padded to the largest count (4 B per stream)."""
    rank, world = _world(group)
    dev = local_lengths.device
    n_local = int(local_lengths.numel())
    if world == 1:
        counts = [n_local]
        all_len = local_lengths.to(torch.int32)
    else:
        cnt = torch.tensor([n_local], dtype=torch.int64, device=dev)
        cnts = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(cnts, cnt, group=group)
        counts = [int(x) for x in cnts.tolist()]
        m = max(counts) if counts else 0
        pad = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
        pad[:n_local] = local_lengths.to(torch.int32)
        buf = torch.empty(world * max(m, 1), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(buf, pad, group=group)
        all_len = torch.cat([buf[r * max(m, 1): r * max(m, 1) + counts[r]] for r in range(world)])
    # lengths are u32 values kept in int32 storage
    l64 = all_len.to(torch.int64) & 0xFFFFFFFF
    byte_off = torch.zeros(l64.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(l64, 0, out=byte_off[1:])
    bounds = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    bo = byte_off[torch.as_tensor(bounds, device=dev)].tolist()
    rank_base = [int(x) for x in bo[:-1]]
    rank_bytes = [int(bo[i + 1] - bo[i]) for i in range(world)]
    return GlobalTable(all_len, byte_off, counts, rank_bytes, rank_base)


def assemble_payload(local_payload: torch.Tensor, table: GlobalTable, group=None) -> torch.Tensor:
    """One contiguous bitstream on every rank: rank r's compacted payload lands at
    table.rank_base[r].  Variable-size all-gather as one broadcast per rank, issued back to back
    (NCCL runs them in order on its stream; over NVSwitch each is a full-bandwidth one-to-all)."""
    rank, world = _world(group)
    total = int(table.byte_off[-1].item()) if table.byte_off.numel() else 0
    out = torch.empty(max(total, 1), dtype=torch.uint8, device=local_payload.device)
    mine = table.rank_bytes[rank]
    out[table.rank_base[rank]: table.rank_base[rank] + mine] = local_payload[:mine]
    if world > 1:
        for r in range(world):
            if table.rank_bytes[r]:
                seg = out[table.rank_base[r]: table.rank_base[r] + table.rank_bytes[r]]
                dist.broadcast(seg, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return out[:total]


def local_slice(table: GlobalTable, rank: int) -> tuple[int, int, int, int]:
    """(first stream, one past last stream, first byte, one past last byte) of `rank` in the global table."""
    a = int(sum(table.stream_counts[:rank]))
    b = a + table.stream_counts[rank]
    return a, b, table.rank_base[rank], table.rank_base[rank] + table.rank_bytes[rank]


def encode_ops_sharded(ops_local, op_off_local, ctx_init, group=None, assemble: bool = False, slab_stride=None):
    """Encode this rank's streams on its GPU, exchange the lengths, optionally assemble the
    payload.  -> (local Payload, GlobalTable, assembled payload or None)."""
    from . import engine as E
    enc = E.encode_ops(ops_local, op_off_local, ctx_init, slab_stride)
    pay = E.compact(enc)
    table = gather_table(enc.lengths, group)
    full = assemble_payload(pay.payload, table, group) if assemble else None
    return pay, table, full


def decode_ops_sharded(full_payload, table: GlobalTable, ops_local, op_off_local, ctx_init, group=None):
    """Decode this rank's streams out of an assembled (replicated) payload."""
    from . import engine as E
    rank, _ = _world(group)
    a, b, b0, b1 = local_slice(table, rank)
    boff = (table.byte_off[a: b + 1] - b0).contiguous()
    return E.decode_ops((full_payload[b0:b1], boff), ops_local, op_off_local, ctx_init)
