"""Stream sharding over the GPUs of one box (one process per GPU, torch.distributed).

Streams are independent (separate start()...finish() lifetimes, separate context sets), so the
coding loop never communicates: every rank codes a contiguous range of stream ids with the
single-GPU entry points of engine.py.  What ranks exchange is the metadata that turns N local
bitstreams into one: per-stream byte lengths (all-gather) -> global offset table, and -- only
when one contiguous payload is wanted on every rank -- the payload bytes themselves
(variable-size all-gather, done as one broadcast per rank into its slice of the output).
On the GPUs this module is a caller of the C ABI (include/isscabac.h: cabac_multi_gpu_*, csrc/multi_gpu.cu): the
library owns an NCCL communicator (MultiGpu), the length exchange + global scan run stream-ordered without a host
round trip, the payload is assembled by grouped NCCL broadcasts or -- fused into the compaction kernel -- by peer
stores over NVLink.  torch.distributed is used for what the brief assigns to it: process-group plumbing (shipping the
NCCL unique id) and, on CPU tensors over gloo in the tests, the same metadata exchange (gather_table /
assemble_payload below; no coding work happens there).

Reference context: the reference codes one matrix = one stream = one file
(ISS/+coder/cabacEncode.m:34-37,97); sharding many such streams is the data parallelism
SURVEY.md 2.2 identifies, there is no reference code to mirror.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from ._lib import CabacError, check, lib, vp


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
    out = torch.empty((max(total, 1) + 3) & ~3, dtype=torch.uint8, device=local_payload.device)[:max(total, 1)]
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


# ------------------------------------------------------------------------------------
# C-ABI path (CUDA tensors, NCCL inside libisscabac.so)
# ------------------------------------------------------------------------------------
class _CudaBuffer:
    """A device allocation owned by the library, visible to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class MultiGpu:
    """The library's multi-GPU handle (isscabac_mgpu): one per process, on the current CUDA device.  The NCCL unique
    id travels from rank 0 to the other ranks through torch.distributed (any backend)."""

    def __init__(self, group=None):
        self.rank, self.world = _world(group)
        L = lib()
        L.cabac_multi_gpu_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        dev = torch.device("cuda", torch.cuda.current_device())
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = np.zeros(128, dtype=np.uint8)
            check(L.cabac_multi_gpu_unique_id(vp(buf)))
            ident = torch.from_numpy(buf)
        if self.world > 1:
            backend = dist.get_backend(group)
            t = ident.to(dev) if backend == "nccl" else ident
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = t.cpu()
        self._h = C.c_void_p(0)
        check(L.cabac_multi_gpu_init(vp(ident.numpy()), self.rank, self.world, C.byref(self._h)))
        self._sym = None

    def close(self):
        if self._h:
            lib().cabac_multi_gpu_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def first_from_counts(counts) -> np.ndarray:
        f = np.zeros(len(counts) + 1, dtype=np.uint32)
        np.cumsum(np.asarray(counts, dtype=np.uint64), out=f[1:])
        return f

    def gather_table(self, first: np.ndarray, local_lengths: torch.Tensor, *, all_lengths=None, byte_off=None, scratch=None):
        """-> (all_lengths int32 [n_total], byte_off int64 [n_total + 1]) on the device, stream-ordered, no host sync."""
        dev = local_lengths.device
        first = np.ascontiguousarray(first, dtype=np.uint32)
        n_total = int(first[-1])
        L = lib()
        if all_lengths is None:
            all_lengths = torch.empty(max(n_total, 1), dtype=torch.int32, device=dev)
        if byte_off is None:
            byte_off = torch.empty(n_total + 1, dtype=torch.int64, device=dev)
        if scratch is None:
            scratch = torch.empty(int(L.cabac_compact_scratch_bytes(C.c_uint32(n_total))), dtype=torch.uint8, device=dev)
        check(L.cabac_multi_gpu_gather_table(self._h, vp(first), vp(local_lengths), vp(all_lengths), vp(byte_off), vp(scratch),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return all_lengths[:n_total], byte_off

    def assemble(self, first: np.ndarray, byte_off: torch.Tensor, local_payload: torch.Tensor, out: torch.Tensor):
        """Grouped NCCL broadcasts: rank r's payload lands at byte_off[first[r]] of `out` on every rank.
        -> rank_byte_first uint64 [world + 1] (host; the one small device-to-host read of the path)."""
        first = np.ascontiguousarray(first, dtype=np.uint32)
        rb = np.zeros(self.world + 1, dtype=np.uint64)
        check(lib().cabac_multi_gpu_assemble(self._h, vp(first), vp(byte_off), vp(local_payload), vp(out), C.c_uint64(out.numel()),
                                             vp(rb), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return rb

    def symmetric_alloc(self, nbytes: int) -> torch.Tensor:
        """One buffer of `nbytes` per rank, every rank's mapped into every other rank (CUDA IPC).  -> this rank's buffer."""
        L = lib()
        L.cabac_multi_gpu_symmetric_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        ptr = C.c_void_p(0)
        check(L.cabac_multi_gpu_symmetric_alloc(self._h, C.c_uint64(int(nbytes)), C.byref(ptr)))
        self._sym = torch.as_tensor(_CudaBuffer(int(ptr.value), int(nbytes)), device=torch.device("cuda", torch.cuda.current_device()))
        return self._sym

    def close_symmetric(self):
        """Unmap the peers' buffers and free this rank's (tensors returned by symmetric_alloc dangle afterwards)."""
        self._sym = None
        check(lib().cabac_multi_gpu_symmetric_free(self._h))

    def compact_p2p(self, first: np.ndarray, enc, byte_off: torch.Tensor):
        """Compaction fused with the payload exchange: every local stream is stored at its global offset into the
        symmetric buffer of EVERY rank (peer stores over NVLink); follow with barrier()."""
        first = np.ascontiguousarray(first, dtype=np.uint32)
        check(lib().cabac_multi_gpu_compact_p2p(self._h, vp(first), vp(enc.slab), C.c_uint64(enc.slab.shape[1]), vp(enc.lengths),
                                                vp(byte_off), vp(enc.overflow), C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def barrier(self):
        check(lib().cabac_multi_gpu_barrier(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))


_DEFAULT: MultiGpu | None = None


def default_handle(group=None) -> MultiGpu:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = MultiGpu(group)
    return _DEFAULT


def _first_of(n_local: int, group=None, first=None) -> np.ndarray:
    """Partition table: given, or from an all-gather of the local stream counts (host metadata, 8 B per rank)."""
    if first is not None:
        return np.ascontiguousarray(first, dtype=np.uint32)
    rank, world = _world(group)
    if world == 1:
        return np.array([0, n_local], dtype=np.uint32)
    counts = [None] * world
    dist.all_gather_object(counts, int(n_local), group=group)
    return MultiGpu.first_from_counts(counts)


def encode_ops_sharded(ops_local, op_off_local, ctx_init, group=None, assemble: bool = False, slab_stride=None, first=None,
                       fused_p2p: bool = False):
    """Encode this rank's streams on its GPU, exchange the lengths, optionally assemble the payload.
    -> (local Payload, GlobalTable, assembled payload or None).  `first` = the partition (global first stream id of every
    rank, world + 1 entries) when the caller knows it; otherwise the counts are all-gathered.
    A stream that outgrows the (heuristic) slab stride is re-encoded with the proof-level stride, like the host ABI does."""
    from . import engine as E
    rank, world = _world(group)
    enc = E.encode_ops(ops_local, op_off_local, ctx_init, slab_stride)
    if int(enc.overflow[0].item()) & 1:
        off_t = torch.as_tensor(np.asarray(op_off_local)) if not isinstance(op_off_local, torch.Tensor) else op_off_local
        longest = int((off_t[1:] - off_t[:-1]).max().item()) if off_t.numel() > 1 else 0
        enc = E.encode_ops(ops_local, op_off_local, ctx_init, E.slab_stride_bound(longest))
        enc.check_overflow()
    n_local = int(enc.lengths.numel())
    first = _first_of(n_local, group, first)
    mg = default_handle(group)
    all_len, byte_off = mg.gather_table(first, enc.lengths)
    full = None
    if assemble and fused_p2p:
        # capacity: host-known bound (2 bits per op is never reached by adaptive CABAC; the kernel flags an excess)
        rb_total = int(byte_off[-1].item())
        full_buf = mg.symmetric_alloc((max(rb_total, 16) + 3) & ~3)      # whole 32-bit words: the decoders read the payload in them
        mg.compact_p2p(first, enc, byte_off)
        mg.barrier()
        pay = E.Payload(full_buf[int(byte_off[int(first[rank])].item()): int(byte_off[int(first[rank + 1])].item())],
                        (byte_off[int(first[rank]): int(first[rank + 1]) + 1] - byte_off[int(first[rank])]).contiguous())
        full = full_buf[:rb_total]
        rbf = byte_off[torch.as_tensor(first.astype(np.int64), device=byte_off.device)].tolist()
    else:
        pay = E.compact(enc)
        if assemble:
            total = int(byte_off[-1].item())
            full = torch.empty((max(total, 1) + 3) & ~3, dtype=torch.uint8, device=enc.slab.device)[:max(total, 1)]
            rbf = [int(x) for x in mg.assemble(first, byte_off, pay.payload, full)]
            full = full[:total]
        else:
            rbf = byte_off[torch.as_tensor(first.astype(np.int64), device=byte_off.device)].tolist()
    if int(enc.overflow[0].item()) & 2:
        raise CabacError(-3, "payload capacity exceeded during compaction")
    counts = [int(first[r + 1] - first[r]) for r in range(world)]
    table = GlobalTable(all_len, byte_off, counts, [int(rbf[r + 1] - rbf[r]) for r in range(world)], [int(x) for x in rbf[:-1]])
    return pay, table, full


def decode_ops_sharded(full_payload, table: GlobalTable, ops_local, op_off_local, ctx_init, group=None):
    """Decode this rank's streams out of an assembled (replicated) payload."""
    from . import engine as E
    rank, _ = _world(group)
    a, b, b0, b1 = local_slice(table, rank)
    boff = (table.byte_off[a: b + 1] - b0).contiguous()
    return E.decode_ops((full_payload[b0:b1], boff), ops_local, op_off_local, ctx_init)
