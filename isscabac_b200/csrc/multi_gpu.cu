// multi_gpu.cu -- the multi-GPU entry points of the C ABI (include/isscabac.h: cabac_multi_gpu_*).
//
// Streams are independent (separate start() ... finish() lifetimes, separate context sets), so every rank codes its
// contiguous range of stream ids on its own GPU with the single-GPU kernels and the coding loop never communicates.
// This file holds what follows the coding: the all-gather-v of the per-stream lengths + the device-wide scan that gives
// every rank the global offset table, and the assembly of ONE contiguous bitstream on every rank -- either as grouped
// NCCL broadcasts, or fused into the compaction kernel as peer stores over NVLink (no host synchronisation).
// NCCL is bound with dlopen/dlsym at the first multi-GPU call: libisscabac.so has no link-time NCCL dependency.
//
// Reference context: nothing upstream is parallel; one stream = one file (ISS/+coder/cabacEncode.m:34-37,97,
// CABAC/SimpleCABACMex.cpp:195,288).  SURVEY.md 8(b).3 / 8(e).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/isscabac.h"
#include "internal.h"

using namespace isscabac_internal;

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

int nccl_api(NcclApi** out) {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (!g_nccl.lib) {
    // the copy already in the process first (PyTorch loads its own libnccl.so.2): communicators and streams must
    // belong to ONE NCCL instance
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return ISSCABAC_ERR_UNSUPPORTED; }
    NcclApi a;
    a.lib = h;
#define BIND(field, name)                                                              \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));                       \
  if (!a.field) { set_error("libnccl.so.2 lacks %s", name); return ISSCABAC_ERR_UNSUPPORTED; }
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(Broadcast, "ncclBroadcast")
    BIND(AllGather, "ncclAllGather")
    BIND(AllReduce, "ncclAllReduce")
    BIND(GroupStart, "ncclGroupStart")
    BIND(GroupEnd, "ncclGroupEnd")
    BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    g_nccl = a;
  }
  *out = &g_nccl;
  return ISSCABAC_OK;
}

int nccl_fail(NcclApi* N, ncclResult_t r, const char* what) {
  set_error("%s: %s", what, N->GetErrorString ? N->GetErrorString(r) : "NCCL error");
  return ISSCABAC_ERR_CUDA;
}
#define NK(call)                                                  \
  do {                                                            \
    ncclResult_t r__ = (call);                                    \
    if (r__ != ncclSuccess) return nccl_fail(N, r__, #call);      \
  } while (0)

constexpr int kMaxRanks = 64;

}  // namespace

struct isscabac_mgpu {
  ncclComm_t comm = nullptr;
  bool own_comm = false;
  int rank = 0, world = 1, device = 0;
  uint32_t* d_token = nullptr;            // barrier operand
  // symmetric allocation (CUDA IPC): the local buffer and every peer's mapping of its own
  uint8_t* sym_local = nullptr;
  uint64_t sym_bytes = 0;
  uint8_t* sym_peer[kMaxRanks] = {};
};

namespace {

int check_partition(const isscabac_mgpu* mg, const uint32_t* h_first, const char* who) {
  if (!mg || !h_first) { set_error("%s: null pointer", who); return ISSCABAC_ERR_INVALID; }
  if (h_first[0] != 0) { set_error("%s: h_first[0] must be 0", who); return ISSCABAC_ERR_INVALID; }
  for (int r = 0; r < mg->world; ++r)
    if (h_first[r + 1] < h_first[r]) { set_error("%s: h_first must be non-decreasing", who); return ISSCABAC_ERR_INVALID; }
  return ISSCABAC_OK;
}

struct PeerList {
  uint8_t* p[kMaxRanks];
};

// k_compact_copy (kernels.cu) with several destinations: one warp per local stream reads its slab row once and stores
// it at the stream's GLOBAL byte offset into the assembled payload of every rank.  All destinations are 256-byte
// aligned allocations, so the destination-word alignment -- and with it the funnel shift -- is the same for all of them.
template <int G>      // lanes per row, as k_compact_copy
__global__ void __launch_bounds__(256) k_compact_copy_peers(const uint8_t* slab, uint64_t stride, const uint32_t* lengths,
                                                             const uint64_t* off, PeerList dsts, int n_dst, uint64_t cap,
                                                             uint32_t n_streams, uint32_t* overflow) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / G, lane = threadIdx.x % G;
  if (warp >= n_streams) return;
  uint32_t len = lengths[warp];
  if (len > stride) len = (uint32_t)stride;
  const uint64_t d0 = off[warp];
  if (d0 + len > cap) {
    if (lane == 0 && overflow) atomicOr(overflow, 2u);
    return;
  }
  const uint8_t* src = slab + (uint64_t)warp * stride;
  uint32_t head = (uint32_t)((4 - (d0 & 3u)) & 3u);
  if (head > len) head = len;
  const uint32_t nwords = (len - head) >> 2;
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(src);
  const uint32_t sh = 8 * head;
  const uint32_t done = head + (nwords << 2);
  if (lane < head) {
    const uint8_t b = src[lane];
    for (int d = 0; d < n_dst; ++d) dsts.p[d][d0 + lane] = b;
  }
  uint32_t w = lane;
  for (; w + 3 * G < nwords; w += 4 * G) {
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t lo = sw[w + G * k], hi = sh ? sw[w + G * k + 1] : 0u;
      v[k] = sh ? __funnelshift_r(lo, hi, sh) : lo;
    }
    for (int d = 0; d < n_dst; ++d) {
      uint32_t* dw = reinterpret_cast<uint32_t*>(dsts.p[d] + d0 + head);
#pragma unroll
      for (int k = 0; k < 4; ++k) dw[w + G * k] = v[k];
    }
  }
  for (; w < nwords; w += G) {
    const uint32_t lo = sw[w], hi = sh ? sw[w + 1] : 0u;
    const uint32_t v = sh ? __funnelshift_r(lo, hi, sh) : lo;
    for (int d = 0; d < n_dst; ++d) reinterpret_cast<uint32_t*>(dsts.p[d] + d0 + head)[w] = v;
  }
  if (lane < len - done) {
    const uint8_t b = src[done + lane];
    for (int d = 0; d < n_dst; ++d) dsts.p[d][d0 + done + lane] = b;
  }
}

}  // namespace

extern "C" {

int cabac_multi_gpu_unique_id(uint8_t* h_id) {
  if (!h_id) { set_error("cabac_multi_gpu_unique_id: null pointer"); return ISSCABAC_ERR_INVALID; }
  NcclApi* N;
  int rc = nccl_api(&N);
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == ISSCABAC_MGPU_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NK(N->GetUniqueId(&id));
  memcpy(h_id, &id, sizeof id);
  return ISSCABAC_OK;
}

static int mgpu_finish_init(isscabac_mgpu* mg) {
  CK(cudaGetDevice(&mg->device));
  CK(cudaMalloc(reinterpret_cast<void**>(&mg->d_token), 256));
  CK(cudaMemset(mg->d_token, 0, 256));
  return ISSCABAC_OK;
}

int cabac_multi_gpu_init(const uint8_t* h_id, int rank, int world, isscabac_mgpu** out) {
  if (!h_id || !out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) {
    set_error("cabac_multi_gpu_init: bad arguments (world 1..%d)", kMaxRanks);
    return ISSCABAC_ERR_INVALID;
  }
  NcclApi* N;
  int rc = nccl_api(&N);
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, h_id, sizeof id);
  isscabac_mgpu* mg = new isscabac_mgpu();
  mg->rank = rank; mg->world = world; mg->own_comm = true;
  ncclResult_t r = N->CommInitRank(&mg->comm, world, id, rank);
  if (r != ncclSuccess) { delete mg; return nccl_fail(N, r, "ncclCommInitRank"); }
  if ((rc = mgpu_finish_init(mg))) { N->CommDestroy(mg->comm); delete mg; return rc; }
  *out = mg;
  return ISSCABAC_OK;
}

int cabac_multi_gpu_attach(void* nccl_comm, int rank, int world, isscabac_mgpu** out) {
  if (!nccl_comm || !out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) {
    set_error("cabac_multi_gpu_attach: bad arguments");
    return ISSCABAC_ERR_INVALID;
  }
  NcclApi* N;
  int rc = nccl_api(&N);
  if (rc) return rc;
  isscabac_mgpu* mg = new isscabac_mgpu();
  mg->comm = static_cast<ncclComm_t>(nccl_comm);
  mg->rank = rank; mg->world = world; mg->own_comm = false;
  if ((rc = mgpu_finish_init(mg))) { delete mg; return rc; }
  *out = mg;
  return ISSCABAC_OK;
}

int cabac_multi_gpu_destroy(isscabac_mgpu* mg) {
  if (!mg) return ISSCABAC_OK;
  cabac_multi_gpu_symmetric_free(mg);
  if (mg->d_token) cudaFree(mg->d_token);
  if (mg->own_comm && mg->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(mg->comm);
  delete mg;
  return ISSCABAC_OK;
}

int cabac_multi_gpu_info(const isscabac_mgpu* mg, int* rank, int* world) {
  if (!mg) { set_error("cabac_multi_gpu_info: null handle"); return ISSCABAC_ERR_INVALID; }
  if (rank) *rank = mg->rank;
  if (world) *world = mg->world;
  return ISSCABAC_OK;
}

int cabac_multi_gpu_gather_table(isscabac_mgpu* mg, const uint32_t* h_first, const uint32_t* d_local_lengths,
                                 uint32_t* d_all_lengths, uint64_t* d_byte_off, void* d_scratch, void* stream) {
  int rc = check_partition(mg, h_first, "cabac_multi_gpu_gather_table");
  if (rc) return rc;
  if (!d_all_lengths || !d_byte_off || !d_scratch) { set_error("cabac_multi_gpu_gather_table: null pointer"); return ISSCABAC_ERR_INVALID; }
  NcclApi* N;
  if ((rc = nccl_api(&N))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t n_total = h_first[mg->world];
  if (mg->world == 1) {
    if (n_total && d_local_lengths != d_all_lengths)
      CK(cudaMemcpyAsync(d_all_lengths, d_local_lengths, (size_t)n_total * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    NK(N->GroupStart());
    for (int r = 0; r < mg->world; ++r) {
      const size_t cnt = h_first[r + 1] - h_first[r];
      if (!cnt) continue;
      uint32_t* dst = d_all_lengths + h_first[r];
      ncclResult_t res = N->Broadcast(r == mg->rank ? (const void*)d_local_lengths : (const void*)dst, dst, cnt, ncclUint32, r, mg->comm, st);
      if (res != ncclSuccess) { N->GroupEnd(); return nccl_fail(N, res, "ncclBroadcast(lengths)"); }
    }
    NK(N->GroupEnd());
  }
  return exclusive_scan_u32_u64(d_all_lengths, d_byte_off, n_total, d_scratch, st);
}

int cabac_multi_gpu_assemble(isscabac_mgpu* mg, const uint32_t* h_first, const uint64_t* d_byte_off,
                             const uint8_t* d_local_payload, uint8_t* d_payload, uint64_t payload_cap,
                             uint64_t* h_rank_byte_first, void* stream) {
  int rc = check_partition(mg, h_first, "cabac_multi_gpu_assemble");
  if (rc) return rc;
  if (!d_byte_off || !d_payload) { set_error("cabac_multi_gpu_assemble: null pointer"); return ISSCABAC_ERR_INVALID; }
  NcclApi* N;
  if ((rc = nccl_api(&N))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // NCCL takes its counts from the host: the world + 1 boundary offsets come back first (the one sync of this path)
  uint64_t base[kMaxRanks + 1];
  for (int r = 0; r <= mg->world; ++r)
    CK(cudaMemcpyAsync(&base[r], d_byte_off + h_first[r], sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (h_rank_byte_first) memcpy(h_rank_byte_first, base, sizeof(uint64_t) * (mg->world + 1));
  if (base[mg->world] > payload_cap) {
    set_error("assembled payload needs %llu bytes, capacity %llu", (unsigned long long)base[mg->world], (unsigned long long)payload_cap);
    return ISSCABAC_ERR_OVERFLOW;
  }
  const uint64_t mine = base[mg->rank + 1] - base[mg->rank];
  if (mine && !d_local_payload) { set_error("cabac_multi_gpu_assemble: local payload is NULL"); return ISSCABAC_ERR_INVALID; }
  if (mg->world == 1) {
    if (mine && d_local_payload != d_payload) CK(cudaMemcpyAsync(d_payload, d_local_payload, mine, cudaMemcpyDeviceToDevice, st));
    return ISSCABAC_OK;
  }
  NK(N->GroupStart());
  for (int r = 0; r < mg->world; ++r) {
    const uint64_t cnt = base[r + 1] - base[r];
    if (!cnt) continue;
    uint8_t* dst = d_payload + base[r];
    ncclResult_t res = N->Broadcast(r == mg->rank ? (const void*)d_local_payload : (const void*)dst, dst, cnt, ncclUint8, r, mg->comm, st);
    if (res != ncclSuccess) { N->GroupEnd(); return nccl_fail(N, res, "ncclBroadcast(payload)"); }
  }
  NK(N->GroupEnd());
  return ISSCABAC_OK;
}

int cabac_multi_gpu_barrier(isscabac_mgpu* mg, void* stream) {
  if (!mg) { set_error("cabac_multi_gpu_barrier: null handle"); return ISSCABAC_ERR_INVALID; }
  if (mg->world == 1) return ISSCABAC_OK;
  NcclApi* N;
  int rc = nccl_api(&N);
  if (rc) return rc;
  NK(N->AllReduce(mg->d_token, mg->d_token + 16, 1, ncclUint32, ncclSum, mg->comm, static_cast<cudaStream_t>(stream)));
  return ISSCABAC_OK;
}

int cabac_multi_gpu_symmetric_free(isscabac_mgpu* mg) {
  if (!mg) return ISSCABAC_OK;
  for (int r = 0; r < mg->world; ++r) {
    if (mg->sym_peer[r] && r != mg->rank) cudaIpcCloseMemHandle(mg->sym_peer[r]);
    mg->sym_peer[r] = nullptr;
  }
  if (mg->sym_local) cudaFree(mg->sym_local);
  mg->sym_local = nullptr;
  mg->sym_bytes = 0;
  return ISSCABAC_OK;
}

int cabac_multi_gpu_symmetric_alloc(isscabac_mgpu* mg, uint64_t bytes, uint8_t** d_local) {
  if (!mg || !d_local || !bytes) { set_error("cabac_multi_gpu_symmetric_alloc: bad arguments"); return ISSCABAC_ERR_INVALID; }
  NcclApi* N;
  int rc = nccl_api(&N);
  if (rc) return rc;
  cabac_multi_gpu_symmetric_free(mg);
  CK(cudaMalloc(reinterpret_cast<void**>(&mg->sym_local), bytes));
  mg->sym_bytes = bytes;
  mg->sym_peer[mg->rank] = mg->sym_local;
  if (mg->world > 1) {
    // exchange the IPC handles through the communicator itself (64 bytes per rank)
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, mg->sym_local));
    uint8_t* d_h = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&d_h), sizeof(mine) * (size_t)(mg->world + 1)));
    CK(cudaMemcpy(d_h + sizeof(mine) * mg->world, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    ncclResult_t res = N->AllGather(d_h + sizeof(mine) * mg->world, d_h, sizeof(mine), ncclUint8, mg->comm, nullptr);
    if (res != ncclSuccess) { cudaFree(d_h); return nccl_fail(N, res, "ncclAllGather(ipc handles)"); }
    std::vector<cudaIpcMemHandle_t> all(mg->world);
    cudaError_t e = cudaMemcpy(all.data(), d_h, sizeof(mine) * (size_t)mg->world, cudaMemcpyDeviceToHost);
    cudaFree(d_h);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(ipc handles)");
    for (int r = 0; r < mg->world; ++r) {
      if (r == mg->rank) continue;
      void* p = nullptr;
      e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle (peer access between the GPUs of the box is required)");
      mg->sym_peer[r] = static_cast<uint8_t*>(p);
    }
  }
  *d_local = mg->sym_local;
  return ISSCABAC_OK;
}

int cabac_multi_gpu_compact_p2p(isscabac_mgpu* mg, const uint32_t* h_first, const uint8_t* d_slab, uint64_t slab_stride,
                                const uint32_t* d_local_lengths, const uint64_t* d_byte_off, uint32_t* d_overflow,
                                void* stream) {
  int rc = check_partition(mg, h_first, "cabac_multi_gpu_compact_p2p");
  if (rc) return rc;
  if (!mg->sym_local) { set_error("cabac_multi_gpu_compact_p2p: call cabac_multi_gpu_symmetric_alloc first"); return ISSCABAC_ERR_STATE; }
  const uint32_t n_local = h_first[mg->rank + 1] - h_first[mg->rank];
  if (!n_local) return ISSCABAC_OK;
  if (!d_slab || !d_local_lengths || !d_byte_off) { set_error("cabac_multi_gpu_compact_p2p: null pointer"); return ISSCABAC_ERR_INVALID; }
  PeerList pl;
  // the own buffer first, then the peers starting behind this rank (spreads the NVLink targets over the ranks)
  for (int k = 0; k < mg->world; ++k) pl.p[k] = mg->sym_peer[(mg->rank + k) % mg->world];
  // A warp per row whatever the row length: the local copy gains from 8 lanes per short row (kernels.cu), the peer stores do not --
  // 32-byte store groups over NVLink instead of 128-byte ones: C4 on 8 GPUs 1.81 -> 2.15 ms with 8 lanes (profiles/r2_v15_bench_n8.json)
  // (a single rank has no peer: 8 lanes per short row like the local copy)
  if (mg->world == 1 && slab_stride <= 2048) {
    const uint32_t blocks = (uint32_t)(((uint64_t)n_local * 8 + 255) / 256);
    k_compact_copy_peers<8><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_slab, slab_stride, d_local_lengths, d_byte_off + h_first[mg->rank], pl, mg->world, mg->sym_bytes, n_local, d_overflow);
  } else {
    const uint32_t blocks = (uint32_t)(((uint64_t)n_local * 32 + 255) / 256);
    k_compact_copy_peers<32><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_slab, slab_stride, d_local_lengths, d_byte_off + h_first[mg->rank], pl, mg->world, mg->sym_bytes, n_local, d_overflow);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_compact_copy_peers");
}

}  // extern "C"
