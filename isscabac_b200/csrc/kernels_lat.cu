// kernels_lat.cu -- op-array kernels for the latency-bound regime (cabac_spec.cuh): few 32-stream tiles per
// SM, so one bin's dependent chain -- not the issue rate -- sets the time of the launch.  Same inputs, outputs
// and block schedule as k_encode_ops_wide / k_decode_ops_wide (kernels.cu); results are byte-identical.
//
// Shared memory per CTA:
//   tab  [128 states][LAT_COLS] SRow   16-byte rows; lane l reads column l % LAT_COLS.  An LDS.128 is served one
//                                      quarter-warp (8 lanes x 16 B = all 32 banks) at a time, so 8 columns already
//                                      keep 32 lanes with 32 unrelated states conflict-free: 16 KB instead of the
//                                      66 KB of the 32-column table of the wide kernels.
//   ctx  [NW][n_ctx + 1][32] SRow      the ROW of every context's state, one per (warp, context, lane)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <type_traits>

#include "../../include/isscabac.h"
#include "cabac_spec.cuh"

#ifndef CABAC_OPS_ASYNC
#define CABAC_OPS_ASYNC 0     // op blocks of the latency decoder through shared memory (cp.async) instead of registers
#endif
#include "codec_params.h"
#include "internal.h"

using namespace cabac;
using namespace isscabac_internal;

namespace {

#ifndef LAT_COLS
#define LAT_COLS 8
#endif
constexpr uint32_t LAT_MAX_WARPS = 16;
constexpr uint32_t LAT_TAB_BYTES = 128u * LAT_COLS * sizeof(SRow);

struct SpecRowTable {
  SRow r[128];
  constexpr SpecRowTable() : r{} {
    for (uint32_t i = 0; i < 128; ++i) r[i] = spec_row(i);
  }
};
__constant__ SpecRowTable c_spec_rows = SpecRowTable();

// row of state byte st with the successor tokens of column `col` (tab0 = shared-window address of the table)
__device__ __forceinline__ SRow lat_row(uint32_t st, uint32_t tab0, uint32_t col) {
  SRow r = c_spec_rows.r[st];
  r.tok_m = tab0 + ((r.tok_m * LAT_COLS + col) << 4);
  r.tok_l = tab0 + ((r.tok_l * LAT_COLS + col) << 4);
  return r;
}

struct LatMem {
  uint32_t ctx0;   // shared-window address of this lane's row of slot 0
  __device__ __forceinline__ SRow ldrow(uint32_t tok) const {
    SRow r;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.lps4), "=r"(r.nn4), "=r"(r.tok_m), "=r"(r.tok_l) : "r"(tok));
    return r;
  }
  // context rows: volatile + memory clobber pins the program order of the look-ahead load against the stores
  __device__ __forceinline__ SRow ldctx(uint32_t c) const {
    SRow r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.lps4), "=r"(r.nn4), "=r"(r.tok_m), "=r"(r.tok_l) : "r"(ctx0 + c * 512u) : "memory");
    return r;
  }
  __device__ __forceinline__ void stctx(uint32_t c, const SRow& r) const {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(ctx0 + c * 512u), "r"(r.lps4), "r"(r.nn4), "r"(r.tok_m), "r"(r.tok_l) : "memory");
  }
};

__device__ __forceinline__ bool lat_setup(const CodecParams& P, uint8_t* smem, uint32_t& s, LatMem& mem, uint32_t& n_ctx, uint32_t& vmask) {
  const uint32_t tab0 = (uint32_t)__cvta_generic_to_shared(smem);
  SRow* t = reinterpret_cast<SRow*>(smem);
  for (uint32_t i = threadIdx.x; i < 128u * LAT_COLS; i += blockDim.x) t[i] = lat_row(i / LAT_COLS, tab0, i % LAT_COLS);
  const uint32_t warp = threadIdx.x >> 5, lane = cb_keep32(threadIdx.x & 31);
  const uint32_t nw = blockDim.x >> 5;
  n_ctx = cb_keep32(P.n_ctx);
  s = (blockIdx.x * nw + warp) * 32 + lane;
  const bool valid = s < P.n_streams;
  mem.ctx0 = cb_keep32(tab0 + LAT_TAB_BYTES + (warp * (n_ctx + 1) * 32 + lane) * (uint32_t)sizeof(SRow));
  const uint8_t* init = P.ctx_init + (P.per_stream_init && valid ? (uint64_t)s * n_ctx : 0);
  const uint32_t col = lane % LAT_COLS;
  for (uint32_t c = 0; c < n_ctx; ++c) mem.stctx(c, lat_row(init[c] & 127u, tab0, col));
  mem.stctx(n_ctx, lat_row(0, tab0, col));     // the slot bypass ops address: any valid row
  __syncthreads();
  vmask = __ballot_sync(0xffffffffu, valid);
  return valid;
}

__global__ void __launch_bounds__(LAT_MAX_WARPS * 32) k_encode_ops_lat(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t s, n_ctx, vmask;
  LatMem mem;
  if (!lat_setup(P, smem, s, mem, n_ctx, vmask)) return;
  const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
  const uint8_t* p = reinterpret_cast<const uint8_t*>(P.ops) + o0;
  const uint64_t n = o1 - o0;

  EncWide E;
  const uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  encw_start(E, P.slab + (uint64_t)s * P.slab_stride, cap);

  uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
  if (head > n) head = n;
  for (uint64_t i = 0; i < head; ++i) encs_general(E, p[i], mem, n_ctx);
  p += head;
  const uint64_t nblk = (n - head) >> 4;
  const uint32_t tail = (uint32_t)((n - head) & 15u);
  const uint32_t common = __reduce_min_sync(vmask, (uint32_t)(nblk > 0xffffffffull ? 0xffffffffull : nblk));
  if (nblk) {
    uint4 cur = __ldg(reinterpret_cast<const uint4*>(p));
    uint64_t b = 0;
    auto block = [&](auto lock) {
      constexpr bool LOCK = decltype(lock)::value;
      uint4 nxt = cur;
      if (b + 1 < nblk) nxt = __ldg(reinterpret_cast<const uint4*>(p + 16));
      const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
      const uint32_t cw[4] = {op_codes4(cur.x), op_codes4(cur.y), op_codes4(cur.z), op_codes4(cur.w)};
      if (cb_any<LOCK>(block_has_trm(cw))) {
        for (int k = 0; k < 16; ++k) encs_general(E, p[k], mem, n_ctx);
      } else {
        encs_block16<LOCK>(E, w, cw, mem, n_ctx);
      }
      cur = nxt;
      p += 16;
    };
    if (vmask == 0xffffffffu)
      for (; b < common; ++b) block(std::true_type{});
    for (; b < nblk; ++b) block(std::false_type{});
  }
  for (uint32_t i = 0; i < tail; ++i) encs_general(E, p[i], mem, n_ctx);

  const uint32_t len = encw_finish(E);
  P.lengths[s] = len;
  if (len > cap && P.overflow) atomicOr(P.overflow, 1u);
}

__global__ void __launch_bounds__(LAT_MAX_WARPS * 32) k_decode_ops_lat(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t s, n_ctx, vmask;
  LatMem mem;
  if (!lat_setup(P, smem, s, mem, n_ctx, vmask)) return;
  const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
  const uint8_t* p = reinterpret_cast<const uint8_t*>(P.ops) + o0;
  uint8_t* q = P.bins + o0;
  const uint64_t n = o1 - o0;
  const uint64_t b0 = P.byte_off[s], b1 = P.byte_off[s + 1];

  DecWide D;
  decw_start(D, P.bytes + b0, (uint32_t)(b1 - b0));

  uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
  if (head > n) head = n;
  for (uint64_t i = 0; i < head; ++i) q[i] = (uint8_t)decs_general(D, p[i], mem, n_ctx);
  p += head;
  q += head;
  const uint64_t nblk = (n - head) >> 4;
  const uint32_t tail = (uint32_t)((n - head) & 15u);
  const bool out_vec = (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
  const uint32_t common = __reduce_min_sync(vmask, (uint32_t)(nblk > 0xffffffffull ? 0xffffffffull : nblk));
  if (nblk) {
#if CABAC_OPS_ASYNC
    // The op blocks come in through shared memory (cp.async, two blocks ahead, four 16-byte slots per lane) instead of a
    // register pair rotated with moves: the rotation's `cur = nxt` waited ~120 cycles per block on the load scoreboard it
    // shares with the window top-up (a word some lane requested a moment ago), although its own load was 16 bins old.
    const uint32_t ostage = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)LAT_TAB_BYTES + (blockDim.x >> 5) * (n_ctx + 1u) * 32u * (uint32_t)sizeof(SRow)
                            + threadIdx.x * 64u;
    const uint8_t* const pbase = p;
    auto feed = [&](uint64_t blk) {
      if (blk < nblk) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(ostage + ((uint32_t)blk & 3u) * 16u), "l"(pbase + 16ull * blk) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#endif
    uint4 cur = __ldg(reinterpret_cast<const uint4*>(p));
    uint64_t b = 0;
#if CABAC_OPS_ASYNC
    feed(1);
    feed(2);
#endif
    auto block = [&](auto lock) {
      constexpr bool LOCK = decltype(lock)::value;
#if !CABAC_OPS_ASYNC
      uint4 nxt = cur;
      if (b + 1 < nblk) nxt = __ldg(reinterpret_cast<const uint4*>(p + 16));
#endif
      const uint32_t cw[4] = {op_codes4(cur.x), op_codes4(cur.y), op_codes4(cur.z), op_codes4(cur.w)};
      if (cb_any<LOCK>(block_has_trm(cw))) {
        for (int k = 0; k < 16; ++k) q[k] = (uint8_t)decs_general(D, p[k], mem, n_ctx);
      } else {
        uint32_t r[4];
        decs_block16<LOCK>(D, cw, r, mem, n_ctx);
        if (out_vec) {
          *reinterpret_cast<uint4*>(q) = make_uint4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) q[k] = (uint8_t)(r[k >> 2] >> (8 * (k & 3)));
        }
      }
#if CABAC_OPS_ASYNC
      p += 16;
      q += 16;
      asm volatile("cp.async.wait_group 1;" ::: "memory");       // block b + 1 has landed (block b + 2 may be on its way)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(cur.x), "=r"(cur.y), "=r"(cur.z), "=r"(cur.w) : "r"(ostage + ((uint32_t)(b + 1) & 3u) * 16u) : "memory");
      feed(b + 3);
#else
      cur = nxt;
      p += 16;
      q += 16;
#endif
    };
    if (vmask == 0xffffffffu)
      for (; b < common; ++b) block(std::true_type{});
    for (; b < nblk; ++b) block(std::false_type{});
  }
  for (uint32_t i = 0; i < tail; ++i) q[i] = (uint8_t)decs_general(D, p[i], mem, n_ctx);

  if (P.finish_ok) P.finish_ok[s] = (uint8_t)decw_finish(D);
}

}  // namespace

namespace isscabac_internal {

int launch_lat_codec(bool encode, const CodecParams& P, cudaStream_t st, bool& done) {
  done = false;
  const size_t lim = smem_limit();
  const size_t warp_ctx = ((size_t)P.n_ctx + 1) * 32 * sizeof(SRow);
  if (!lim || P.n_ctx > 125 || LAT_TAB_BYTES + warp_ctx > lim) return ISSCABAC_OK;
  const uint32_t sms = (uint32_t)sm_count();
  const uint32_t tiles = (P.n_streams + 31) / 32;
  uint32_t nw_max = (uint32_t)((lim - LAT_TAB_BYTES) / (warp_ctx + (CABAC_OPS_ASYNC ? 32 * 64 : 0)));
  if (nw_max > LAT_MAX_WARPS) nw_max = LAT_MAX_WARPS;
  // tiles spread evenly over the SMs in whole CTAs
  const uint32_t ctas_per_sm = (tiles + sms * nw_max - 1) / (sms * nw_max);
  uint32_t nw = (tiles + sms * ctas_per_sm - 1) / (sms * ctas_per_sm);
  if (nw > nw_max) nw = nw_max;
  if (nw < 1) nw = 1;
  const uint32_t grid = (tiles + nw - 1) / nw;
  const size_t smem = LAT_TAB_BYTES + warp_ctx * nw + (CABAC_OPS_ASYNC && !encode ? (size_t)nw * 32 * 64 : 0);
  auto kernel = encode ? k_encode_ops_lat : k_decode_ops_lat;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<grid, nw * 32, smem, st>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, encode ? "k_encode_ops_lat" : "k_decode_ops_lat");
  done = true;
  return ISSCABAC_OK;
}

}  // namespace isscabac_internal
