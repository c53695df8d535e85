// stats.cu -- context-model statistics computed on the device from the op arrays
// (SURVEY.md 8(f) rank 3).
//
// What the reference collects, one MEX call per bin:
//   * ctxHist / ctxCost of cabacEncode.m:40-41,61-65: bins coded per context and the bits
//     getNumBits() advanced by while a bin of that context was coded;
//   * under RWTH_TRACE_CABAC_STATES (CommonDef.h:39-43, Windows builds): per context a
//     histogram over the 128 "trace states" visited BEFORE each update
//     (ContextModel.cpp:97-104: index = mps == 0 ? 63 - state : state + 64), a 128x128
//     matrix of (state before -> state after) transition counts and the step log
//     [bin, state_p, mps_p, state_a, mps_a] (ContextModel.cpp:126-134,
//     SimpleCABACMex.cpp:231-241), exported by getEncoderStats / getDecoderStats (:356-466).
// The state sequence of a context depends only on the bins coded with it, so all of these
// except the cost need no arithmetic coder; the cost runs the 32-bit lane encoder
// (cabac_lane.cuh) without a sink and reads its getNumberOfWrittenBits() counter.
//
// One lane per stream; context states live in a [ctx][stream] byte scratch (coalesced per
// context, any n_ctx up to 999); counters are u64 / u32 atomics in global memory.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <type_traits>

#include "../../include/isscabac.h"
#include "cabac_lane.cuh"
#include "internal.h"

using namespace cabac;
using namespace isscabac_internal;

namespace {

constexpr int NT = 128;

struct RowTable {
  uint2 r[128];
  constexpr RowTable() : r{} {
    for (uint32_t i = 0; i < 128; ++i) r[i] = fused_row(i);
  }
};
__constant__ RowTable c_rows_s = RowTable();

struct TraceParams {
  uint32_t n_streams, n_ctx, spg;
  int per_stream_init;
  const uint64_t* op_off;
  const void* ops;
  const uint8_t* ctx_init;
  uint8_t* scratch;                 // [n_ctx][n_streams]
  unsigned long long* hist;         // [groups][n_ctx][128]
  uint32_t* trans;                  // [groups][n_ctx][128][128] or null
  unsigned long long* cost;         // [groups][n_ctx + 1] or null
  uint8_t* steps;                   // [n_ops][2] or null
  uint8_t* final_ctx;               // [n_streams][n_ctx] or null
};

__device__ __forceinline__ uint32_t trace_state(uint32_t st) {   // ContextModel.cpp:99-101
  return (st & 1u) ? (st >> 1) + 64u : 63u - (st >> 1);
}

template <int W>
__global__ void __launch_bounds__(NT) k_ctx_trace(TraceParams P) {
  __shared__ uint2 tab[128];
  if (threadIdx.x < 128) tab[threadIdx.x] = c_rows_s.r[threadIdx.x];
  __syncthreads();
  const uint32_t s = blockIdx.x * NT + threadIdx.x;
  if (s >= P.n_streams) return;
  uint8_t* my = P.scratch + s;
  const uint64_t cstride = P.n_streams;
  const uint8_t* init = P.ctx_init + (P.per_stream_init ? (uint64_t)s * P.n_ctx : 0);
  for (uint32_t c = 0; c < P.n_ctx; ++c) my[c * cstride] = init[c] & 127u;
  const uint64_t g = s / P.spg;
  unsigned long long* hist = P.hist + g * P.n_ctx * 128ull;
  uint32_t* trans = P.trans ? P.trans + g * P.n_ctx * 16384ull : nullptr;
  unsigned long long* cost = P.cost ? P.cost + g * (P.n_ctx + 1ull) : nullptr;

  typedef typename std::conditional<W == 1, uint8_t, uint16_t>::type OpT;
  constexpr uint32_t TRM = W == 1 ? ISSCABAC_OP8_TRM : ISSCABAC_OP16_TRM;
  const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
  const OpT* ops = reinterpret_cast<const OpT*>(P.ops);
  EncLane L;
  enc_start(L, nullptr, 0);
  for (uint64_t i = o0; i < o1; ++i) {
    const uint32_t o = ops[i], code = o >> 1, bin = o & 1u;
    const uint32_t before = cost ? enc_bits_written(L) : 0u;
    uint32_t slot = P.n_ctx;
    uint32_t sp = 0xffu, sa = 0xffu;
    if (code < TRM && code < P.n_ctx) {
      slot = code;
      uint32_t st = my[code * cstride];
      sp = st;
      const uint2 row = tab[st];
      if (cost) enc_bin_ctx<true>(L, bin, st, row);
      else st = cb_perm(row.y, 0, ((st ^ bin) & 1u) | 0x4440u);
      sa = st;
      my[code * cstride] = (uint8_t)st;
      atomicAdd(hist + code * 128ull + trace_state(sp), 1ull);
      if (trans) atomicAdd(trans + code * 16384ull + trace_state(sp) * 128u + trace_state(sa), 1u);
    } else if (cost) {
      if (code == TRM) enc_bin_trm<true>(L, bin);
      else enc_bin_ep<true>(L, bin);
    }
    if (cost) {
      const uint32_t d = enc_bits_written(L) - before;
      if (d) atomicAdd(cost + slot, (unsigned long long)d);
    }
    if (P.steps) {
      P.steps[2 * i] = (uint8_t)sp;
      P.steps[2 * i + 1] = (uint8_t)sa;
    }
  }
  if (P.final_ctx)
    for (uint32_t c = 0; c < P.n_ctx; ++c) P.final_ctx[(uint64_t)s * P.n_ctx + c] = my[c * cstride];
}

}  // namespace

extern "C" int cabac_ctx_trace_ops(uint32_t n_streams, const uint64_t* d_op_off, const void* d_ops, int op_width,
                                   const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                                   uint32_t streams_per_group, uint64_t* d_state_hist, uint32_t* d_trans,
                                   uint64_t* d_cost_bits, uint8_t* d_step_states, uint8_t* d_final_ctx, void* stream) {
  if (op_width != 1 && op_width != 2) { set_error("op_width must be 1 or 2"); return ISSCABAC_ERR_INVALID; }
  if (n_ctx > ISSCABAC_MAX_CTX || (op_width == 1 && n_ctx > 125)) { set_error("cabac_ctx_trace_ops: n_ctx %u out of range for this op width", n_ctx); return ISSCABAC_ERR_INVALID; }
  if (!streams_per_group) { set_error("streams_per_group must be >= 1"); return ISSCABAC_ERR_INVALID; }
  if (n_streams && (!d_op_off || !d_state_hist || (n_ctx && !d_ctx_init))) { set_error("cabac_ctx_trace_ops: null pointer"); return ISSCABAC_ERR_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  const uint64_t groups = ((uint64_t)n_streams + streams_per_group - 1) / streams_per_group;
  if (groups && n_ctx) {
    CK(cudaMemsetAsync(d_state_hist, 0, groups * n_ctx * 128ull * 8, st));
    if (d_trans) CK(cudaMemsetAsync(d_trans, 0, groups * n_ctx * 16384ull * 4, st));
  }
  if (groups && d_cost_bits) CK(cudaMemsetAsync(d_cost_bits, 0, groups * (n_ctx + 1ull) * 8, st));
  if (!n_streams) return ISSCABAC_OK;
  int rc = keep_pool_cached();
  if (rc) return rc;
  void* scratch = nullptr;
  CK(cudaMallocAsync(&scratch, (size_t)n_ctx * n_streams + 16, st));
  TraceParams P;
  memset(&P, 0, sizeof P);
  P.n_streams = n_streams; P.n_ctx = n_ctx; P.spg = streams_per_group; P.per_stream_init = per_stream_init;
  P.op_off = d_op_off; P.ops = d_ops; P.ctx_init = d_ctx_init; P.scratch = static_cast<uint8_t*>(scratch);
  P.hist = reinterpret_cast<unsigned long long*>(d_state_hist); P.trans = d_trans;
  P.cost = reinterpret_cast<unsigned long long*>(d_cost_bits); P.steps = d_step_states; P.final_ctx = d_final_ctx;
  const uint32_t grid = (n_streams + NT - 1) / NT;
  if (op_width == 1) k_ctx_trace<1><<<grid, NT, 0, st>>>(P);
  else k_ctx_trace<2><<<grid, NT, 0, st>>>(P);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(scratch, st);
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_ctx_trace");
}
