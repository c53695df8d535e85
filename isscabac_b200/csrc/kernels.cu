// kernels.cu -- sm_100a kernels of the many-stream CABAC engine + the device-pointer C ABI.
//
// Work decomposition: one LANE (thread) per CABAC stream -- a stream is a serial
// recurrence on (low, range) resp. (value, range), so the data parallelism is across
// streams.  Per CTA of 128 lanes, shared memory holds
//   * the fused state table (cabac_lane.cuh: LPS sub-ranges + both next states, 8 B per
//     state byte), REPLICATED PER LANE: row st of lane l lives at tab[st*32 + l], so the
//     32 lanes of a warp always hit 32 distinct banks (LDS.64, conflict-free by
//     construction whatever the 32 states are);
//   * the context states, one 32-bit word per (context, lane): ctx[c*128 + tid] -- bank =
//     lane, again conflict-free for arbitrary per-lane context indices.
// The arithmetic-coder registers (low/range/bitsLeft, output accumulator) stay in
// registers; output bytes are packed four at a time and stored as 32-bit words to the
// lane's slab (the L2 merges the partial sectors); a later scan+compaction pass builds the
// contiguous bitstream.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <type_traits>

#include "../../include/isscabac.h"
#include "cabac_lane.cuh"
#include "cabac_wide.cuh"
#include "wide_common.cuh"
#include "internal.h"
#include "codec_params.h"

using namespace cabac;
using isscabac_internal::CodecParams;

namespace {

constexpr int NT = 128;            // lanes (streams) per CTA
constexpr int TAB_WORDS = 128 * 32;  // uint2 entries of the lane-replicated table

struct RowTable {
  uint2 r[128];
  constexpr RowTable() : r{} {
    for (uint32_t i = 0; i < 128; ++i) r[i] = fused_row(i);
  }
};
__constant__ RowTable c_rows = RowTable();

__device__ __forceinline__ void fill_table(uint2* tab) {
  for (int i = threadIdx.x; i < TAB_WORDS; i += blockDim.x) tab[i] = c_rows.r[i >> 5];
}

template <int W>
struct OpTraits;
template <>
struct OpTraits<1> {
  typedef uint8_t T;
  static constexpr uint32_t EP = ISSCABAC_OP8_EP, TRM = ISSCABAC_OP8_TRM;
};
template <>
struct OpTraits<2> {
  typedef uint16_t T;
  static constexpr uint32_t EP = ISSCABAC_OP16_EP, TRM = ISSCABAC_OP16_TRM;
};

// Context storage policies ----------------------------------------------------
// smem: word per (context, lane).  gmem: byte per (context, stream) in a scratch
// array laid out [ctx][stream] so that a warp's accesses to one context coalesce
// (used when n_ctx is too large for shared memory, up to the API limit of 999).
struct CtxSmem {
  uint32_t* p;  // this lane's column
  __device__ __forceinline__ uint32_t load(uint32_t c) const { return p[c * NT]; }
  __device__ __forceinline__ void store(uint32_t c, uint32_t v) const { p[c * NT] = v; }
};
struct CtxGmem {
  uint8_t* p;  // this stream's column
  uint64_t stride;
  __device__ __forceinline__ uint32_t load(uint32_t c) const { return p[c * stride]; }
  __device__ __forceinline__ void store(uint32_t c, uint32_t v) const { p[c * stride] = (uint8_t)v; }
};

template <class Ctx>
__device__ __forceinline__ Ctx make_ctx(const CodecParams& P, uint8_t* smem_after_tab, uint32_t s, bool valid);

template <>
__device__ __forceinline__ CtxSmem make_ctx<CtxSmem>(const CodecParams& P, uint8_t* smem_after_tab, uint32_t s, bool valid) {
  uint32_t* base = reinterpret_cast<uint32_t*>(smem_after_tab);
  const uint8_t* init = P.ctx_init + (P.per_stream_init && valid ? (uint64_t)s * P.n_ctx : 0);
  for (uint32_t c = 0; c < P.n_ctx; ++c) base[c * NT + threadIdx.x] = init[c];
  return CtxSmem{base + threadIdx.x};
}
template <>
__device__ __forceinline__ CtxGmem make_ctx<CtxGmem>(const CodecParams& P, uint8_t*, uint32_t s, bool valid) {
  CtxGmem g{P.ctx_scratch + s, P.n_streams};
  if (valid) {
    const uint8_t* init = P.ctx_init + (P.per_stream_init ? (uint64_t)s * P.n_ctx : 0);
    for (uint32_t c = 0; c < P.n_ctx; ++c) g.store(c, init[c]);
  }
  return g;
}

// ---------------------------------------------------------------------------
// encode: one op
// ---------------------------------------------------------------------------
// Context-coded and bypass bins share one straight-line sequence (the 32 lanes of a warp
// hold unrelated streams, so a branch on the op kind would execute both sides anyway):
// the context path is computed unconditionally on a clamped context index, the bypass
// result is blended in with selects, and a single write-out test follows.  Only the
// terminate bin -- at most a handful per stream -- is a real branch.
template <int W, class Ctx>
__device__ __forceinline__ void encode_one(EncLane& L, uint32_t o, const Ctx& ctx, const uint2* mytab, uint32_t ctx_max, uint32_t n_ctx) {
  typedef OpTraits<W> OT;
  const uint32_t code = o >> 1, bin = o & 1u;
  if (code == OT::TRM) {
    enc_bin_trm<false>(L, bin);
    return;
  }
  const bool is_ep = code >= n_ctx;       // the bypass code, and by definition any code that is not a context of this call
  const uint32_t c = min(code, ctx_max);  // clamped: the context path is computed unconditionally (never out of bounds)
  uint32_t st = ctx.load(c);
  const uint2 row = mytab[st * 32];
  // encodeBin (Encoder.cpp:113-178), see enc_bin_ctx in cabac_lane.cuh
  const uint32_t lps = cb_perm(0, row.x, L.range >> 6);
  const uint32_t rmps = L.range - lps;
  const uint32_t is_lps = (st ^ o) & 1u;
  const uint32_t rsel = is_lps ? lps : rmps;
  int n = cb_clz(rsel) - 23;
  n = n > 6 ? 6 : n;
  const uint32_t low_c = (L.low + (is_lps ? rmps : 0u)) << n;
  const uint32_t range_c = rsel << n;
  st = cb_perm(row.y, 0, is_lps | 0x4440u);
  // encodeBinEP (Encoder.cpp:250-270)
  const uint32_t low_e = (L.low << 1) + (bin ? L.range : 0u);
  L.low = is_ep ? low_e : low_c;
  L.range = is_ep ? L.range : range_c;
  L.bits_left -= is_ep ? 1 : n;
  if (!is_ep) ctx.store(c, st);
  if (L.bits_left < 12) enc_write_out<false>(L);
}

template <int W, class Ctx>
__global__ void __launch_bounds__(NT) k_encode_ops(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint2* tab = reinterpret_cast<uint2*>(smem);
  fill_table(tab);
  const uint32_t s = blockIdx.x * NT + threadIdx.x;
  const bool valid = s < P.n_streams;
  Ctx ctx = make_ctx<Ctx>(P, smem + TAB_WORDS * sizeof(uint2), valid ? s : 0, valid);
  __syncthreads();
  if (!valid) return;
  const uint2* mytab = tab + (threadIdx.x & 31);
  const uint32_t ctx_max = P.n_ctx ? P.n_ctx - 1 : 0;

  typedef typename OpTraits<W>::T OpT;
  const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
  const OpT* p = reinterpret_cast<const OpT*>(P.ops) + o0;
  uint64_t n = o1 - o0;

  EncLane L;
  uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  enc_start(L, P.slab + (uint64_t)s * P.slab_stride, cap);

  uint64_t i = 0;
  // head: up to the first 16-byte boundary
  while (i < n && (reinterpret_cast<uintptr_t>(p + i) & 15u)) {
    encode_one<W, Ctx>(L, p[i], ctx, mytab, ctx_max, P.n_ctx);
    ++i;
  }
  // body: 16 bytes of ops per load, next block prefetched while this one is coded
  constexpr int PER = 16 / W;
  if (i + PER <= n) {
    uint4 cur = __ldg(reinterpret_cast<const uint4*>(p + i));
    for (;;) {
      const bool more = i + 2 * PER <= n;
      uint4 nxt = cur;
      if (more) nxt = __ldg(reinterpret_cast<const uint4*>(p + i + PER));
      const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int b = 0; b < 4 / W; ++b) {
          uint32_t o = (W == 1) ? ((w[k] >> (8 * b)) & 0xffu) : ((w[k] >> (16 * b)) & 0xffffu);
          encode_one<W, Ctx>(L, o, ctx, mytab, ctx_max, P.n_ctx);
        }
      }
      i += PER;
      if (!more) break;
      cur = nxt;
    }
  }
  // tail
  for (; i < n; ++i) encode_one<W, Ctx>(L, p[i], ctx, mytab, ctx_max, P.n_ctx);

  enc_finish<false>(L);
  enc_flush_pending(L);
  P.lengths[s] = L.nbytes;
  if ((L.overflow || L.nbytes > cap) && P.overflow) atomicOr(P.overflow, 1u);
}

// ---------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------
template <int W, class Ctx>
__device__ __forceinline__ uint32_t decode_one(DecLane& D, uint32_t o, const Ctx& ctx, const uint2* mytab, uint32_t ctx_max, uint32_t n_ctx) {
  typedef OpTraits<W> OT;
  const uint32_t code = o >> 1;
  if (code == OT::TRM) return dec_bin_trm(D);
  const bool is_ep = code >= n_ctx;
  const uint32_t c = min(code, ctx_max);
  uint32_t st = ctx.load(c);
  const uint2 row = mytab[st * 32];
  // decodeBinEP shifts (and possibly reads) BEFORE its compare (Decoder.cpp:288-331),
  // decodeBin compares first and renormalises after (Decoder.cpp:87-190).  One byte is
  // consumed per op at most, so both orders share a single read: the bypass pre-shift
  // reserves the byte position, the byte itself is added when it is fetched below --
  // adding it before or after the compare/subtract is the same because the compare is
  // made against a value whose low bits (the not-yet-read byte) are below one unit only
  // for the context path; for the bypass path the byte is needed first, so it is fetched
  // up front there.
  uint32_t value = D.value;
  int bn = D.bits_needed;
  uint32_t bin;
  if (is_ep) {
    value <<= 1;
    if (++bn >= 0) { bn = -8; value += dec_read(D); }
    const uint32_t scaled = D.range << 7;
    bin = value >= scaled ? 1u : 0u;
    value -= bin ? scaled : 0u;
  } else {
    const uint32_t lps = cb_perm(0, row.x, D.range >> 6);
    const uint32_t rmps = D.range - lps;
    const uint32_t scaled = rmps << 7;
    const uint32_t is_lps = value >= scaled ? 1u : 0u;
    const uint32_t rsel = is_lps ? lps : rmps;
    int n = cb_clz(rsel) - 23;
    n = n > 6 ? 6 : n;
    value = (value - (is_lps ? scaled : 0u)) << n;
    D.range = rsel << n;
    bin = (st ^ is_lps) & 1u;
    st = cb_perm(row.y, 0, is_lps | 0x4440u);
    ctx.store(c, st);
    bn += n;
    if (bn >= 0) { value += dec_read(D) << bn; bn -= 8; }
  }
  D.value = value;
  D.bits_needed = bn;
  return bin;
}

template <int W, class Ctx>
__global__ void __launch_bounds__(NT) k_decode_ops(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint2* tab = reinterpret_cast<uint2*>(smem);
  fill_table(tab);
  const uint32_t s = blockIdx.x * NT + threadIdx.x;
  const bool valid = s < P.n_streams;
  Ctx ctx = make_ctx<Ctx>(P, smem + TAB_WORDS * sizeof(uint2), valid ? s : 0, valid);
  __syncthreads();
  if (!valid) return;
  const uint2* mytab = tab + (threadIdx.x & 31);
  const uint32_t ctx_max = P.n_ctx ? P.n_ctx - 1 : 0;

  typedef typename OpTraits<W>::T OpT;
  const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
  const OpT* p = reinterpret_cast<const OpT*>(P.ops) + o0;
  uint8_t* q = P.bins + o0;
  uint64_t n = o1 - o0;

  const uint64_t b0 = P.byte_off[s], b1 = P.byte_off[s + 1];
  DecLane D;
  dec_start(D, P.bytes + b0, (uint32_t)(b1 - b0));

  uint64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(p + i) & 15u)) {
    q[i] = (uint8_t)decode_one<W, Ctx>(D, p[i], ctx, mytab, ctx_max, P.n_ctx);
    ++i;
  }
  constexpr int PER = 16 / W;
  // the bins of one op block are packed and stored with one vector store when the output
  // address is aligned too (it is whenever P.bins is 16-byte aligned and W == 1)
  const bool out_vec = (W == 1) && ((reinterpret_cast<uintptr_t>(q + i) & 15u) == 0);
  if (i + PER <= n) {
    uint4 cur = __ldg(reinterpret_cast<const uint4*>(p + i));
    for (;;) {
      const bool more = i + 2 * PER <= n;
      uint4 nxt = cur;
      if (more) nxt = __ldg(reinterpret_cast<const uint4*>(p + i + PER));
      const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
      uint32_t r[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int b = 0; b < 4 / W; ++b) {
          uint32_t o = (W == 1) ? ((w[k] >> (8 * b)) & 0xffu) : ((w[k] >> (16 * b)) & 0xffffu);
          uint32_t bin = decode_one<W, Ctx>(D, o, ctx, mytab, ctx_max, P.n_ctx);
          if (W == 1) r[k] |= bin << (8 * b);
          else q[i + k * 2 + b] = (uint8_t)bin;
        }
      }
      if (W == 1) {
        if (out_vec) {
          *reinterpret_cast<uint4*>(q + i) = make_uint4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) q[i + k] = (uint8_t)(r[k >> 2] >> (8 * (k & 3)));
        }
      }
      i += PER;
      if (!more) break;
      cur = nxt;
    }
  }
  for (; i < n; ++i) q[i] = (uint8_t)decode_one<W, Ctx>(D, p[i], ctx, mytab, ctx_max, P.n_ctx);

  if (P.finish_ok) P.finish_ok[s] = (uint8_t)dec_finish(D);
}

// ---------------------------------------------------------------------------
// wide-window kernels (cabac_wide.cuh): the hot path for the u8 op format
// ---------------------------------------------------------------------------
// One WARP per tile of 32 streams, NW warps per CTA (NW chosen by the host so that the tiles
// spread evenly over the SMs: 65,536 streams = 2,048 tiles = 14 warps on each of 147 SMs).
// Shared memory per CTA:
//   tab  [129][32] uint2   fused state rows, replicated per lane: row st of lane l at
//                          tab[st*32 + l] -> the 32 lanes always hit 32 distinct bank pairs
//   ctx  [NW][n_ctx+1][32] u32   context state bytes, one word per (warp, context, lane);
//                          slot n_ctx is the bypass slot (state 0x80)
// Per lane and 16 ops: one 16-byte op load (next block prefetched), 16 branch-free bin steps,
// 4 word emissions (encode) or refills (decode), one 16-byte store of the bins (decode).
__global__ void __launch_bounds__(WIDE_MAX_WARPS * 32) k_encode_ops_wide(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t s;
  WCtx ctx;
  WTab tab;
  uint32_t n_ctx, vmask;
  if (!wide_setup(P.n_streams, P.n_ctx, P.ctx_init, P.per_stream_init, smem, s, ctx, tab, n_ctx, &vmask)) return;
  const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
  const uint8_t* p = reinterpret_cast<const uint8_t*>(P.ops) + o0;
  const uint64_t n = o1 - o0;

  EncWide E;
  const uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  encw_start(E, P.slab + (uint64_t)s * P.slab_stride, cap);

  // head: general path up to the first 16-byte boundary of this lane's op array
  uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
  if (head > n) head = n;
  for (uint64_t i = 0; i < head; ++i) encw_general(E, p[i], ctx, tab, n_ctx);
  p += head;
  const uint64_t nblk = (n - head) >> 4;
  const uint32_t tail = (uint32_t)((n - head) & 15u);
  // body: 16 ops per load, next block prefetched while this one is coded.  In a full warp the
  // blocks every lane has are walked in lockstep (LOCK): the lazy emission votes across the warp
  // and a block with a terminate op sends the whole warp down the general path.  Lanes with
  // longer streams -- and warps with idle lanes -- go on without votes.
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
        for (int k = 0; k < 16; ++k) encw_general(E, p[k], ctx, tab, n_ctx);
      } else {
        encw_block16<LOCK>(E, w, cw, ctx, tab, n_ctx);
      }
      cur = nxt;
      p += 16;
    };
    if (vmask == 0xffffffffu)
      for (; b < common; ++b) block(std::true_type{});
    for (; b < nblk; ++b) block(std::false_type{});
  }
  for (uint32_t i = 0; i < tail; ++i) encw_general(E, p[i], ctx, tab, n_ctx);

  const uint32_t len = encw_finish(E);
  P.lengths[s] = len;
  if (len > cap && P.overflow) atomicOr(P.overflow, 1u);
}

// ---------------------------------------------------------------------------
// k_encode_ops_wide with ONE hand-over per SM (round 2, last part; ISSCABAC_HANDOVER=0 switches it off).
// 2,048 tiles on 592 schedulers is 3.46 warps per scheduler: with 14 warps per CTA two schedulers of every SM hold 4
// warps and two hold 3 (warp w issues on scheduler w % 4), the launch lasts as long as a 4-warp scheduler needs, and the
// 3-warp ones idle at its end.  Here the CTA has two more warps, NW and NW + 1, on the light schedulers 2 and 3.  They
// sleep on a named barrier.  Warps NW - 2 and NW - 1 (the fourth warps of schedulers 0 and 1) code the first HALF of
// their tile's lockstep blocks, park the lane state (window, range, pending word, op pointer) in shared memory, arrive
// at the barrier and exit; the sleeping warp adopts state and context block (the tokens are lane-relative, so they stay
// valid) and codes the rest.  In the throughput-bound picture every scheduler then carries 3.5 tiles instead of 4 / 3.
// Only for full tiles of streams with >= 64 common blocks; otherwise the giver says "no hand-over" and codes on.
// ---------------------------------------------------------------------------
#if !WIDE_CTX_ROWS
constexpr uint32_t HO_WORDS = 12;                                    // per lane: W (2), range, n, pend, wp, slot (2), p (2), blocks left, tail
constexpr uint32_t HO_SLOT_BYTES = HO_WORDS * 32 * 4 + 16;           // + the flag word: 0 = no hand-over, else lockstep blocks left + 1
constexpr uint32_t HO_BYTES = 2 * HO_SLOT_BYTES;

__global__ void __launch_bounds__(WIDE_MAX_WARPS * 32) k_encode_ops_wide_ho(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t nwarps = blockDim.x >> 5, home = nwarps - 2u;
  const uint32_t warp = threadIdx.x >> 5, lane = cb_keep32(threadIdx.x & 31);
  const bool receiver = warp >= home;                     // warp home adopts warp home - 2, warp home + 1 adopts warp home - 1
  const uint32_t ewarp = receiver ? warp - 2u : warp;     // the warp whose tile and context block this warp works on
  const bool giver = !receiver && warp + 2u >= home;
  const uint32_t xid = ewarp - (home - 2u);               // exchange slot (givers and receivers only)
  // set-up (wide_setup with `home` warps of tiles per CTA; the receivers initialise nothing)
  WRow* t = reinterpret_cast<WRow*>(smem);
  const uint32_t tab0 = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = threadIdx.x; i < kNumRows * WIDE_COLS; i += blockDim.x) {
    WRow r = c_wide_rows.r[i / WIDE_COLS];
    const uint32_t col = tab0 + (i % WIDE_COLS) * (uint32_t)sizeof(WRow);
    r.next_mps = col + r.next_mps * WIDE_ROW_STRIDE;
    r.next_lps = col + r.next_lps * WIDE_ROW_STRIDE;
    t[i] = r;
  }
  const uint32_t n_ctx = cb_keep32(P.n_ctx);
  const uint32_t s = (blockIdx.x * home + ewarp) * 32 + lane;
  const bool valid = s < P.n_streams;
  WTab tab;
  WCtx ctx;
  tab.base = cb_keep32(tab0 + (lane % WIDE_COLS) * (uint32_t)sizeof(WRow));
  ctx.p = reinterpret_cast<uint32_t*>(smem + WIDE_TAB_BYTES) + cb_keep32(ewarp * (n_ctx + 1) * 32 + lane);
  if (!receiver) {
    const uint8_t* init = P.ctx_init + (P.per_stream_init && valid ? (uint64_t)s * n_ctx : 0);
    for (uint32_t c = 0; c < n_ctx; ++c) ctx.store(c, tab.token(init[c] & 127u));
    ctx.store(n_ctx, tab.token(kEpState));
  }
  uint32_t* xch = reinterpret_cast<uint32_t*>(smem + WIDE_TAB_BYTES + (size_t)home * (n_ctx + 1) * WIDE_CTX_STRIDE + (size_t)(giver || receiver ? xid : 0u) * HO_SLOT_BYTES);
  uint32_t* flag = xch + HO_WORDS * 32;
  __syncthreads();
  const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
  const uint32_t bar_id = 1u + xid;

  EncWide E;
  const uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  const uint8_t* p = nullptr;
  uint64_t nblk = 0;
  uint32_t tail = 0, common = 0, hand_at = 0xffffffffu;
  if (receiver) {
    // every lane waits (the barrier counts threads); then the giver's verdict
    asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory");
    const uint32_t f = *reinterpret_cast<volatile uint32_t*>(flag);
    if (f == 0u) return;
    volatile uint32_t* x = xch + lane;
    E.W = (uint64_t)x[0 * 32] | ((uint64_t)x[1 * 32] << 32);
    E.range = x[2 * 32];
    E.n = (int32_t)x[3 * 32];
    E.pend = x[4 * 32];
    E.wp = x[5 * 32];
    E.cap_words = cap >> 2;
    E.slot = cb_keep(reinterpret_cast<uint32_t*>((uintptr_t)x[6 * 32] | ((uintptr_t)x[7 * 32] << 32)));
    p = reinterpret_cast<const uint8_t*>((uintptr_t)x[8 * 32] | ((uintptr_t)x[9 * 32] << 32));
    nblk = x[10 * 32];
    tail = x[11 * 32];
    common = f - 1u;
  } else {
    if (giver && vmask != 0xffffffffu) {     // not a full tile: no hand-over (all 32 lanes arrive, then the idle ones leave)
      if (lane == 0) *reinterpret_cast<volatile uint32_t*>(flag) = 0u;
      __syncwarp();
      __threadfence_block();
      asm volatile("bar.arrive %0, 64;" :: "r"(bar_id) : "memory");
    }
    if (!valid) return;
    const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
    p = reinterpret_cast<const uint8_t*>(P.ops) + o0;
    const uint64_t n = o1 - o0;
    encw_start(E, P.slab + (uint64_t)s * P.slab_stride, cap);
    uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
    if (head > n) head = n;
    for (uint64_t i = 0; i < head; ++i) encw_general(E, p[i], ctx, tab, n_ctx);
    p += head;
    nblk = (n - head) >> 4;
    tail = (uint32_t)((n - head) & 15u);
    common = __reduce_min_sync(vmask, (uint32_t)(nblk > 0xffffffffull ? 0xffffffffull : nblk));
    if (giver && vmask == 0xffffffffu) {
      const uint32_t worth = __reduce_max_sync(0xffffffffu, (uint32_t)(nblk > 0xfffffff0ull ? 1u : 0u));   // block counts that do not fit a word: no hand-over
      if (common >= 64u && !worth) {
        hand_at = (uint32_t)(((uint64_t)common * P.ho_eighths) >> 3);
      } else {
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(flag) = 0u;
        __syncwarp();
        __threadfence_block();
        asm volatile("bar.arrive %0, 64;" :: "r"(bar_id) : "memory");
      }
    }
  }
  const bool lockstep = receiver || vmask == 0xffffffffu;
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
        for (int k = 0; k < 16; ++k) encw_general(E, p[k], ctx, tab, n_ctx);
      } else {
        encw_block16<LOCK>(E, w, cw, ctx, tab, n_ctx);
      }
      cur = nxt;
      p += 16;
    };
    if (lockstep) {
      const uint32_t stop = hand_at < common ? hand_at : common;
      for (; b < stop; ++b) block(std::true_type{});
      if (hand_at != 0xffffffffu) {
        // the first half is coded: park the lane state, tell the sleeping warp, leave
        volatile uint32_t* x = xch + lane;
        x[0 * 32] = (uint32_t)E.W;
        x[1 * 32] = (uint32_t)(E.W >> 32);
        x[2 * 32] = E.range;
        x[3 * 32] = (uint32_t)E.n;
        x[4 * 32] = E.pend;
        x[5 * 32] = E.wp;
        x[6 * 32] = (uint32_t)reinterpret_cast<uintptr_t>(E.slot);
        x[7 * 32] = (uint32_t)(reinterpret_cast<uintptr_t>(E.slot) >> 32);
        x[8 * 32] = (uint32_t)reinterpret_cast<uintptr_t>(p);
        x[9 * 32] = (uint32_t)(reinterpret_cast<uintptr_t>(p) >> 32);
        x[10 * 32] = (uint32_t)(nblk - b);
        x[11 * 32] = tail;
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(flag) = common - hand_at + 1u;
        __syncwarp();
        __threadfence_block();
        asm volatile("bar.arrive %0, 64;" :: "r"(bar_id) : "memory");
        return;
      }
    }
    for (; b < nblk; ++b) block(std::false_type{});
  } else if (hand_at != 0xffffffffu) {
    // (cannot happen: hand_at is set only with common >= 64 blocks)
  }
  for (uint32_t i = 0; i < tail; ++i) encw_general(E, p[i], ctx, tab, n_ctx);

  const uint32_t len = encw_finish(E);
  P.lengths[s] = len;
  if (len > cap && P.overflow) atomicOr(P.overflow, 1u);
}
#endif

// ---------------------------------------------------------------------------
// The encoder as two warps per tile of 32 streams (used when there are few tiles per SM, see run_codec).
// The context side of a bin (token -> row -> next token) does not depend on the arithmetic side
// (range / low), so a PRODUCER warp walks the ops and the context states and hands, per bin, the row's
// four LPS sub-ranges plus {is LPS, is bypass, terminate, no-op} flags to a CONSUMER warp through a
// double-buffered ring in shared memory (one named barrier per pair and 16-op block); the consumer
// runs the window arithmetic and the emission.  Twice the warps per scheduler for the same streams, and two
// short dependent chains per bin instead of one long one: a lone tile codes a bin in 126 cycles instead of 161.
// ---------------------------------------------------------------------------
constexpr uint32_t SPLIT_GROUP = 32 * 16 + 32 * 4;   // 4 bins: [lane][4 x lps4] + [lane] flags word
constexpr uint32_t SPLIT_STAGE = 4 * SPLIT_GROUP;
constexpr uint32_t SPLIT_RING = 2 * SPLIT_STAGE;
constexpr uint32_t SPLIT_MAX_PAIRS = 15;
constexpr uint32_t SF_LPS = 1, SF_EP = 2, SF_TRM = 4, SF_NOP = 8;

template <int B>
__device__ __forceinline__ void split_bin(EncWide& E, uint32_t lps4, uint32_t f) {
  const uint32_t lps = cb_prmt(0, lps4, E.range >> 6);
  const uint32_t rmps = E.range - lps;
  const bool is_lps = (f & (SF_LPS << (8 * B))) != 0, is_ep = (f & (SF_EP << (8 * B))) != 0;
  const uint32_t x2 = is_ep ? E.range : 2u * rmps;   // plain select: as a predicated multiply-add it measured 7.58 against 7.30 ms here
  const uint32_t rsel = is_lps ? lps : rmps;
  const int nn = cb_renorm(rsel);
  const int ns = is_ep ? 1 : nn;
  uint64_t W = E.W;
  if (is_lps) W += x2;
  E.W = W << ns;
  E.range = is_ep ? E.range : (rsel << nn);
  E.n += ns;
}

__global__ void __launch_bounds__(2 * SPLIT_MAX_PAIRS * 32) k_encode_ops_split(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t np = blockDim.x >> 6;
  const uint32_t warp = threadIdx.x >> 5, lane = cb_keep32(threadIdx.x & 31);
  const bool producer = warp < np;
  const uint32_t pair = producer ? warp : warp - np;
  // table + context blocks (producers) as in wide_setup
  WRow* t = reinterpret_cast<WRow*>(smem);
  const uint32_t tab0 = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = threadIdx.x; i < kNumRows * WIDE_COLS; i += blockDim.x) {
    WRow r = c_wide_rows.r[i / WIDE_COLS];
    const uint32_t col = tab0 + (i % WIDE_COLS) * (uint32_t)sizeof(WRow);
    r.next_mps = col + r.next_mps * WIDE_ROW_STRIDE;
    r.next_lps = col + r.next_lps * WIDE_ROW_STRIDE;
    t[i] = r;
  }
  const uint32_t n_ctx = cb_keep32(P.n_ctx);
  const uint32_t s = (blockIdx.x * np + pair) * 32 + lane;
  const bool valid = s < P.n_streams;
  WTab tab;
  tab.base = cb_keep32(tab0 + (lane % WIDE_COLS) * (uint32_t)sizeof(WRow));
  WCtx ctx;
#if WIDE_CTX_ROWS
  ctx.base = cb_keep32(tab0 + (uint32_t)WIDE_TAB_BYTES + (pair * (n_ctx + 1) * 32 + lane) * WIDE_SLOT_BYTES);
  __syncthreads();
#else
  ctx.p = reinterpret_cast<uint32_t*>(smem + WIDE_TAB_BYTES) + cb_keep32(pair * (n_ctx + 1) * 32 + lane);
#endif
  if (producer) {
    const uint8_t* init = P.ctx_init + (P.per_stream_init && valid ? (uint64_t)s * n_ctx : 0);
    for (uint32_t c = 0; c < n_ctx; ++c) ctx.store(c, tab.token(init[c] & 127u));
    ctx.store(n_ctx, tab.token(kEpState));
  }
  __syncthreads();
  uint8_t* ring = smem + WIDE_TAB_BYTES + (size_t)np * (n_ctx + 1) * WIDE_CTX_STRIDE + (size_t)pair * SPLIT_RING;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring) + lane * 16u;   // this lane's lps4 quad of group 0, stage 0
  const uint32_t flag_s = (uint32_t)__cvta_generic_to_shared(ring) + 512u + lane * 4u;
  const uint32_t bar_id = pair + 1;

  // every stream is walked in the aligned 16-byte windows of its op array (positions outside the stream are no-ops),
  // so both warps of a pair know the number of blocks without talking to each other
  const uint64_t o0 = valid ? P.op_off[s] : 0, o1 = valid ? P.op_off[s + 1] : 0;
  const uintptr_t a_begin = reinterpret_cast<uintptr_t>(P.ops) + o0, a_end = reinterpret_cast<uintptr_t>(P.ops) + o1;
  const uintptr_t w0 = a_begin & ~(uintptr_t)15;
  const uint32_t T = o1 > o0 ? (uint32_t)((((a_end + 15) & ~(uintptr_t)15) - w0) >> 4) : 0u;
  const uint32_t Tmax = __reduce_max_sync(0xffffffffu, T);

  if (producer) {
    uint4 nxt = make_uint4(0, 0, 0, 0);
    // windows [bf0, bf1) lie inside the stream: one aligned 16-byte load each; the (at most two) others byte by byte
    const uint32_t bf0 = (a_begin & 15u) ? 1u : 0u, bf1 = T ? ((a_end & 15u) ? T - 1u : T) : 0u;
    auto full = [&](uint32_t b) { return b >= bf0 && b < bf1; };
    if (full(0)) nxt = __ldg(reinterpret_cast<const uint4*>(w0));
    uintptr_t wa = w0;
    for (uint32_t b = 0; b < Tmax; ++b, wa += 16) {
      const uint32_t st_off = (b & 1u) * SPLIT_STAGE;
      uint32_t w[4] = {nxt.x, nxt.y, nxt.z, nxt.w};
      uint32_t nopm = 0;                 // bit e: position e of the window holds no op of this stream
      if (!full(b)) {
        nopm = 0xffffu;
        w[0] = w[1] = w[2] = w[3] = 0;
        if (b < T) {
          for (uint32_t e = 0; e < 16; ++e) {
            const uintptr_t a = wa + e;
            if (a >= a_begin && a < a_end) {
              nopm &= ~(1u << e);
              w[e >> 2] |= (uint32_t)(*reinterpret_cast<const uint8_t*>(a)) << (8 * (e & 3));
            }
          }
        }
      }
      if (full(b + 1)) nxt = __ldg(reinterpret_cast<const uint4*>(wa + 16));
      const uint32_t cw[4] = {op_codes4(w[0]), op_codes4(w[1]), op_codes4(w[2]), op_codes4(w[3])};
      if (nopm == 0 && !block_has_trm(cw)) {
        // the common block: 16 context / bypass ops.  Per op: token, row, "is it the LPS" straight into the flag
        // byte of the op's position, next token; the bypass flags of four ops come from one SIMD-in-register test
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t codes = cw[g];
          uint32_t fl = 0;
          uint32_t l4[4];
#define SPLIT_OP(K)                                                                                   \
          {                                                                                           \
            const uint32_t code = cb_prmt(codes, 0, 0x4440u + K);                                     \
            const uint32_t c = code < n_ctx ? code : n_ctx;                                           \
            const WRow row = tab.row(ctx.load(c));                                                    \
            const uint32_t lb = cb_xor_and<(1u << (8 * K))>(row.mps4, w[g]);                          \
            ctx.store_sel(c, lb, row.next_lps, row.next_mps);                                         \
            fl |= lb | (code >= n_ctx ? (SF_EP << (8 * K)) : 0u);   /* bypass: 126 or any non-context code */ \
            l4[K] = row.lps4;                                                                         \
          }
          SPLIT_OP(0) SPLIT_OP(1) SPLIT_OP(2) SPLIT_OP(3)
#undef SPLIT_OP
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(ring_s + st_off + g * SPLIT_GROUP), "r"(l4[0]), "r"(l4[1]), "r"(l4[2]), "r"(l4[3]) : "memory");
          asm volatile("st.shared.u32 [%0], %1;" :: "r"(flag_s + st_off + g * SPLIT_GROUP), "r"(fl) : "memory");
        }
      } else {
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          uint32_t l4[4], fl = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t op = (w[g] >> (8 * k)) & 0xffu, code = op >> 1;
            l4[k] = 0;
            if ((nopm >> (4 * g + k)) & 1u) {
              fl |= SF_NOP << (8 * k);
            } else if (code == kOpTrmCode) {
              fl |= (SF_TRM | (op & 1u)) << (8 * k);
            } else {
              const uint32_t c = code < n_ctx ? code : n_ctx;
              const WRow row = tab.row(ctx.load(c));
              const uint32_t is_lps = (row.mps4 ^ op) & 1u;
              ctx.store(c, is_lps ? row.next_lps : row.next_mps);
              l4[k] = row.lps4;
              fl |= (is_lps | (code >= n_ctx ? SF_EP : 0u)) << (8 * k);
            }
          }
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(ring_s + st_off + g * SPLIT_GROUP), "r"(l4[0]), "r"(l4[1]), "r"(l4[2]), "r"(l4[3]) : "memory");
          asm volatile("st.shared.u32 [%0], %1;" :: "r"(flag_s + st_off + g * SPLIT_GROUP), "r"(fl) : "memory");
        }
      }
      asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory");
    }
    return;
  }

  // consumer
  EncWide E;
  const uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  encw_start(E, valid ? P.slab + (uint64_t)s * P.slab_stride : nullptr, valid ? cap : 0u);
  for (uint32_t b = 0; b < Tmax; ++b) {
    const uint32_t st_off = (b & 1u) * SPLIT_STAGE;
    asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory");
    // the whole block's sub-ranges and flags at once (eight loads in flight), and ONE vote on "some lane holds a terminate
    // op or a position outside its stream" per block instead of one per group: a vote + branch costs a lone pair of warps
    // as much as two bins' arithmetic (profiles/r2_tree_decoder_latency.txt, the same finding on this kernel's consumer)
    uint32_t l[4][4], fl[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(l[g][0]), "=r"(l[g][1]), "=r"(l[g][2]), "=r"(l[g][3]) : "r"(ring_s + st_off + g * SPLIT_GROUP) : "memory");
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(fl[g]) : "r"(flag_s + st_off + g * SPLIT_GROUP) : "memory");
    }
    if (__any_sync(0xffffffffu, ((fl[0] | fl[1] | fl[2] | fl[3]) & 0x0c0c0c0cu) != 0u)) {
      // one bin at a time
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        uint32_t l0, l1, l2, l3, f4;      // read again: this path is rare, the registers above are indexed at compile time only
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(l0), "=r"(l1), "=r"(l2), "=r"(l3) : "r"(ring_s + st_off + g * SPLIT_GROUP) : "memory");
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(f4) : "r"(flag_s + st_off + g * SPLIT_GROUP) : "memory");
        const uint32_t l4[4] = {l0, l1, l2, l3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t f = (f4 >> (8 * k)) & 0xffu;
          if (f & SF_NOP) continue;
          if (f & SF_TRM) encw_trm(E, f & 1u);
          else split_bin<0>(E, l4[k], f);
          encw_emit(E);
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        split_bin<0>(E, l[g][0], fl[g]);
        split_bin<1>(E, l[g][1], fl[g]);
        if (E.n >= kLazy) encw_emit(E);
        split_bin<2>(E, l[g][2], fl[g]);
        split_bin<3>(E, l[g][3], fl[g]);
        if (__any_sync(0xffffffffu, E.n >= kLazy)) encw_emit(E);
      }
    }
  }
  if (valid) {
    const uint32_t len = encw_finish(E);
    P.lengths[s] = len;
    if (len > cap && P.overflow) atomicOr(P.overflow, 1u);
  }
}

#ifndef WIDE_DEC_MINBLOCKS
#define WIDE_DEC_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(WIDE_MAX_WARPS * 32, WIDE_DEC_MINBLOCKS) k_decode_ops_wide(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t s;
  WCtx ctx;
  WTab tab;
  uint32_t n_ctx, vmask;
  if (!wide_setup(P.n_streams, P.n_ctx, P.ctx_init, P.per_stream_init, smem, s, ctx, tab, n_ctx, &vmask)) return;
  const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
  const uint8_t* p = reinterpret_cast<const uint8_t*>(P.ops) + o0;
  uint8_t* q = P.bins + o0;
  const uint64_t n = o1 - o0;
  const uint64_t b0 = P.byte_off[s], b1 = P.byte_off[s + 1];

  DecWide D;
#if CABAC_DEC_TMA
  {   // this lane's ring + mbarriers behind the context blocks (see run_codec for the size)
    const uint32_t nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)WIDE_TAB_BYTES + nwarps * (n_ctx + 1) * WIDE_CTX_STRIDE;
    D.ring = base + (warp * 32u + lane) * kTmaLaneStride;
    D.mbar = base + nwarps * 32u * kTmaLaneStride + (warp * 32u + lane) * 16u;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(D.mbar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(D.mbar + 8u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    D.gend = (reinterpret_cast<uint64_t>(P.bytes) + P.byte_off[P.n_streams] + 15ull) & ~15ull;
  }
#endif
  decw_start(D, P.bytes + b0, (uint32_t)(b1 - b0));

  uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
  if (head > n) head = n;
  for (uint64_t i = 0; i < head; ++i) q[i] = (uint8_t)decw_general(D, p[i], ctx, tab, n_ctx);
  p += head;
  q += head;
  const uint64_t nblk = (n - head) >> 4;
  const uint32_t tail = (uint32_t)((n - head) & 15u);
  const bool out_vec = (reinterpret_cast<uintptr_t>(q) & 15u) == 0;  // true whenever ops and bins share their alignment
  // lockstep blocks first, like the encoder
  const bool pf_ops = P.prefetch_ops != 0u;
  const uint32_t common = __reduce_min_sync(vmask, (uint32_t)(nblk > 0xffffffffull ? 0xffffffffull : nblk));
  if (nblk) {
    uint4 cur = __ldg(reinterpret_cast<const uint4*>(p));
    uint64_t b = 0;
    auto block = [&](auto lock) {
      constexpr bool LOCK = decltype(lock)::value;
      uint4 nxt = cur;
      if (b + 1 < nblk) nxt = __ldg(reinterpret_cast<const uint4*>(p + 16));
      // The op-block load and the window top-up's word share a load scoreboard: a top-up two groups into the block waited for
      // the block's own next-op load when that went to DRAM (ncu, lone tile: 78 + 52 + 14 cycles per block on the top-ups'
      // first instruction, 78 on the rotation's move).  Asking the op stream into L1 four blocks ahead makes that load an L1
      // hit.  With few tiles per SM only (run_codec): 8,192 streams 4.43 -> 4.21 ms, 16,384 4.47 -> 4.24, but 65,536 7.05 ->
      // 7.24 (the encoders gain nothing at any size).
      if (pf_ops && b + 6 < nblk) asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 16 + 64));
      const uint32_t cw[4] = {op_codes4(cur.x), op_codes4(cur.y), op_codes4(cur.z), op_codes4(cur.w)};
      if (cb_any<LOCK>(block_has_trm(cw))) {
        if (LOCK && CABAC_REFILL_P == 3 && CABAC_REFILL_FIRST) decw_refill(D);
        for (int k = 0; k < 16; ++k) q[k] = (uint8_t)decw_general(D, p[k], ctx, tab, n_ctx);
      } else {
        uint32_t r[4];
        decw_block16<LOCK>(D, cw, r, ctx, tab, n_ctx);
        if (out_vec) {
          *reinterpret_cast<uint4*>(q) = make_uint4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) q[k] = (uint8_t)(r[k >> 2] >> (8 * (k & 3)));
        }
      }
      cur = nxt;
      p += 16;
      q += 16;
    };
    if (vmask == 0xffffffffu) {
      for (; b < common; ++b) block(std::true_type{});
      if (CABAC_REFILL_P == 3 && CABAC_REFILL_FIRST) decw_refill(D);
    }
    for (; b < nblk; ++b) block(std::false_type{});
  }
  for (uint32_t i = 0; i < tail; ++i) q[i] = (uint8_t)decw_general(D, p[i], ctx, tab, n_ctx);

  if (P.finish_ok) P.finish_ok[s] = (uint8_t)decw_finish(D);
}

// ---------------------------------------------------------------------------
// k_decode_ops_wide with the encoder's hand-over (k_encode_ops_wide_ho: one tile per 4-warp scheduler moves, part-decoded,
// to a warp asleep on a 3-warp scheduler).  The lane state is the window (DecWide: 15 words with the two payload words in
// flight) + the op / bin pointers and counts: 20 words.
// ---------------------------------------------------------------------------
#if !WIDE_CTX_ROWS && !CABAC_DEC_TMA
constexpr uint32_t HOD_WORDS = 20;
constexpr uint32_t HOD_SLOT_BYTES = HOD_WORDS * 32 * 4 + 16;
constexpr uint32_t HOD_BYTES = 2 * HOD_SLOT_BYTES;

__global__ void __launch_bounds__(WIDE_MAX_WARPS * 32, WIDE_DEC_MINBLOCKS) k_decode_ops_wide_ho(CodecParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t nwarps = blockDim.x >> 5, home = nwarps - 2u;
  const uint32_t warp = threadIdx.x >> 5, lane = cb_keep32(threadIdx.x & 31);
  const bool receiver = warp >= home;
  const uint32_t ewarp = receiver ? warp - 2u : warp;
  const bool giver = !receiver && warp + 2u >= home;
  const uint32_t xid = ewarp - (home - 2u);
  WRow* t = reinterpret_cast<WRow*>(smem);
  const uint32_t tab0 = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = threadIdx.x; i < kNumRows * WIDE_COLS; i += blockDim.x) {
    WRow r = c_wide_rows.r[i / WIDE_COLS];
    const uint32_t col = tab0 + (i % WIDE_COLS) * (uint32_t)sizeof(WRow);
    r.next_mps = col + r.next_mps * WIDE_ROW_STRIDE;
    r.next_lps = col + r.next_lps * WIDE_ROW_STRIDE;
    t[i] = r;
  }
  const uint32_t n_ctx = cb_keep32(P.n_ctx);
  const uint32_t s = (blockIdx.x * home + ewarp) * 32 + lane;
  const bool valid = s < P.n_streams;
  WTab tab;
  WCtx ctx;
  tab.base = cb_keep32(tab0 + (lane % WIDE_COLS) * (uint32_t)sizeof(WRow));
  ctx.p = reinterpret_cast<uint32_t*>(smem + WIDE_TAB_BYTES) + cb_keep32(ewarp * (n_ctx + 1) * 32 + lane);
  if (!receiver) {
    const uint8_t* init = P.ctx_init + (P.per_stream_init && valid ? (uint64_t)s * n_ctx : 0);
    for (uint32_t c = 0; c < n_ctx; ++c) ctx.store(c, tab.token(init[c] & 127u));
    ctx.store(n_ctx, tab.token(kEpState));
  }
  uint32_t* xch = reinterpret_cast<uint32_t*>(smem + WIDE_TAB_BYTES + (size_t)home * (n_ctx + 1) * WIDE_CTX_STRIDE + (size_t)(giver || receiver ? xid : 0u) * HOD_SLOT_BYTES);
  uint32_t* flag = xch + HOD_WORDS * 32;
  __syncthreads();
  const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
  const uint32_t bar_id = 1u + xid;

  DecWide D;
  const uint8_t* p = nullptr;
  uint8_t* q = nullptr;
  uint64_t nblk = 0;
  uint32_t tail = 0, common = 0, hand_at = 0xffffffffu;
  if (receiver) {
    asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory");
    const uint32_t f = *reinterpret_cast<volatile uint32_t*>(flag);
    if (f == 0u) return;
    volatile uint32_t* x = xch + lane;
    D.lo = x[0 * 32]; D.hi = x[1 * 32]; D.range = x[2 * 32]; D.f = (int32_t)x[3 * 32]; D.p = x[4 * 32]; D.len = x[5 * 32];
    D.cur = x[6 * 32]; D.nxt = x[7 * 32]; D.sel = x[8 * 32]; D.widx = x[9 * 32]; D.wcnt = x[10 * 32];
    D.wbase = cb_keep(reinterpret_cast<const uint32_t*>((uintptr_t)x[11 * 32] | ((uintptr_t)x[12 * 32] << 32)));
    D.in = reinterpret_cast<const uint8_t*>((uintptr_t)x[13 * 32] | ((uintptr_t)x[14 * 32] << 32));
    p = reinterpret_cast<const uint8_t*>((uintptr_t)x[15 * 32] | ((uintptr_t)x[16 * 32] << 32));
    q = reinterpret_cast<uint8_t*>((uintptr_t)x[17 * 32] | ((uintptr_t)x[18 * 32] << 32));
    nblk = x[19 * 32] >> 4;
    tail = x[19 * 32] & 15u;
    common = f - 1u;
  } else {
    if (giver && vmask != 0xffffffffu) {
      if (lane == 0) *reinterpret_cast<volatile uint32_t*>(flag) = 0u;
      __syncwarp();
      __threadfence_block();
      asm volatile("bar.arrive %0, 64;" :: "r"(bar_id) : "memory");
    }
    if (!valid) return;
    const uint64_t o0 = P.op_off[s], o1 = P.op_off[s + 1];
    p = reinterpret_cast<const uint8_t*>(P.ops) + o0;
    q = P.bins + o0;
    const uint64_t n = o1 - o0;
    const uint64_t b0 = P.byte_off[s], b1 = P.byte_off[s + 1];
    decw_start(D, P.bytes + b0, (uint32_t)(b1 - b0));
    uint64_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
    if (head > n) head = n;
    for (uint64_t i = 0; i < head; ++i) q[i] = (uint8_t)decw_general(D, p[i], ctx, tab, n_ctx);
    p += head;
    q += head;
    nblk = (n - head) >> 4;
    tail = (uint32_t)((n - head) & 15u);
    common = __reduce_min_sync(vmask, (uint32_t)(nblk > 0xffffffffull ? 0xffffffffull : nblk));
    if (giver && vmask == 0xffffffffu) {
      const uint32_t huge = __reduce_max_sync(0xffffffffu, (uint32_t)(nblk > 0x0ffffff0ull ? 1u : 0u));   // (blocks left, tail) share a word
      if (common >= 64u && !huge) {
        hand_at = (uint32_t)(((uint64_t)common * P.ho_eighths) >> 3);
      } else {
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(flag) = 0u;
        __syncwarp();
        __threadfence_block();
        asm volatile("bar.arrive %0, 64;" :: "r"(bar_id) : "memory");
      }
    }
  }
  const bool out_vec = (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
  const bool lockstep = receiver || vmask == 0xffffffffu;
  if (nblk) {
    uint4 cur = __ldg(reinterpret_cast<const uint4*>(p));
    uint64_t b = 0;
    auto block = [&](auto lock) {
      constexpr bool LOCK = decltype(lock)::value;
      uint4 nxt = cur;
      if (b + 1 < nblk) nxt = __ldg(reinterpret_cast<const uint4*>(p + 16));
      const uint32_t cw[4] = {op_codes4(cur.x), op_codes4(cur.y), op_codes4(cur.z), op_codes4(cur.w)};
      if (cb_any<LOCK>(block_has_trm(cw))) {
        if (LOCK && CABAC_REFILL_P == 3 && CABAC_REFILL_FIRST) decw_refill(D);
        for (int k = 0; k < 16; ++k) q[k] = (uint8_t)decw_general(D, p[k], ctx, tab, n_ctx);
      } else {
        uint32_t r[4];
        decw_block16<LOCK>(D, cw, r, ctx, tab, n_ctx);
        if (out_vec) {
          *reinterpret_cast<uint4*>(q) = make_uint4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) q[k] = (uint8_t)(r[k >> 2] >> (8 * (k & 3)));
        }
      }
      cur = nxt;
      p += 16;
      q += 16;
    };
    if (lockstep) {
      const uint32_t stop = hand_at < common ? hand_at : common;
      for (; b < stop; ++b) block(std::true_type{});
      if (hand_at != 0xffffffffu) {
        volatile uint32_t* x = xch + lane;
        x[0 * 32] = D.lo; x[1 * 32] = D.hi; x[2 * 32] = D.range; x[3 * 32] = (uint32_t)D.f; x[4 * 32] = D.p; x[5 * 32] = D.len;
        x[6 * 32] = D.cur; x[7 * 32] = D.nxt; x[8 * 32] = D.sel; x[9 * 32] = D.widx; x[10 * 32] = D.wcnt;
        x[11 * 32] = (uint32_t)reinterpret_cast<uintptr_t>(D.wbase);
        x[12 * 32] = (uint32_t)(reinterpret_cast<uintptr_t>(D.wbase) >> 32);
        x[13 * 32] = (uint32_t)reinterpret_cast<uintptr_t>(D.in);
        x[14 * 32] = (uint32_t)(reinterpret_cast<uintptr_t>(D.in) >> 32);
        x[15 * 32] = (uint32_t)reinterpret_cast<uintptr_t>(p);
        x[16 * 32] = (uint32_t)(reinterpret_cast<uintptr_t>(p) >> 32);
        x[17 * 32] = (uint32_t)reinterpret_cast<uintptr_t>(q);
        x[18 * 32] = (uint32_t)(reinterpret_cast<uintptr_t>(q) >> 32);
        x[19 * 32] = ((uint32_t)(nblk - b) << 4) | tail;
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(flag) = common - hand_at + 1u;
        __syncwarp();
        __threadfence_block();
        asm volatile("bar.arrive %0, 64;" :: "r"(bar_id) : "memory");
        return;
      }
      if (CABAC_REFILL_P == 3 && CABAC_REFILL_FIRST) decw_refill(D);
    }
    for (; b < nblk; ++b) block(std::false_type{});
  }
  for (uint32_t i = 0; i < tail; ++i) q[i] = (uint8_t)decw_general(D, p[i], ctx, tab, n_ctx);

  if (P.finish_ok) P.finish_ok[s] = (uint8_t)decw_finish(D);
}
#endif

// ---------------------------------------------------------------------------
// device-wide exclusive scan (u32 -> u64), single pass, decoupled look-back
// ---------------------------------------------------------------------------
// Tiles of SCAN_TILE elements; tile order is taken from an atomic ticket so a tile only
// ever waits for tiles that already started.  Tile descriptor: bit 63 = inclusive prefix
// available, bit 62 = tile aggregate available, low 62 bits = value.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long FLAG_P = 1ull << 63, FLAG_A = 1ull << 62, VAL_MASK = (1ull << 62) - 1;

__global__ void k_scan_init(unsigned long long* desc, uint32_t n_tiles, uint32_t* ticket) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_tiles) desc[i] = 0;
  if (i == 0) *ticket = 0;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_u32_u64(const uint32_t* in, uint64_t* out, uint64_t n,
                                                                unsigned long long* desc, uint32_t* ticket) {
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
  __shared__ unsigned long long s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t base = (uint64_t)tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  unsigned long long local = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    local += v[k];
  }
  // warp inclusive scan of the per-thread sums
  unsigned long long inc = local;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  unsigned long long warp_off = 0, tile_sum = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    if (w < wid) warp_off += s_warp[w];
    tile_sum += s_warp[w];
  }
  // publish the aggregate, then look back (warp 0)
  if (wid == 0) {
    if (lane == 0) {
      unsigned long long d = (tile == 0 ? FLAG_P : FLAG_A) | tile_sum;
      atomicExch(&desc[tile], d);
    }
    unsigned long long prefix = 0;
    if (tile > 0) {
      int64_t look = (int64_t)tile - 1 - lane;
      for (;;) {
        unsigned long long d = 0;
        bool have;
        do {
          d = look >= 0 ? atomicAdd(&desc[look >= 0 ? look : 0], 0ull) : FLAG_P;
          have = (d & (FLAG_P | FLAG_A)) != 0;
        } while (!__all_sync(0xffffffffu, have));
        // first lane (closest predecessor first) holding an inclusive prefix ends the walk
        unsigned has_p = __ballot_sync(0xffffffffu, (d & FLAG_P) != 0);
        int stop = has_p ? __ffs(has_p) - 1 : 32;
        unsigned long long contrib = (lane <= stop && look >= 0) ? (d & VAL_MASK) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_down_sync(0xffffffffu, contrib, o);
        contrib = __shfl_sync(0xffffffffu, contrib, 0);
        prefix += contrib;
        if (has_p) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(&desc[tile], FLAG_P | (prefix + tile_sum));
    }
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  unsigned long long run = s_prefix + warp_off + (inc - local);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = run;  // total
}

// One warp per stream: copy lengths[s] bytes from the 16-byte aligned slab row to an
// arbitrarily aligned payload position.  Destination-aligned 32-bit words are built from
// two aligned source words with a funnel shift, so every store is a full, coalesced word.
constexpr uint64_t kCompactShortRow = 2048;   // slab stride up to which a row is copied by 8 lanes
// G lanes per stream: a whole warp for long rows; 8 lanes for short ones (C4: 2^20 rows of ~270 bytes -- with a warp per row the
// launch is bound by how many rows are in flight, three dependent loads each, not by bandwidth: 0.48 -> see DESIGN.md 3.2)
template <int G>
__global__ void __launch_bounds__(256) k_compact_copy(const uint8_t* slab, uint64_t stride, const uint32_t* lengths,
                                                       const uint64_t* off, uint8_t* payload, uint64_t cap,
                                                       uint32_t n_streams, uint32_t* overflow) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / G, lane = threadIdx.x % G;      // "warp" = the row's lane group
  if (warp >= n_streams) return;
  uint32_t len = lengths[warp];
  if (len > stride) len = (uint32_t)stride;
  const uint64_t d0 = off[warp];
  if (d0 + len > cap) {
    if (lane == 0 && overflow) atomicOr(overflow, 2u);
    return;
  }
  const uint8_t* src = slab + (uint64_t)warp * stride;
  uint8_t* dst = payload + d0;
  // head bytes up to the first aligned destination word
  uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u);
  if (head > len) head = len;
  if (lane < head) dst[lane] = src[lane];
  const uint32_t nwords = (len - head) >> 2;
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(src);
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
  const uint32_t sh = 8 * head;  // source byte offset of destination word 0 is `head` (0..3)
  // four independent word pairs in flight per lane: the copy is bound by the loads a warp keeps outstanding
  uint32_t w = lane;
  for (; w + 3 * G < nwords; w += 4 * G) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo[k] = sw[w + G * k];
      hi[k] = sh ? sw[w + G * k + 1] : 0u;   // sw[.. + 1] stays inside the slab row: head > 0 => bytes remain
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) dw[w + G * k] = sh ? __funnelshift_r(lo[k], hi[k], sh) : lo[k];
  }
  for (; w < nwords; w += G) {
    uint32_t lo = sw[w], hi = sh ? sw[w + 1] : 0u;
    dw[w] = sh ? __funnelshift_r(lo, hi, sh) : lo;
  }
  const uint32_t done = head + (nwords << 2);
  if (lane < len - done) dst[done + lane] = src[done + lane];
}

// bins one per byte -> one per bit: thread t packs bytes [16t, 16t + 16) of the range into two bytes
__global__ void __launch_bounds__(256) k_pack_bins(const uint8_t* bins, uint64_t n, uint8_t* out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i0 = t * 16;
  if (i0 >= n) return;
  uint32_t w[4] = {0, 0, 0, 0};
  if (i0 + 16 <= n && (reinterpret_cast<uintptr_t>(bins + i0) & 15u) == 0) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(bins + i0));
    w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
  } else {
    for (uint32_t k = 0; k < 16 && i0 + k < n; ++k) w[k >> 2] |= (uint32_t)bins[i0 + k] << (8 * (k & 3));
  }
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) bits |= ((((w[k] & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * k);   // byte j bit 0 -> bit j
  const uint64_t left = n - i0;
  out[2 * t] = (uint8_t)bits;
  if (left > 8) out[2 * t + 1] = (uint8_t)(bits >> 8);
}

}  // namespace

// ===========================================================================
// host side of the device-pointer C ABI
// ===========================================================================
namespace isscabac_internal {

thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return ISSCABAC_ERR_CUDA;
}

// per-device attribute caches (a host thread may switch devices between calls)
constexpr int kMaxDevs = 64;
static int current_dev() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevs ? dev : -1;
}
int sm_count() {
  static std::atomic<int> n[kMaxDevs] = {};
  const int dev = current_dev();
  if (dev < 0) return 1;
  int v = n[dev].load(std::memory_order_relaxed);
  if (!v && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) n[dev].store(v, std::memory_order_relaxed);
  return v > 0 ? v : 1;
}

// A zeroed u32 work counter for persistent kernels: slots of a small per-device ring that is
// allocated once (no stream-ordered allocation on the launch path); the slot is cleared on the
// caller's stream right before use.  512 launches may be in flight per device.
int work_counter(cudaStream_t st, uint32_t** out) {
  constexpr int kSlots = 512;
  static uint32_t* ring[kMaxDevs] = {};
  static std::atomic<unsigned> next[kMaxDevs] = {};
  static std::mutex ring_mutex;
  const int dev = current_dev();
  if (dev < 0) { set_error("no current CUDA device (or device index out of range)"); return ISSCABAC_ERR_CUDA; }
  uint32_t* base;
  {
    std::lock_guard<std::mutex> lock(ring_mutex);
    if (!ring[dev]) {
      cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ring[dev]), kSlots * sizeof(uint32_t));
      if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(work counters)");
    }
    base = ring[dev];
  }
  uint32_t* slot = base + (next[dev].fetch_add(1u, std::memory_order_relaxed) % kSlots);
  cudaError_t e = cudaMemsetAsync(slot, 0, sizeof(uint32_t), st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(work counter)");
  *out = slot;
  return ISSCABAC_OK;
}

// The launch paths that need stream-ordered scratch use the device's default memory pool; by
// default that pool hands its memory back to the driver at every synchronisation, which makes
// each later cudaMallocAsync a slow driver call.  Keep freed blocks cached instead (once per device).
int keep_pool_cached() {
  static std::atomic<bool> done[kMaxDevs] = {};
  const int dev = current_dev();
  if (dev < 0) return cuda_fail(cudaErrorInvalidDevice, "cudaGetDevice");
  if (!done[dev].load(std::memory_order_acquire)) {
    cudaMemPool_t pool;
    cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetDefaultMemPool");
    uint64_t thr = ~0ull;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemPoolSetAttribute");
    done[dev].store(true, std::memory_order_release);
  }
  return ISSCABAC_OK;
}

size_t smem_limit() {
  static std::atomic<size_t> lim[kMaxDevs] = {};
  const int dev = current_dev();
  if (dev < 0) return 0;
  size_t v = lim[dev].load(std::memory_order_relaxed);
  if (!v) {
    int a = 0;
    if (cudaDeviceGetAttribute(&a, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess) {
      v = (size_t)a;
      lim[dev].store(v, std::memory_order_relaxed);
    }
  }
  return v;
}

}  // namespace isscabac_internal
using namespace isscabac_internal;

namespace {

template <class K>
int launch_codec(K kernel, const CodecParams& P, size_t smem, cudaStream_t st, const char* name) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  }
  uint32_t grid = (P.n_streams + NT - 1) / NT;
  kernel<<<grid, NT, smem, st>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, name);
  return ISSCABAC_OK;
}

// picks the context-storage policy; allocates the global scratch when needed
#ifndef LAT_TILES_PER_SM
#define LAT_TILES_PER_SM 0      // 0: the latency kernels run only when forced (see run_codec)
#endif
constexpr uint32_t kLatTilesPerSm = LAT_TILES_PER_SM;
#ifndef PF_OPS_TILES_PER_SM
#define PF_OPS_TILES_PER_SM 8
#endif
constexpr uint32_t kPrefetchOpsTilesPerSm = PF_OPS_TILES_PER_SM;   // the wide decoder's op-stream prefetch (k_decode_ops_wide)

// the hand-over encoder (k_encode_ops_wide_ho) runs when the wide geometry is one CTA per SM with 4k + 2 warps
inline bool handover_geometry(uint32_t nw, uint32_t grid, size_t wsmem, size_t lim) {
#if WIDE_CTX_ROWS
  return false;
#else
  const char* ho_env = getenv("ISSCABAC_HANDOVER");
  return !(ho_env && ho_env[0] == '0') && nw % 4u == 2u && nw + 2u <= (uint32_t)WIDE_MAX_WARPS && grid <= (uint32_t)sm_count() &&
         wsmem + HO_BYTES <= lim;
#endif
}
inline bool encoder_split_on(uint32_t n_streams) {
  const char* split_env = getenv("ISSCABAC_ENC_SPLIT");
  const uint32_t split_tiles = (n_streams + 31) / 32;
  return split_env && (split_env[0] == '0' || split_env[0] == '1') ? split_env[0] == '1' : split_tiles <= 11u * (uint32_t)sm_count();
}

// the hand-over decoder: the same geometry, and only where the plain kernel runs without its op-stream prefetch (more than
// kPrefetchOpsTilesPerSm tiles per SM: the hand-over kernel has none)
inline bool decode_handover_geometry(uint32_t n_streams, uint32_t nw, uint32_t grid, size_t wsmem, size_t lim) {
#if WIDE_CTX_ROWS || CABAC_DEC_TMA
  return false;
#else
  return (n_streams + 31) / 32 > kPrefetchOpsTilesPerSm * (uint32_t)sm_count() && !getenv("ISSCABAC_HANDOVER_ENC_ONLY") &&
         handover_geometry(nw, grid, wsmem + HOD_BYTES - HO_BYTES, lim);
#endif
}

template <bool ENC>
int run_codec(CodecParams P, int op_width, cudaStream_t st) {
  if (op_width != 1 && op_width != 2) { set_error("op_width must be 1 or 2"); return ISSCABAC_ERR_INVALID; }
  if (P.n_ctx > ISSCABAC_MAX_CTX) { set_error("n_ctx > %u", ISSCABAC_MAX_CTX); return ISSCABAC_ERR_INVALID; }
  if (P.n_streams == 0) return ISSCABAC_OK;
  const size_t tab_bytes = TAB_WORDS * sizeof(uint2);
  const size_t smem_ctx = tab_bytes + (size_t)(P.n_ctx ? P.n_ctx : 1) * NT * 4;   // n_ctx == 0: slot 0 is still loaded (value unused)
  const size_t lim = smem_limit();
  if (!lim) return cuda_fail(cudaErrorNoDevice, "no CUDA device");
  // hot path: u8 ops, context block of one warp fits shared memory next to the table
  uint32_t nw, grid;
  size_t wsmem;
  // Encoder with few tiles per SM: the two-warp formulation (k_encode_ops_split).  Measured on B200 over 65,536-bin
  // streams: 4.20 against 5.36 ms up to 16 K streams, 5.11 / 5.84 at 32 K, 6.28 / 6.49 at 48 K, equal from 56 K on, where
  // both are bound by the ALU pipe and the ring traffic only adds to it.  ISSCABAC_ENC_SPLIT=0 / 1 forces the choice.
  const uint32_t split_tiles = (P.n_streams + 31) / 32;
  const bool split_on = encoder_split_on(P.n_streams);
  // Few tiles per SM: the latency decoder (kernels_lat.cu, cabac_spec.cuh) -- context rows in the slots, successor rows
  // loaded ahead, LPS arm by table: fewer exposed latencies per bin, more instructions.  Measured on B200 over 65,536-bin
  // streams (profiles/r2_v1_ops_latency.jsonl): decode 4.42 / 5.42 / 5.32 ms against 4.97 / 5.92 / 5.67 ms of the wide
  // kernel at 32 / 8,192 / 16,384 streams, 6.51 against 5.98 at 32,768 -- so up to kLatTilesPerSm tiles per SM.  The latency
  // ENCODER (4.63 ms) loses to the two-warp encoder (4.24 ms) and is only used when forced.  ISSCABAC_LAT=0 / 1 forces
  // the choice for both directions.
  // Second half of round 2: with the window top-up of decw_refill_p the WIDE decoder is ahead at every size -- 3.95 / 4.52 /
  // 4.47 / 4.99 ms against 4.33 / 4.83 / 4.80 / 6.33 ms at 32 / 8,192 / 16,384 / 32,768 streams (118.6 against 129.7 cycles per
  // bin for a lone tile): what the latency formulation bought was the wait on the top-up's load, and it pays for it with 47
  // instructions per bin against 37 at one ALU-pipe instruction per 2.14 cycles.  kLatTilesPerSm = 0: forced use only.
  const char* lat_env = getenv("ISSCABAC_LAT");
  const bool lat_forced = lat_env && (lat_env[0] == '0' || lat_env[0] == '1');
  const bool lat_on = lat_forced ? lat_env[0] == '1' : (!ENC && split_tiles <= kLatTilesPerSm * (uint32_t)sm_count());
  if (lat_on && op_width == 1 && P.n_ctx <= 125) {
    bool done = false;
    const int rc_lat = launch_lat_codec(ENC, P, st, done);
    if (rc_lat || done) return rc_lat;
  }
  if (ENC && split_on && op_width == 1 && P.n_ctx <= 125) {
    const uint32_t sms = (uint32_t)sm_count();
    const uint32_t tiles = split_tiles;
    const size_t per_pair = ((size_t)P.n_ctx + 1) * WIDE_CTX_STRIDE + SPLIT_RING;
    uint32_t np_max = (uint32_t)((lim - WIDE_TAB_BYTES) / per_pair);
    if (np_max > SPLIT_MAX_PAIRS) np_max = SPLIT_MAX_PAIRS;
    if (np_max >= 1) {
      const uint32_t ctas_per_sm = (tiles + sms * np_max - 1) / (sms * np_max);
      uint32_t np = (tiles + sms * ctas_per_sm - 1) / (sms * ctas_per_sm);
      if (np > np_max) np = np_max;
      if (np < 1) np = 1;
      const uint32_t sgrid = (tiles + np - 1) / np;
      const size_t ssmem = WIDE_TAB_BYTES + per_pair * np;
      cudaError_t e = cudaFuncSetAttribute(k_encode_ops_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
      k_encode_ops_split<<<sgrid, np * 64, ssmem, st>>>(P);
      e = cudaGetLastError();
      return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_encode_ops_split");
    }
  }
  if (op_width == 1 && wide_geometry(P.n_streams, P.n_ctx, nw, grid, wsmem)) {
#if !WIDE_CTX_ROWS
    // one CTA per SM with 4k + 2 warps: the hand-over variant (see k_encode_ops_wide_ho)
    const char* ho_frac = getenv("ISSCABAC_HANDOVER_EIGHTHS");       // measured at C3: 4/8 7.05, 5/8 7.03, 6/8 7.02 ms (plain kernel 7.19)
    P.ho_eighths = ho_frac && ho_frac[0] >= '1' && ho_frac[0] <= '7' ? (uint32_t)(ho_frac[0] - '0') : 6u;
#if !CABAC_DEC_TMA
    if (!ENC && decode_handover_geometry(P.n_streams, nw, grid, wsmem, lim)) {
      const size_t hsmem = wsmem + HOD_BYTES;
      P.prefetch_ops = 0u;
      cudaError_t e = cudaFuncSetAttribute(k_decode_ops_wide_ho, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
      k_decode_ops_wide_ho<<<grid, (nw + 2u) * 32u, hsmem, st>>>(P);
      e = cudaGetLastError();
      return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_decode_ops_wide_ho");
    }
#endif
    if (ENC && handover_geometry(nw, grid, wsmem, lim)) {
      const size_t hsmem = wsmem + HO_BYTES;
      cudaError_t e = cudaFuncSetAttribute(k_encode_ops_wide_ho, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
      k_encode_ops_wide_ho<<<grid, (nw + 2u) * 32u, hsmem, st>>>(P);
      e = cudaGetLastError();
      return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_encode_ops_wide_ho");
    }
#endif
    auto kernel = ENC ? k_encode_ops_wide : k_decode_ops_wide;
    P.prefetch_ops = (!ENC && split_tiles <= kPrefetchOpsTilesPerSm * (uint32_t)sm_count()) ? 1u : 0u;
#if CABAC_DEC_TMA
    if (!ENC) wsmem += (size_t)nw * 32 * (kTmaLaneStride + 16);     // rings + mbarriers of the bulk-copy experiment
#endif
    if (wsmem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    }
    kernel<<<grid, nw * 32, wsmem, st>>>(P);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, ENC ? "k_encode_ops_wide" : "k_decode_ops_wide");
  }
  // keep at least 2 CTAs per SM resident on the shared-memory path
  const bool use_smem = smem_ctx <= lim / 2;
  int rc;
  if (use_smem) {
    if (op_width == 1)
      rc = ENC ? launch_codec(k_encode_ops<1, CtxSmem>, P, smem_ctx, st, "k_encode_ops")
               : launch_codec(k_decode_ops<1, CtxSmem>, P, smem_ctx, st, "k_decode_ops");
    else
      rc = ENC ? launch_codec(k_encode_ops<2, CtxSmem>, P, smem_ctx, st, "k_encode_ops")
               : launch_codec(k_decode_ops<2, CtxSmem>, P, smem_ctx, st, "k_decode_ops");
    return rc;
  }
  void* scratch = nullptr;
  if ((rc = keep_pool_cached())) return rc;
  cudaError_t e = cudaMallocAsync(&scratch, (size_t)P.n_ctx * P.n_streams, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(ctx scratch)");
  P.ctx_scratch = static_cast<uint8_t*>(scratch);
  if (op_width == 1)
    rc = ENC ? launch_codec(k_encode_ops<1, CtxGmem>, P, tab_bytes, st, "k_encode_ops")
             : launch_codec(k_decode_ops<1, CtxGmem>, P, tab_bytes, st, "k_decode_ops");
  else
    rc = ENC ? launch_codec(k_encode_ops<2, CtxGmem>, P, tab_bytes, st, "k_encode_ops")
             : launch_codec(k_decode_ops<2, CtxGmem>, P, tab_bytes, st, "k_decode_ops");
  cudaFreeAsync(scratch, st);
  return rc;
}

}  // namespace

extern "C" {

int cabac_encode_ops(uint32_t n_streams, const uint64_t* d_op_off, const void* d_ops, int op_width,
                     const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                     uint8_t* d_slab, uint64_t slab_stride, uint32_t* d_lengths,
                     uint32_t* d_overflow, void* stream) {
  if (n_streams && (!d_op_off || !d_slab || !d_lengths || (n_ctx && !d_ctx_init))) {
    set_error("cabac_encode_ops: null pointer");
    return ISSCABAC_ERR_INVALID;
  }
  if ((slab_stride & 15u) || slab_stride < 16 || (reinterpret_cast<uintptr_t>(d_slab) & 15u)) {
    set_error("cabac_encode_ops: slab and slab_stride must be 16-byte aligned");
    return ISSCABAC_ERR_INVALID;
  }
  CodecParams P;
  memset(&P, 0, sizeof P);
  P.n_streams = n_streams; P.n_ctx = n_ctx; P.per_stream_init = per_stream_init;
  P.op_off = d_op_off; P.ops = d_ops; P.ctx_init = d_ctx_init;
  P.slab = d_slab; P.slab_stride = slab_stride; P.lengths = d_lengths; P.overflow = d_overflow;
  return run_codec<true>(P, op_width, static_cast<cudaStream_t>(stream));
}

// Which kernel formulation cabac_encode_ops picks for a call of this shape (u8 ops): the name the profiler will show.
const char* cabac_encode_ops_kernel(uint32_t n_streams, uint32_t n_ctx) {
  const size_t lim = smem_limit();
  if (!lim || n_streams == 0) return "";
  const char* lat_env = getenv("ISSCABAC_LAT");
  if (lat_env && lat_env[0] == '1' && n_ctx <= 125) return "k_encode_ops_lat";
  if (encoder_split_on(n_streams) && n_ctx <= 125) return "k_encode_ops_split";
  uint32_t nw, grid;
  size_t wsmem;
  if (wide_geometry(n_streams, n_ctx, nw, grid, wsmem)) return handover_geometry(nw, grid, wsmem, lim) ? "k_encode_ops_wide_ho" : "k_encode_ops_wide";
  return "k_encode_ops";
}

const char* cabac_decode_ops_kernel(uint32_t n_streams, uint32_t n_ctx) {
  const size_t lim = smem_limit();
  if (!lim || n_streams == 0) return "";
  const char* lat_env = getenv("ISSCABAC_LAT");
  if (lat_env && lat_env[0] == '1' && n_ctx <= 125) return "k_decode_ops_lat";
  uint32_t nw, grid;
  size_t wsmem;
  if (wide_geometry(n_streams, n_ctx, nw, grid, wsmem)) {
    return decode_handover_geometry(n_streams, nw, grid, wsmem, lim) ? "k_decode_ops_wide_ho" : "k_decode_ops_wide";
  }
  return "k_decode_ops";
}

// <= 6 shift bits per context bin, 1 per bypass bin, 7 per terminate bin, <= 13 tail bits
// in finish(): 7 bits per op is a proof-level bound.
uint64_t cabac_slab_stride_bound(uint64_t max_ops) {
  uint64_t b = (max_ops * 7 + 7) / 8 + 8;
  return (b + 15) & ~15ull;
}

int cabac_decode_ops(uint32_t n_streams, const uint64_t* d_byte_off, const uint8_t* d_bytes,
                     const uint64_t* d_op_off, const void* d_ops, int op_width,
                     const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                     uint8_t* d_bins, uint8_t* d_finish_ok, void* stream) {
  if (n_streams && (!d_byte_off || !d_bytes || !d_op_off || !d_bins || (n_ctx && !d_ctx_init))) {
    set_error("cabac_decode_ops: null pointer");
    return ISSCABAC_ERR_INVALID;
  }
  CodecParams P;
  memset(&P, 0, sizeof P);
  P.n_streams = n_streams; P.n_ctx = n_ctx; P.per_stream_init = per_stream_init;
  P.op_off = d_op_off; P.ops = d_ops; P.ctx_init = d_ctx_init;
  P.byte_off = d_byte_off; P.bytes = d_bytes; P.bins = d_bins; P.finish_ok = d_finish_ok;
  return run_codec<false>(P, op_width, static_cast<cudaStream_t>(stream));
}

int cabac_pack_bins(const uint8_t* d_bins, uint64_t bit_begin, uint64_t bit_end, uint8_t* d_packed, void* stream) {
  if (bit_begin & 7u) { set_error("cabac_pack_bins: bit_begin must be a multiple of 8"); return ISSCABAC_ERR_INVALID; }
  if (bit_end <= bit_begin) return ISSCABAC_OK;
  if (!d_bins || !d_packed) { set_error("cabac_pack_bins: null pointer"); return ISSCABAC_ERR_INVALID; }
  const uint64_t n = bit_end - bit_begin, threads = (n + 15) / 16;
  k_pack_bins<<<(uint32_t)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_bins + bit_begin, n, d_packed + (bit_begin >> 3));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_pack_bins");
}

size_t cabac_compact_scratch_bytes(uint32_t n_streams) {
  size_t tiles = ((size_t)n_streams + SCAN_TILE - 1) / SCAN_TILE;
  return (tiles + 1) * sizeof(unsigned long long) + 256;
}

}  // extern "C"

namespace isscabac_internal {

int exclusive_scan_u32_u64(const uint32_t* d_in, uint64_t* d_out, uint64_t n, void* d_scratch, cudaStream_t st) {
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(uint64_t), st);
    return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "cudaMemsetAsync");
  }
  uint32_t tiles = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
  unsigned long long* desc = static_cast<unsigned long long*>(d_scratch);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(desc + tiles);
  k_scan_init<<<(tiles + 255) / 256, 256, 0, st>>>(desc, tiles, ticket);
  k_scan_u32_u64<<<tiles, SCAN_THREADS, 0, st>>>(d_in, d_out, n, desc, ticket);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_scan_u32_u64");
}

}  // namespace isscabac_internal

extern "C" int cabac_compact(uint32_t n_streams, const uint8_t* d_slab, uint64_t slab_stride,
                             const uint32_t* d_lengths, uint8_t* d_payload, uint64_t payload_cap,
                             uint64_t* d_byte_off, void* d_scratch, uint32_t* d_overflow, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!d_byte_off || (n_streams && (!d_slab || !d_lengths || !d_scratch))) {
    set_error("cabac_compact: null pointer");
    return ISSCABAC_ERR_INVALID;
  }
  int rc = exclusive_scan_u32_u64(d_lengths, d_byte_off, n_streams, d_scratch, st);
  if (rc || n_streams == 0 || !d_payload) return rc;
  if (slab_stride <= kCompactShortRow) {      // short rows: 8 lanes per row, four times the rows in flight
    const uint32_t blocks = (uint32_t)(((uint64_t)n_streams * 8 + 255) / 256);
    k_compact_copy<8><<<blocks, 256, 0, st>>>(d_slab, slab_stride, d_lengths, d_byte_off, d_payload, payload_cap, n_streams, d_overflow);
  } else {
    const uint32_t blocks = (uint32_t)(((uint64_t)n_streams * 32 + 255) / 256);
    k_compact_copy<32><<<blocks, 256, 0, st>>>(d_slab, slab_stride, d_lengths, d_byte_off, d_payload, payload_cap, n_streams, d_overflow);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_compact_copy");
}
