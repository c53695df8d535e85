// cabac_spec.cuh -- the per-lane coder restated for the LATENCY-bound regime (few streams per SM:
// strong scaling of the stream set over GPUs, the chunk kernels of the host-buffer path, ragged jobs
// that end on their longest stream).  Same arithmetic as cabac_wide.cuh (and therefore as
// CABAC_ArithmeticEncoder.cpp:113-178,250-270 / CABAC_ArithmeticDecoder.cpp:87-190,288-331), same
// 64-bit windows and the same block schedule; what changes is what sits on the dependent chain of a bin.
//
// cabac_wide.cuh per bin (one warp alone on its scheduler: 148 cycles decode, 161 encode on B200):
//     LDS token -> LDS.128 row -> PRMT lps -> SUB -> SHL -> ISETP -> SEL -> BFIND -> IMAD -> SHF
//     -> STS token, and the next bin's token load has to follow that store.
// Here:
//  * a context slot holds the 16-byte ROW of its state, not a token for it: one load, issued a whole
//    bin ahead (the op codes are known up front), gives everything the next bin needs;
//  * both successor rows (after an MPS / after an LPS) of the current row are loaded while the decision
//    is still being computed; the decision only selects between registers.  A bin on the same
//    context as its predecessor takes the selected successor, any other bin the row loaded ahead;
//  * a row carries the LPS arm precomputed per range quartile -- range and renormalisation shift
//    after an LPS depend only on (state, quartile) -- and the MPS arm needs one compare (the shift is
//    0 or 1), so BFIND/IMAD/SHF leave the chain: both arms are ready before the decision;
//  * bypass bins: the quartile selector is range >> 16 = 0 instead of range >> 6, which makes the LPS
//    sub-range 0, rMPS = range, and scaledRange is rMPS << 21 instead of << 22; range is kept by
//    one select off the chain.  No dummy row, no branch.
// The loop-carried chain of a bin is SHF -> PRMT -> SUB -> SHF -> ISETP -> SEL (range) resp.
// SEL -> SEL -> [LOP3 -> LDS.128] (same context twice in a row).
//
// __host__ __device__ like the other headers: tests/emul/lane_emul.cpp compiles this file with g++.
#pragma once
#include "cabac_wide.cuh"

namespace cabac {

// Row of a context state byte st = (state << 1) | mps.  Tokens name rows: on the device the shared-window
// address of the row in this lane's column of the table, in the host emulation the state byte.
struct alignas(16) SRow {
  uint32_t lps4;   // four LPS sub-ranges (CABAC_ArithmeticEncoder.cpp:414-480)
  uint32_t nn4;    // per quartile: renormalisation shift after an LPS (1..6, Encoder.cpp:482-492) | MPS value << 7
  uint32_t tok_m;  // token of the row after an MPS (ContextModel.cpp:136-146)
  uint32_t tok_l;  // token of the row after an LPS (ContextModel.cpp:148-158)
};
constexpr uint32_t kSpecMpsBit = 0x80u;     // in every byte of nn4

CB_HD constexpr uint32_t spec_lps_shift(uint32_t lps) {      // min(clz(lps) - 23, 6) for 2 <= lps <= 240
  return lps >= 128 ? 1u : lps >= 64 ? 2u : lps >= 32 ? 3u : lps >= 16 ? 4u : lps >= 8 ? 5u : 6u;
}
// the row in terms of state bytes (tokens = state bytes)
CB_HD constexpr SRow spec_row(uint32_t st) {
  const uint32_t l4 = kTables.lps[(st >> 1) & 63u];
  uint32_t nn = 0;
  for (int q = 0; q < 4; ++q) nn |= (spec_lps_shift((l4 >> (8 * q)) & 0xffu) | ((st & 1u) << 7)) << (8 * q);
  return SRow{l4, nn, next_mps(st), next_lps(st)};
}

// what a bin needs to know about its op besides the row: hoisted out of the chain (the ops are known up front)
struct SpecOp {
  uint32_t qsh;    // range >> qsh is the quartile selector: 6, or 16 for a bypass bin
  uint32_t ksh;    // scaledRange = rMPS << ksh: 22, bypass 21 (rMPS = range there)
  bool ep;
};
CB_HD SpecOp spec_op(uint32_t code, uint32_t n_ctx) {
  const bool ep = code >= n_ctx;        // bypass, incl. any code that is not a context of this call (see cabac_wide.cuh)
  return SpecOp{ep ? 16u : 6u, ep ? 21u : 22u, ep};
}

// ---------------------------------------------------------------------------
// decoder: one bin (decodeBin / decodeBinEP).  Returns the bin (0/1); lps_out = LPS decided (bypass: bin).
// ---------------------------------------------------------------------------
CB_HD uint32_t decs_bin(DecWide& D, const SpecOp op, const SRow& R, bool& lps_out) {
  const uint32_t q = D.range >> op.qsh;                        // 4..7, bypass: 0
  const uint32_t lps = cb_prmt(0, R.lps4, q);                  // bypass: 0
  const uint32_t rmps = D.range - lps;
  const uint32_t scaled = rmps << op.ksh;
  const bool p = D.hi >= scaled;                               // unsigned like the reference: a corrupt stream may hold any value
  const uint32_t h = p ? D.hi - scaled : D.hi;
  // LPS arm, from the row: shift by table, range = lps << shift (a bypass bin keeps its range and shifts by one)
  const uint32_t nnm = cb_prmt(0, R.nn4, q);                   // bypass: 0
  const uint32_t nnL = nnm & 7u;
  const uint32_t rangeL = op.ep ? D.range : lps << nnL;
  const uint32_t nsL = op.ep ? 1u : nnL;
  // MPS arm: at most one shift (Decoder.cpp:156-166)
  const bool pm = rmps < 256u;
  const uint32_t rangeM = pm ? 2u * rmps : rmps;
  const uint32_t nsM = (pm || op.ep) ? 1u : 0u;
  const uint32_t ns = p ? nsL : nsM;
  D.range = p ? rangeL : rangeM;
  D.hi = cb_funnel_l(D.lo, h, ns);
  D.lo <<= ns;
  D.f += (int32_t)ns;
  lps_out = p;
  return (p ? 1u : 0u) ^ (nnm >> 7);                           // bin = MPS ^ isLPS; bypass: nnm = 0
}

// ---------------------------------------------------------------------------
// encoder: one bin (encodeBin / encodeBinEP); p = the bin is the LPS (bypass: the bin itself)
// ---------------------------------------------------------------------------
CB_HD void encs_bin(EncWide& E, const SpecOp op, const SRow& R, bool p) {
  const uint32_t q = E.range >> op.qsh;
  const uint32_t lps = cb_prmt(0, R.lps4, q);
  const uint32_t rmps = E.range - lps;
  const uint32_t x2 = op.ep ? E.range : 2u * rmps;
  const uint32_t nnL = cb_prmt(0, R.nn4, q) & 7u;
  const uint32_t rangeL = op.ep ? E.range : lps << nnL;
  const uint32_t nsL = op.ep ? 1u : nnL;
  const bool pm = rmps < 256u;
  const uint32_t rangeM = pm ? 2u * rmps : rmps;
  const uint32_t nsM = (pm || op.ep) ? 1u : 0u;
  const uint32_t ns = p ? nsL : nsM;
  uint64_t W = E.W;
  if (p) W += x2;
  E.W = W << ns;
  E.range = p ? rangeL : rangeM;
  E.n += (int32_t)ns;
}

CB_HD SRow spec_sel(bool p, const SRow& a, const SRow& b) {
  return SRow{p ? a.lps4 : b.lps4, p ? a.nn4 : b.nn4, p ? a.tok_m : b.tok_m, p ? a.tok_l : b.tok_l};
}

// Context memory as the kernels see it (Mem): ldrow(token) = a table row, ldctx(c) / stctx(c, row) = the row
// held by context slot c of this lane (slot n_ctx = the slot bypass ops address; its content is never used).

// One op of a block, decoder.  R = the row of this op's slot on entry, the row of the NEXT op's slot on return
// (c_next = that slot; c_next == c: the successor just selected, otherwise the row loaded ahead -- loaded before this
// op's store, which it can only miss when c_next == c).
template <bool HAS_NEXT, class Mem>
CB_HD uint32_t decs_step(DecWide& D, SRow& R, uint32_t code, uint32_t c, uint32_t c_next, const Mem& mem, uint32_t n_ctx) {
  SRow rowN = R;
  if (HAS_NEXT) rowN = mem.ldctx(c_next);
  const SRow rowM = mem.ldrow(R.tok_m), rowL = mem.ldrow(R.tok_l);
  bool p;
  const uint32_t bin = decs_bin(D, spec_op(code, n_ctx), R, p);
  const bool ctx_lps = code >= n_ctx ? false : p;          // a bypass bin leaves its (dummy) slot on the MPS path
  const SRow nr = spec_sel(ctx_lps, rowL, rowM);
  mem.stctx(c, nr);
  if (HAS_NEXT) R = spec_sel(c_next == c, nr, rowN);
  return bin;
}

// One op of a block, encoder: the bin is known, so only the successor that will be taken is loaded.
template <bool HAS_NEXT, class Mem>
CB_HD void encs_step(EncWide& E, SRow& R, uint32_t code, uint32_t bin, uint32_t c, uint32_t c_next, const Mem& mem, uint32_t n_ctx) {
  SRow rowN = R;
  if (HAS_NEXT) rowN = mem.ldctx(c_next);
  const bool ep = code >= n_ctx;
  const bool is_lps = (((R.nn4 >> 7) ^ bin) & 1u) != 0u;
  const SRow nr = mem.ldrow((is_lps && !ep) ? R.tok_l : R.tok_m);
  encs_bin(E, spec_op(code, n_ctx), R, ep ? bin != 0u : is_lps);
  mem.stctx(c, nr);
  if (HAS_NEXT) R = spec_sel(c_next == c, nr, rowN);
}

CB_HD uint32_t spec_slot(uint32_t code, uint32_t n_ctx) { return code < n_ctx ? code : n_ctx; }

// 16 ops without a terminate op; same groups of four, same window budget and the same (voted) emission /
// refill points as encw_block16 / decw_block16.  c_after = slot of the op that follows the block
// (n_ctx when there is none or it is unknown: the row is then reloaded by whoever needs it).
template <bool VOTE, class Mem>
CB_HD void encs_block16(EncWide& E, const uint32_t w[4], const uint32_t cw[4], const Mem& mem, uint32_t n_ctx) {
  uint32_t code[16], c[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    code[k] = cb_prmt(cw[k >> 2], 0, 0x4440u + (k & 3));
    c[k] = spec_slot(code[k], n_ctx);
  }
  SRow R = mem.ldctx(c[0]);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 4 * g + j;
      const uint32_t bin = (w[g] >> (8 * j)) & 1u;
      if (k < 15) encs_step<true>(E, R, code[k], bin, c[k], c[k < 15 ? k + 1 : 15], mem, n_ctx);
      else encs_step<false>(E, R, code[k], bin, c[k], c[k], mem, n_ctx);
      if (j == 1 && E.n >= kLazy) encw_emit(E);
    }
    if (cb_any<VOTE>(E.n >= kLazy)) encw_emit(E);
  }
}

template <bool VOTE, class Mem>
CB_HD void decs_block16(DecWide& D, const uint32_t cw[4], uint32_t r[4], const Mem& mem, uint32_t n_ctx) {
  uint32_t code[16], c[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    code[k] = cb_prmt(cw[k >> 2], 0, 0x4440u + (k & 3));
    c[k] = spec_slot(code[k], n_ctx);
  }
  SRow R = mem.ldctx(c[0]);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 4 * g + j;
      uint32_t bin;
      if (k < 15) bin = decs_step<true>(D, R, code[k], c[k], c[k < 15 ? k + 1 : 15], mem, n_ctx);
      else bin = decs_step<false>(D, R, code[k], c[k], c[k], mem, n_ctx);
      acc |= bin << (8 * j);
      if (CABAC_LAZY_DEC == 1 && j == 1 && D.f >= kLazyDec) decw_refill(D);
    }
    r[g] = acc;
    if (VOTE && CABAC_REFILL_P_LAT == 3) {
      decw_refill_p<true>(D);
    } else if (cb_any<VOTE>(D.f >= kLazyDec)) {
      if (VOTE && CABAC_REFILL_P_LAT) decw_refill_p<(CABAC_REFILL_P_LAT > 1)>(D);   // cabac_wide.cuh, CABAC_REFILL_P
      else decw_refill(D);
    }
  }
}

// general path: any op kind, one at a time (stream heads up to the 16-byte boundary, tails, blocks with terminate ops)
template <class Mem>
CB_HD void encs_general(EncWide& E, uint32_t o, const Mem& mem, uint32_t n_ctx) {
  const uint32_t code = o >> 1;
  if (code == kOpTrmCode) {
    encw_trm(E, o & 1u);
  } else {
    const uint32_t c = spec_slot(code, n_ctx);
    SRow R = mem.ldctx(c);
    encs_step<false>(E, R, code, o & 1u, c, c, mem, n_ctx);
  }
  encw_emit(E);
}
template <class Mem>
CB_HD uint32_t decs_general(DecWide& D, uint32_t o, const Mem& mem, uint32_t n_ctx) {
  const uint32_t code = o >> 1;
  uint32_t bin;
  if (code == kOpTrmCode) {
    bin = decw_trm(D);
  } else {
    const uint32_t c = spec_slot(code, n_ctx);
    SRow R = mem.ldctx(c);
    bin = decs_step<false>(D, R, code, c, c, mem, n_ctx);
  }
  decw_refill(D);
  return bin;
}

}  // namespace cabac
