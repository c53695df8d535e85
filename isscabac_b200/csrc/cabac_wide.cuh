// cabac_wide.cuh -- the wide-window formulation of the per-lane coder (the hot path).
//
// Same arithmetic as the reference engine (CABAC_ArithmeticEncoder.cpp:113-178,250-270,
// 326-412; CABAC_ArithmeticDecoder.cpp:54-190,288-331,423-472), restated so that a lane
// touches memory once per 32 coded BITS instead of once per byte, and so that the per-bin
// step is one straight-line sequence with no data-dependent branch:
//
//  encoder  The reference keeps `low` in 32 bits and moves one byte out whenever fewer
//           than 12 free bits remain, parking it in bufferedByte/numBufferedBytes until a
//           carry can no longer reach it (writeOut, :380-412).  Here the window is 64 bits
//           (kept as W = 2*low, see below), whole big-endian 32-bit words leave it at fixed
//           points of the instruction stream (after every 4th bin), and a carry is added to
//           the previously emitted word, which is still in a register (`pend`); only a carry
//           into a 0xFFFFFFFF word has to walk back through memory.  Every emitted bit is a
//           bit >= 9 of the reference's `low`, i.e. one that later bins can only change by
//           a carry, so the bytes are identical.
//  decoder  The reference keeps 9+7 bits of `value` plus up to 8 look-ahead bits and reads
//           one byte whenever they run out (:137-141,156-162,295-299).  Here the window
//           holds value<<15 followed by up to 47 look-ahead bits and is topped up with one
//           32-bit word after every 4th bin.  All decisions compare bits the reference has
//           already read, so the decoded bins are identical; bytes past the end of a stream
//           read as 0xFF like CABAC_BitstreamFile.cpp:153-158.
//  bypass   W = 2*low makes encodeBinEP the same expression as the LPS branch of encodeBin:
//           W' = (W + add) << n with add = bin ? range : 0, n = 1 (context bins: add =
//           isLPS ? 2*rMPS : 0, n = renorm shift).  A bypass op is therefore coded as a
//           context bin on a dummy context slot whose state byte is 0x80 (mps bit 0, both
//           next states 0x80); three selects pick range / 1 / range instead of the
//           context-path values.  The decoder mirrors this with value - (bin ? range<<21 : 0).
//
// Everything is __host__ __device__ so that tests/emul compiles exactly this code with g++.
#pragma once
#include "cabac_lane.cuh"

namespace cabac {

constexpr uint32_t kEpState = 128;     // dummy state byte of the bypass slot
constexpr uint32_t kNumRows = 129;     // fused rows 0..127 + the bypass row

// Rows as the wide kernels use them.  A context slot does not hold the state byte but a TOKEN for
// it: on the device the shared-memory address of this lane's copy of the state's row, in the host
// emulation the state byte itself.  The row load therefore needs no address arithmetic, and the
// next token comes straight out of the row (one select) instead of being extracted from a packed
// byte pair and turned into an address again.  mps4 = the MPS bit replicated into bit 0 of every
// byte, so that "is this bin the LPS" is one logic op against the op word whatever byte of the
// word the op sits in.
struct alignas(16) WRow {
  uint32_t lps4;       // four LPS sub-ranges (0 for the bypass row)
  uint32_t next_mps;   // token of the state after an MPS
  uint32_t next_lps;   // token of the state after an LPS
  uint32_t mps4;       // (st & 1) * 0x01010101
};
// the row in terms of state bytes (next_* = state bytes): the host emulation uses it as is, the
// device table replaces the two next states by tokens
CB_HD constexpr WRow wide_row(uint32_t st) {
  return st < 128 ? WRow{fused_row(st).x, next_mps(st), next_lps(st), (st & 1u) * 0x01010101u}
                  : WRow{0u, kEpState, kEpState, 0u};
}

CB_HD uint32_t cb_bswap(uint32_t x) { return cb_perm(x, 0, 0x0123); }

// Warp vote: true when `p` holds for any lane executing this call together with the caller (the
// host emulation is one lane).  Used to move words out of / into the window LAZILY: with 32
// unrelated streams per warp some lane crosses the 32-bit mark at almost every check, so an eager
// "if (n >= 32) emit" makes the whole warp walk through the emission code on every check with a
// tenth of its lanes active.  Waiting until some lane is about to run out of window lets all lanes
// that have a full word move it at the same time, a third as often.
#ifndef CABAC_LAZY
#define CABAC_LAZY 1
#endif
#ifndef CABAC_DEC_TMA
#define CABAC_DEC_TMA 0      // decoder input staged by bulk copies: compile-time experiment, see below
#endif
// VOTE = all 32 lanes of the warp execute this call together (the caller guarantees it: the kernels
// walk the blocks every lane of a full warp has in lockstep); otherwise this lane decides alone
// CABAC_EXPECT: rare branches marked as such, so that the straight path of a block has no taken branch (a taken branch or
// the BSYNC behind it costs a lone warp 15 - 45 cycles, profiles/r2_tree_decoder_latency.txt)
#ifndef CABAC_EXPECT
#define CABAC_EXPECT 1
#endif
#if CABAC_EXPECT
#define CB_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define CB_UNLIKELY(x) (x)
#endif
template <bool VOTE>
CB_HD bool cb_any(bool p) {
#if defined(__CUDA_ARCH__) && CABAC_LAZY
  return VOTE ? __any_sync(0xffffffffu, p) != 0 : p;
#else
  return p;
#endif
}

// Window budget of a 4-bin group (<= 6 bits per bin, the window holds 53): a lane enters the group
// with at most kLazy-1 = 41 bits, may reach 53 after two bins, is brought back under 32 by the
// mid-group guard if it passed 41, and leaves the group with at most 53.
constexpr int kLazy = CABAC_LAZY ? 42 : 32;
// The decoder's refill: 0 = eager, 1 = the encoder's schedule (voted at 42 + mid-group guard),
// 2 = voted at 36 without a guard.  Measured on B200 at C3: 523 / 517 / 546 Gbins/s.
#ifndef CABAC_LAZY_DEC
#define CABAC_LAZY_DEC 2
#endif
// 2 = voted refill without the mid-group guard: a lane enters a group with at most 35 unfilled bits,
// so the fourth decision of the group still sees f <= 35 + 18 = 53
constexpr int kLazyDec = CABAC_LAZY_DEC == 2 ? 36 : (CABAC_LAZY_DEC ? 42 : 32);

// renormalisation shift min(clz(rsel) - 23, 6) = 8 - bfind(rsel | 4) (Encoder.cpp:482-492 incl. the state-63 row).
// CABAC_FMA_RENORM: the subtraction as a multiply-add, i.e. on the FMA pipe instead of the ALU pipe, which is the one
// that binds the hot kernels (B200, C3: decode 546 -> 556 Gbins/s, encode 587 -> 590; the same trick for the bit
// counters -- a 3-input add covers two bins -- and for range >> 6 as a high multiply measured slower).
#ifndef CABAC_FMA_RENORM
#define CABAC_FMA_RENORM 1
#endif
#if defined(__CUDACC__)
// a multiplier ptxas cannot fold: with a literal -1 it turns x * -1 + 8 back into an ALU-pipe subtraction
static __constant__ int c_cb_neg1 = -1;
#endif
CB_HD int cb_renorm(uint32_t rsel) {
#if defined(__CUDA_ARCH__) && CABAC_FMA_RENORM
  uint32_t fl;
  int nn;
  asm("bfind.u32 %0, %1;" : "=r"(fl) : "r"(rsel | 4u));
  asm("mad.lo.s32 %0, %1, %2, 8;" : "=r"(nn) : "r"(fl), "r"(c_cb_neg1));
  return nn;
#else
  return cb_clz(rsel | 4u) - 23;
#endif
}

// x -= p ? y : 0.  CABAC_PRED_SUB: as a PREDICATED multiply-add (x = y * -1 + x, the multiplier in constant memory so
// that ptxas keeps the IMAD form): the select leaves the ALU pipe (B200, C3: decode 556 -> 561 Gbins/s).  The encoder's
// 64-bit "low += p ? y : 0" did not gain from the same treatment: 590 -> 541 as a predicated IMAD.WIDE, -> 566 as a
// predicated add / add-with-carry pair.
#ifndef CABAC_PRED_SUB
#define CABAC_PRED_SUB 1
#endif
CB_HD uint32_t cb_sub_if(uint32_t x, bool p, uint32_t y) {
#if defined(__CUDA_ARCH__) && CABAC_PRED_SUB
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q mad.lo.u32 %0, %2, %3, %0;\n\t}" : "+r"(x) : "r"((uint32_t)p), "r"(y), "r"(c_cb_neg1));
  return x;
#else
  return x - (p ? y : 0u);
#endif
}

// p ? a : b.  CABAC_SEL_FMA (bit mask, one bit per use site): as "r = b; @p r = a * 1 + 0" with the multiplier in constant
// memory, i.e. the select on the FMA pipe instead of the ALU pipe.  Sites: 0 / 4 = the bypass override of the added term
// (encoder / decoder), 1 / 5 = LPS or MPS sub-range, 2 / 6 = the bypass override of the shift, 3 = next token.  Measured on
// B200 at C3 against none (590 / 561 Gbins/s): sites 0 + 4: 593 / 574 (kept); 2 + 6: 585 / 565; 1 + 5: 570 / 552 (the
// sub-range select sits on the range chain); 3: 577 / 561; all: 573 / 563.  TP = the caller is a throughput-bound kernel
// (the 16-op blocks of the op-array kernels): where a launch is bound by the latency of one chain -- the two-warp encoder at
// low occupancy, the fused symbol kernels -- the extra pipe crossing costs more than the ALU slot saves (4.39 -> 4.62 ms, C5
// decode 24.2 -> 25.8 ms), so those keep the plain select.
#ifndef CABAC_SEL_FMA
#define CABAC_SEL_FMA 0x11
#endif
#if defined(__CUDACC__)
static __constant__ uint32_t c_cb_one = 1u;
#endif
template <int SITE, bool TP = true>
CB_HD uint32_t cb_sel(uint32_t p, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  if (TP && ((CABAC_SEL_FMA >> SITE) & 1)) {
    uint32_t r = b;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q mad.lo.u32 %0, %2, %3, 0;\n\t}" : "+r"(r) : "r"(p), "r"(a), "r"(c_cb_one));
    return r;
  }
#endif
  return p ? a : b;
}

// ---------------------------------------------------------------------------
// encoder
// ---------------------------------------------------------------------------
struct EncWide {
  uint64_t W;         // 2 * (the reference's low), bits below the not-yet-emitted point
  uint32_t range;
  int32_t n;          // bits shifted in since the last emitted word boundary (reference: 23 - bitsLeft, mod emission)
  uint32_t pend;      // last emitted word (numeric value): still in a register so that a carry is one add
  uint32_t wp;        // index of the pending word = words emitted - 1 (0xffffffff before the first one);
                      // keeps counting past the capacity so that the caller sees len > cap
  uint32_t cap_words; // slab capacity in words
  uint32_t* slot;     // where the pending word will be stored (row + wp)
};

CB_HD void encw_start(EncWide& E, uint8_t* out, uint32_t cap_bytes) {  // Encoder.cpp:54-61
  E.W = 0; E.range = 510; E.n = 0; E.pend = 0; E.wp = 0xffffffffu;
  E.cap_words = cap_bytes >> 2;
  E.slot = cb_keep(reinterpret_cast<uint32_t*>(out) - 1);
}

// A carry out of a pending word that was 0xFFFFFFFF: +1 into the words already stored
// (those before the pending one).  Needs 32 one-bits in a row, i.e. practically never;
// replaces the buffered-0xFF-run handling of Encoder.cpp:394-404 and :76-87.
CB_HD_NOINLINE void encw_carry_walk(uint32_t wp, uint32_t cap_words, uint32_t* slot) {
  volatile uint32_t* o = slot - wp;   // row start
  for (int64_t i = (int64_t)wp - 1; i >= 0; --i) {
    if ((uint64_t)i >= cap_words) continue;
    uint32_t v = cb_bswap(o[i]) + 1u;
    o[i] = cb_bswap(v);
    if (v != 0) break;
  }
}

// the pending word leaves for memory, `carry` (0/1) added
CB_HD void encw_retire(EncWide& E, uint32_t carry) {
  const uint32_t prev = E.pend + carry;
  if (carry > prev && E.wp != 0xffffffffu) encw_carry_walk(E.wp, E.cap_words, E.slot);   // prev wrapped to 0
  if (E.wp < E.cap_words) cb_stg32(E.slot, cb_bswap(prev));   // false while nothing is pending (wp = 0xffffffff)
}

// Move one 32-bit word out when at least 32 settled bits are waiting.  Emits the bits
// [n-23, n+8] of the reference's low: the four bytes the reference would hand to
// writeOut() over its next four calls.  The word stays in `pend`; what is stored is the
// PREVIOUS word, after the carry of this emission has been added to it.
CB_HD void encw_emit(EncWide& E) {
  if (E.n >= 32) {
    const uint32_t sh = (uint32_t)E.n - 22u;  // 10..31
    const uint32_t lo = (uint32_t)E.W, hi = (uint32_t)(E.W >> 32);
    const uint32_t word = cb_funnel_r(lo, hi, sh);
    const uint32_t carry = hi >> sh;          // W < 2^(n+11): one carry bit above the word
    E.W = lo & ~(0xffffffffu << sh);
    E.n -= 32;
    encw_retire(E, carry);
    E.pend = word;
    E.wp++;
    E.slot++;
  }
}

// One bin, context-coded or bypass (see the file header).  row = the row of the slot the op
// addresses (the bypass row for a bypass op); the bin is bit 8*B of `w`.  Returns non-zero when the bin was the LPS.
template <int B, bool TP = false>
CB_HD uint32_t encw_bin(EncWide& E, uint32_t w, bool is_ep, const WRow& row) {
  const uint32_t lps = cb_prmt(0, row.lps4, E.range >> 6);   // selector 4..7 = lps4 byte q (range is 256..510)
  const uint32_t rmps = E.range - lps;
  const uint32_t is_lps = cb_xor_and<(1u << (8 * B))>(row.mps4, w);   // non-zero = LPS
  const uint32_t x2 = cb_sel<0, TP>(is_ep, E.range, 2u * rmps);
  const uint32_t rsel = cb_sel<1, TP>(is_lps, lps, rmps);
  const int nn = cb_renorm(rsel);
  const int ns = (int)cb_sel<2, TP>(is_ep, 1u, (uint32_t)nn);
  uint64_t W = E.W;
  if (is_lps) W += x2;                       // predicated 64-bit add
  E.W = W << ns;
  E.range = is_ep ? E.range : (rsel << nn);
  E.n += ns;
  return is_lps;
}

// encodeBinsEP (Encoder.cpp:278-319; byte-identical to n single encodeBinEP calls) in
// one step: low' = (low << n) + range * bits.  The caller keeps the window budget: E.n + n <= 53
// (the fused symbol kernels emit before and after a run, so n <= kEpRunMax with E.n < 32 on entry).
constexpr uint32_t kEpRunMax = 16;
CB_HD void encw_ep_run(EncWide& E, uint32_t bits, uint32_t n) {
  E.W = (E.W << n) + ((uint64_t)(E.range * bits) << 1);
  E.n += (int32_t)n;
}

// encodeBinTrm, Encoder.cpp:326-367 (the only op that is a real branch; a handful per stream)
CB_HD void encw_trm(EncWide& E, uint32_t bin) {
  E.range -= 2;
  if (bin) {
    E.W = (E.W + 2ull * E.range) << 7;
    E.range = 256;
    E.n += 7;
  } else if (E.range < 256) {
    E.W <<= 1; E.range <<= 1; E.n += 1;
  }
}

// finish, Encoder.cpp:70-105; returns the stream length in bytes
CB_HD uint32_t encw_finish(EncWide& E) {
  encw_emit(E);
  encw_trm(E, 1);
  encw_emit(E);
  const uint32_t cbit = (uint32_t)E.n + 10u;            // bit 9+n of low = the carry finish() tests (:76)
  const uint32_t carry = (uint32_t)(E.W >> cbit) & 1u;
  E.W &= ~(1ull << cbit);
  encw_retire(E, carry);                                // the pending word goes out now, with the final carry
  // write(low >> 8, 24 - bitsLeft) = n+1 bits, the stop bit, zero padding (:100-104)
  const uint32_t tb = (uint32_t)E.n + 2u;               // <= 33
  const uint64_t tail = ((((E.W >> 9) & ((1ull << (E.n + 1)) - 1ull)) << 1) | 1ull) << (64u - tb);
  const uint32_t nb = (tb + 7u) >> 3;
  uint8_t* o8 = reinterpret_cast<uint8_t*>(E.slot + 1);
  const uint32_t base = 4u * (E.wp + 1u);
  for (uint32_t j = 0; j < nb; ++j)
    if (base + j < 4u * E.cap_words) o8[j] = (uint8_t)(tail >> (56u - 8u * j));
  return base + nb;
}

// ---------------------------------------------------------------------------
// decoder
// ---------------------------------------------------------------------------
// Window R = hi:lo.  Stream byte m sits at bits [55-8m+s, 62-8m+s] after s shifts, so the
// reference's 16+ bit `value` is hi >> 15.  f = number of low bits of R not yet filled;
// with p = stream bytes loaded so far, f = 63 + s - 8p, which also gives the total shift
// count s back for finish().
// Input: aligned 32-bit words of the payload, two of them always in flight (cur, nxt) so that a
// refill never waits for memory; PRMT with a per-stream selector turns {cur,nxt} into the
// big-endian word of the next four stream bytes whatever the stream's byte alignment.
struct DecWide {
  uint32_t lo, hi;
  uint32_t range;
  int32_t f;
  uint32_t p;            // stream bytes loaded into the window (3 mod 4 after start)
  uint32_t len;          // stream length in bytes
  uint32_t cur, nxt;     // aligned little-endian words: cur holds stream byte p
  uint32_t sel;          // byte selector for {cur,nxt} -> big-endian word of stream bytes p..p+3
  uint32_t widx, wcnt;   // next aligned word to load / number of aligned words that hold stream bytes
  const uint32_t* wbase; // aligned word that holds stream byte 3
  const uint8_t* in;     // first byte of the stream
#if CABAC_DEC_TMA && defined(__CUDACC__)
  // experiment (see the CABAC_DEC_TMA block below): the stream's bytes staged in shared memory by 1-D bulk copies
  uint32_t ring = 0;     // shared-window address of this lane's two 64-byte stages (0: plain global loads)
  uint32_t mbar;         // shared-window address of this lane's two mbarriers
  uint32_t cwait;        // chunk (relative to g0) this lane has already waited for
  uint64_t g0;           // global address of chunk 0 (64-byte aligned)
  uint64_t gend;         // no chunk is copied past this address (end of the payload, rounded up to 16)
#endif
};

// ---------------------------------------------------------------------------
// CABAC_DEC_TMA (compile-time experiment, north_star (4): "pulls each stream's bytes into shared memory with TMA bulk
// copies").  Every lane owns a ring of two 64-byte stages and two mbarriers; chunk c of its stream (the 64-byte aligned
// piece of the payload that holds it) is brought in with ONE cp.async.bulk.shared.global (UBLKCP in SASS) that
// completes on the stage's mbarrier; the aligned words the window refill needs are then LDS.32 from the ring instead
// of per-lane 32-bit global loads.  Chunk c + 2 is requested when the last word of chunk c has been taken.
// Built only into tuning variants (tools/tune.py build tma:CABAC_DEC_TMA=1); result in profiles/r2_tma_decode_experiment.txt.
// ---------------------------------------------------------------------------
#if CABAC_DEC_TMA && defined(__CUDACC__)
constexpr uint32_t kTmaLaneStride = 144;     // 2 x 64 B + 16 B pad (keeps the stages 16-byte aligned, spreads the banks)
__device__ __forceinline__ void decw_tma_issue(DecWide& D, uint32_t c) {
  const uint64_t src = D.g0 + 64ull * c;
  if (src >= D.gend || src >= reinterpret_cast<uint64_t>(D.in) + D.len) return;   // nothing of this stream in the chunk
  const uint64_t left = D.gend - src;
  const uint32_t bytes = left < 64 ? (uint32_t)left : 64u;
  const uint32_t st = c & 1u, bar = D.mbar + 8u * st, dst = D.ring + 64u * st;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the stage was read through the generic proxy
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void decw_tma_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
               :: "r"(bar), "r"(parity) : "memory");
}
#endif


CB_HD uint32_t decw_load(DecWide& D) {
#if CABAC_DEC_TMA && defined(__CUDA_ARCH__)
  uint32_t v = 0u;
  if (!D.ring) {
    if (D.widx < D.wcnt) v = D.wbase[D.widx];
  } else if (D.widx < D.wcnt) {
    const uint64_t a = reinterpret_cast<uint64_t>(D.wbase) + 4ull * D.widx;
    const uint32_t c = (uint32_t)((a - D.g0) >> 6), st = c & 1u, o = (uint32_t)a & 63u;
    if (c != D.cwait) {                     // first word taken from this chunk
      decw_tma_wait(D.mbar + 8u * st, (c >> 1) & 1u);
      D.cwait = c;
    }
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(D.ring + 64u * st + o) : "memory");
    if (o == 60u) decw_tma_issue(D, c + 2u);
  }
  D.widx++;
  return v;
#else
  const uint32_t v = D.widx < D.wcnt ? D.wbase[D.widx] : 0u;
  D.widx++;
  return v;
#endif
}

// big-endian word of stream bytes p..p+3, 0xFF past the end (BitstreamFile.cpp:153-158)
CB_HD uint32_t decw_fetch(DecWide& D) {
  uint32_t w = cb_prmt(D.cur, D.nxt, D.sel);
  D.cur = D.nxt;
  D.nxt = decw_load(D);                       // needed two refills from now
  if (D.p + 4u > D.len) {                     // only at the very end of a stream
    const uint32_t rem = D.len > D.p ? D.len - D.p : 0u;
    w = rem ? (w | (0xffffffffu >> (8u * rem))) : 0xffffffffu;
  }
  D.p += 4u;
  return w;
}

CB_HD void decw_refill(DecWide& D) {
  if (D.f >= 32) {
    const uint32_t w = decw_fetch(D);
    const uint32_t k = (uint32_t)D.f - 32u;                // 0..23
    D.lo |= w << k;
    D.hi |= cb_funnel_l(w, 0u, k);                         // w >> (32-k), 0 when k == 0
    D.f -= 32;
  }
}

// decw_refill without a branch on the lane's own fill level: meant for call sites the WARP reaches together (behind a
// vote) and where some lanes top up and others do not.  Inside the divergent `if (f >= 32)` of decw_refill the rotation
// cur <- nxt <- load ends in a register move FROM the load's destination at the join, i.e. the warp waits there for the
// word it has only just requested (ncu on a warp of long streams in the tree decoder: 180 cycles per top-up on that one
// move, long_scoreboard).  Here the new word is loaded IN PLACE by a predicated load (the asm operand is read-write, so
// no copy can follow it) and everything else is a select.  Words past the end of the stream are not loaded; what the
// register then holds does not matter, decw_fetch's end-of-stream rule overwrites those bytes with 0xFF.
// CABAC_REFILL_P / CABAC_REFILL_P_LAT: the top-up behind the 16-op blocks of the op-array decoders (wide / latency
// formulation) -- 0: voted, decw_refill (branch on the lane's fill level); 1: voted, decw_refill_p without the prefetch;
// 2: voted, with it; 3: decw_refill_p after every group of four bins without a vote.  B200, decode of 65,536-bin streams:
//   65,536 streams (wide kernel)   0: 7.48 ms   1: 7.23   2: 7.19   3: 7.04
//    8,192 streams (latency kernel) 0: 5.43 ms   1: 4.98   2: 4.84   3: 5.26
//       32 streams (latency kernel) 0: 134.7 cycles per bin   1: 127.7   2: 129.5   3: 127.6
#ifndef CABAC_REFILL_P
#define CABAC_REFILL_P 3
#endif
#ifndef CABAC_REFILL_P_LAT
#define CABAC_REFILL_P_LAT 2
#endif
// CABAC_REFILL_FIRST (wide kernel, CABAC_REFILL_P = 3): the top-up in front of every group of four bins instead of behind it, so
// that the block's last top-up load is four bins old when the next block starts.  The kernel tops up once more when it leaves the
// lockstep blocks or takes the general path.
#ifndef CABAC_REFILL_FIRST
#define CABAC_REFILL_FIRST 1      // B200: lone tile 118.7 -> 109.9 cycles per bin, 8,192 streams 4.22 -> 3.99 ms, 65,536 7.11 -> 7.04
#endif
template <bool PF = true>
CB_HD void decw_refill_p(DecWide& D) {
#if defined(__CUDA_ARCH__) && !CABAC_DEC_TMA
  const bool need = D.f >= 32;
  uint32_t w = cb_prmt(D.cur, D.nxt, D.sel);
  D.cur = need ? D.nxt : D.cur;
  const uint32_t ld = (need && D.widx < D.wcnt) ? 1u : 0u;
  // The scoreboard of a load is the WARP's: whatever lane issued it, the next instruction that reads the register waits
  // for it -- with 32 lanes topping up in turn that is the prmt above, in every step.  So the sector after the one being
  // read is asked into L1 now (a hint: no register, no scoreboard); by the time a lane loads from it, eight top-ups
  // later, the load is an L1 hit.
#ifndef CABAC_PF_DIST
#define CABAC_PF_DIST 8      // words ahead of the one being loaded: the next sector
#endif
  const uint32_t pf = (PF && CABAC_PF_DIST && need && D.widx + CABAC_PF_DIST < D.wcnt) ? 1u : 0u;
  asm volatile("{\n\t.reg .pred q, r;\n\tsetp.ne.u32 q, %2, 0;\n\tsetp.ne.u32 r, %3, 0;\n\t@q ld.global.u32 %0, [%1];\n\t@r prefetch.global.L1 [%4];\n\t}"
               : "+r"(D.nxt) : "l"(D.wbase + D.widx), "r"(ld), "r"(pf), "l"(D.wbase + D.widx + CABAC_PF_DIST));
  D.widx += need ? 1u : 0u;
  // end of the stream: bytes past it read as 0xFF (decw_fetch), without a branch -- rem = stream bytes left, 0 .. 4
  const int32_t left = (int32_t)D.len - (int32_t)D.p;
  const uint32_t rem = (uint32_t)(left < 0 ? 0 : (left > 4 ? 4 : left));
  w |= __funnelshift_rc(0xffffffffu, 0u, 8u * rem);      // 0xffffffff >> (8 * rem), 0 for rem = 4
  w = need ? w : 0u;
  const uint32_t k = (uint32_t)D.f & 31u;     // need: f - 32 (0..23)
  D.lo |= w << k;
  D.hi |= cb_funnel_l(w, 0u, k);              // w >> (32-k), 0 when k == 0 or w == 0
  D.p += need ? 4u : 0u;
  D.f -= need ? 32 : 0;
#else
  decw_refill(D);
#endif
}

// start, Decoder.cpp:54-60: the reference reads two bytes; here the first seven go in.
CB_HD void decw_start(DecWide& D, const uint8_t* in, uint32_t len) {
  D.in = in; D.len = len; D.range = 510;
  const uint32_t b0 = len > 0 ? in[0] : 0xffu, b1 = len > 1 ? in[1] : 0xffu, b2 = len > 2 ? in[2] : 0xffu;
  D.hi = ((b0 << 24) | (b1 << 16) | (b2 << 8)) >> 1;
  D.lo = 0;
  D.p = 3; D.f = 63 - 24;
  const uintptr_t a = reinterpret_cast<uintptr_t>(in) + 3u;
  const uint32_t k = (uint32_t)(a & 3u);
  D.sel = (k + 3u) | ((k + 2u) << 4) | ((k + 1u) << 8) | (k << 12);
  D.wbase = cb_keep(reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3));
  // aligned words from the one holding byte 3 to the one holding byte len-1
  D.wcnt = len > 3 ? (uint32_t)(((a + (len - 4u)) >> 2) - (a >> 2)) + 1u : 0u;
  D.widx = 0;
#if CABAC_DEC_TMA && defined(__CUDA_ARCH__)
  D.g0 = reinterpret_cast<uint64_t>(D.wbase) & ~63ull;
  D.cwait = 0xffffffffu;
  if (D.wcnt && D.ring) { decw_tma_issue(D, 0u); decw_tma_issue(D, 1u); }
#endif
  D.cur = decw_load(D);
  D.nxt = decw_load(D);
  decw_refill(D);   // bytes 3..6
}

// One bin, context-coded or bypass; returns the bin in bit 8*B (all other bits 0), lps_out = the bin was the LPS.
template <int B, bool TP = false>
CB_HD uint32_t decw_bin(DecWide& D, bool is_ep, bool& lps_out, const WRow& row) {
  const uint32_t lps = cb_prmt(0, row.lps4, D.range >> 6);
  const uint32_t rmps = D.range - lps;
  // reference: scaledRange << 15 (bypass: compare before the shift); rMPS << 22 directly instead of (2 * rMPS) << 21 keeps
  // the chain range -> lps -> rMPS -> scaled -> decision one operation shorter
  // (in the 16-op blocks of the op-array kernels; the fused symbol decoder measured 10 % slower with it at C4 and keeps the
  // plain form)
  uint32_t scaled;
#if defined(__CUDA_ARCH__)
  if (TP && ((CABAC_SEL_FMA >> 4) & 1)) {
    scaled = rmps << 22;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q mad.lo.u32 %0, %2, 2097152, 0;\n\t}" : "+r"(scaled) : "r"((uint32_t)is_ep), "r"(D.range));
  } else
#endif
  {
    const uint32_t x2 = is_ep ? D.range : 2u * rmps;
    scaled = x2 << 21;
  }
  const bool is_lps = D.hi >= scaled;
  const uint32_t rsel = cb_sel<5, TP>(is_lps, lps, rmps);
  const int nn = cb_renorm(rsel);
  const int ns = (int)cb_sel<6, TP>(is_ep, 1u, (uint32_t)nn);
  const uint32_t h = cb_sub_if(D.hi, is_lps, scaled);
  D.hi = cb_funnel_l(D.lo, h, (uint32_t)ns);                // (h:lo) << ns, ns in 0..6
  D.lo <<= ns;
  D.range = is_ep ? D.range : (rsel << nn);
  D.f += ns;
  lps_out = is_lps;
  return (row.mps4 ^ (is_lps ? 0xffffffffu : 0u)) & (1u << (8 * B));   // bin = mps ^ isLPS
}

// decodeBinsEP (Decoder.cpp:333-421 == n single decodeBinEP calls) in one step.  n bypass decisions
// are a long division: with R the window, every step is bin = R >= (range << 53); R = (R - bin *
// (range << 53)) << 1, so after n steps the bins are q = (R >> (54 - n)) / range (R < range << 54
// on entry keeps q < 2^n) and R' = (rem << 54) + ((R mod 2^(54-n)) << n).  Reads the top 10 + n
// bits of the window: needs f <= 54 - n, i.e. n <= kEpRunMax after a refill (f < 32).
CB_HD uint32_t decw_ep_run(DecWide& D, uint32_t n) {
  const uint32_t A = D.hi >> (22u - n);
  const uint32_t q = A / D.range;
  const uint32_t rem = A - q * D.range;
  const uint32_t h = D.hi & ((1u << (22u - n)) - 1u);
  D.hi = (rem << 22) | cb_funnel_l(D.lo, h, n);   // (h:lo) << n, n in 1..16
  D.lo <<= n;
  D.f += (int32_t)n;
  return q;
}

// The same n bypass decisions (n <= kEpRunMax, f <= 54 - n on entry) as a restoring division on the upper word alone:
// the shifted window of step k compares >= range << 21 exactly when the UNSHIFTED remainder compares >= (range << 21) >> k
// (both sides of the shifted comparison are multiples of 2^k up to the k look-ahead bits shifted in below them), so the
// window is shifted once, by n, at the end.  Two dependent operations per bin and no divide: for the 2 - 4 suffix bits of
// a small exp-Golomb value this is both shorter and far lower in latency than decw_ep_run's quotient.
CB_HD uint32_t decw_ep_bits(DecWide& D, uint32_t n) {
  uint32_t S = D.range << 21, H = D.hi, q = 0;
#pragma unroll 1
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t t = H - S;
    const bool p = H >= S;
    H = p ? t : H;
    q = 2u * q + (p ? 1u : 0u);
    S >>= 1;
  }
  D.hi = cb_funnel_l(D.lo, H, n);   // (H:lo) << n, n in 1..16
  D.lo <<= n;
  D.f += (int32_t)n;
  return q;
}

// The same n bypass decisions (0 <= n <= 13, f <= 54 - n on entry) as ONE division by a reciprocal from a table: the
// quotient of decw_ep_run is floor(A / range) with A = hi >> (22 - n) < range << n < 2^23, and with M = floor(2^32 /
// range) + 1 the product A * M >> 32 equals it exactly (the excess A * (M - 2^32 / range) / 2^32 stays below 2^-9 <
// 1 / range).  rcp = shared-window address of the table M[range & 255], range = 256 .. 511 (a corrupt stream may leave
// a range outside 256 .. 510: the index stays inside the table whatever it is).  n = 0 leaves the window as it is, so
// the call needs no branch around it; device only.
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t decw_ep_recip(DecWide& D, uint32_t n, uint32_t rcp) {
  uint32_t M;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(M) : "r"(rcp + 4u * (D.range & 255u)));
  const uint32_t sh = 22u - n;
  const uint32_t A = D.hi >> sh;
  const uint32_t q = __umulhi(A, M);
  const uint32_t rem = A - q * D.range;
  const uint32_t h = D.hi & ~(0xffffffffu << sh);
  D.hi = (rem << 22) | cb_funnel_l(D.lo, h, n);   // (h:lo) << n
  D.lo <<= n;
  D.f += (int32_t)n;
  return q;
}
#endif

// decodeBinTrm, Decoder.cpp:423-472
CB_HD uint32_t decw_trm(DecWide& D) {
  D.range -= 2;
  const uint32_t scaled = D.range << 22;
  if (D.hi >= scaled) return 1u;
  if (D.range < 256u) {
    D.range <<= 1;
    D.hi = (D.hi << 1) | (D.lo >> 31);
    D.lo <<= 1;
    D.f += 1;
  }
  return 0u;
}

// finish, Decoder.cpp:73-85: terminate bin must be 1 and the unread part of the last byte
// the reference would have read must be the stop bit followed by zeros.
CB_HD uint32_t decw_finish(DecWide& D) {
  const uint32_t t = decw_trm(D);
  const uint32_t s = (uint32_t)(D.f + 8 * (int32_t)D.p - 63);   // total shift count
  const uint32_t idx = 1u + (s >> 3);                           // last byte the reference has read
  const uint32_t last = idx < D.len ? D.in[idx] : 0xffu;
  const uint32_t stop = ((last << (s & 7u)) & 0xffu) == 0x80u;
  return (t == 1u && stop) ? 1u : 0u;
}

// ---------------------------------------------------------------------------
// op blocks: the schedule both kernels and the host emulation follow
// ---------------------------------------------------------------------------
// u8 op format: op = code << 1 | bin, code 0..124 context, 125 terminate, 126 bypass.  A code >= n_ctx that is not the
// terminate code is coded as a bypass bin (defined behaviour for malformed op arrays, the same in every kernel).
constexpr uint32_t kOpTrmCode = 125u;

// the four op codes of a word of ops, one per byte
CB_HD uint32_t op_codes4(uint32_t w) { return (w >> 1) & 0x7f7f7f7fu; }

// true when one of the 16 op codes of a block may be a terminate op (false positives are
// possible -- never false negatives -- and only send the block down the general path)
CB_HD bool block_has_trm(const uint32_t cw[4]) {
  // code bytes are <= 0x7f, so "byte == 0" <=> borrow into bit 7 of (x - 0x01): exact for the
  // lowest zero byte, possibly a false positive above it
  const uint32_t h0 = (cw[0] ^ 0x7d7d7d7du) - 0x01010101u;
  const uint32_t h1 = (cw[1] ^ 0x7d7d7d7du) - 0x01010101u;
  const uint32_t h2 = (cw[2] ^ 0x7d7d7d7du) - 0x01010101u;
  const uint32_t h3 = (cw[3] ^ 0x7d7d7d7du) - 0x01010101u;
  return ((h0 | h1 | h2 | h3) & 0x80808080u) != 0u;
}

// Context storage as the kernels see it: slot c of this lane holds a token, c == n_ctx is the
// bypass slot.  Tab::token(st) = token of a state byte, Tab::row(tok) = its row.  `code` = op >> 1,
// the bin is bit 8*B of `w`.
template <int B, bool TP = false, class Ctx, class Tab>
CB_HD void encw_op(EncWide& E, uint32_t code, uint32_t w, const Ctx& ctx, const Tab& tab, uint32_t n_ctx) {
  const bool is_ep = code >= n_ctx;     // 126, and by definition any code that is not a context of this call (isscabac.h)
  const uint32_t c = code < n_ctx ? code : n_ctx;
  const WRow row = tab.row(ctx.load(c));
  ctx.store_sel(c, encw_bin<B, TP>(E, w, is_ep, row), row.next_lps, row.next_mps);
}

// 16 ops without a terminate op: 4 x (2 bins, guard, 2 bins, voted emit); see kLazy for the
// window budget.  The guard fires for a lane that gained more than 10 bits in two bins.
template <bool VOTE, class Ctx, class Tab>
CB_HD void encw_block16(EncWide& E, const uint32_t w[4], const uint32_t cw[4],
                        const Ctx& ctx, const Tab& tab, uint32_t n_ctx) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t codes = cw[g];   // the four op codes, one per byte
    encw_op<0, true>(E, cb_prmt(codes, 0, 0x4440u), w[g], ctx, tab, n_ctx);
    encw_op<1, true>(E, cb_prmt(codes, 0, 0x4441u), w[g], ctx, tab, n_ctx);
    if (CB_UNLIKELY(E.n >= kLazy)) encw_emit(E);   // the guard: practically never taken, kept off the straight path
    encw_op<2, true>(E, cb_prmt(codes, 0, 0x4442u), w[g], ctx, tab, n_ctx);
    encw_op<3, true>(E, cb_prmt(codes, 0, 0x4443u), w[g], ctx, tab, n_ctx);
    if (cb_any<VOTE>(E.n >= kLazy)) encw_emit(E);
  }
}

// general path: any op kind, any count (stream heads up to the 16-byte boundary, tails,
// blocks with terminate ops)
template <class Ctx, class Tab>
CB_HD void encw_general(EncWide& E, uint32_t o, const Ctx& ctx, const Tab& tab, uint32_t n_ctx) {
  if ((o >> 1) == kOpTrmCode) encw_trm(E, o & 1u);
  else encw_op<0>(E, o >> 1, o, ctx, tab, n_ctx);
  encw_emit(E);
}

template <int B, bool TP = false, class Ctx, class Tab>
CB_HD uint32_t decw_op(DecWide& D, uint32_t code, const Ctx& ctx, const Tab& tab, uint32_t n_ctx) {
  const bool is_ep = code >= n_ctx;
  const uint32_t c = code < n_ctx ? code : n_ctx;
  const WRow row = tab.row(ctx.load(c));
  bool is_lps;
  const uint32_t bin = decw_bin<B, TP>(D, is_ep, is_lps, row);
  ctx.store_sel(c, is_lps, row.next_lps, row.next_mps);
  return bin;
}

// 16 ops without a terminate op -> 16 bins packed one per byte (little-endian in r[0..3]).
// Decisions need f <= 53; same lazy schedule and window budget as the encoder (kLazy).
template <bool VOTE, class Ctx, class Tab>
CB_HD void decw_block16(DecWide& D, const uint32_t cw[4], uint32_t r[4],
                        const Ctx& ctx, const Tab& tab, uint32_t n_ctx) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t codes = cw[g];
    if (VOTE && CABAC_LAZY_DEC && CABAC_REFILL_P == 3 && CABAC_REFILL_FIRST) decw_refill_p<true>(D);   // top-up in FRONT of the group
    uint32_t acc = decw_op<0, true>(D, cb_prmt(codes, 0, 0x4440u), ctx, tab, n_ctx);
    acc |= decw_op<1, true>(D, cb_prmt(codes, 0, 0x4441u), ctx, tab, n_ctx);
    if (CABAC_LAZY_DEC == 1 && D.f >= kLazyDec) decw_refill(D);
    acc |= decw_op<2, true>(D, cb_prmt(codes, 0, 0x4442u), ctx, tab, n_ctx);
    acc |= decw_op<3, true>(D, cb_prmt(codes, 0, 0x4443u), ctx, tab, n_ctx);
    r[g] = acc;
    if (VOTE && CABAC_LAZY_DEC && CABAC_REFILL_P == 3 && CABAC_REFILL_FIRST) {
    } else if (VOTE && CABAC_LAZY_DEC && CABAC_REFILL_P == 3) {
      decw_refill_p<true>(D);                     // no vote: every lane holding 32 unfilled bits tops up, after every group
    } else if (cb_any<(VOTE && CABAC_LAZY_DEC)>(D.f >= kLazyDec)) {
      if (VOTE && CABAC_LAZY_DEC && CABAC_REFILL_P) decw_refill_p<(CABAC_REFILL_P > 1)>(D);   // the warp is here together
      else decw_refill(D);
    }
  }
}

template <class Ctx, class Tab>
CB_HD uint32_t decw_general(DecWide& D, uint32_t o, const Ctx& ctx, const Tab& tab, uint32_t n_ctx) {
  uint32_t bin;
  if ((o >> 1) == kOpTrmCode) bin = decw_trm(D);
  else bin = decw_op<0>(D, o >> 1, ctx, tab, n_ctx);
  decw_refill(D);
  return bin;
}

}  // namespace cabac
