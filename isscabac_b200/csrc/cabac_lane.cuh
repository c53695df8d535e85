// cabac_lane.cuh -- per-lane (one thread = one CABAC stream) coder logic.
//
// Everything here is `__host__ __device__` so that the exact code the kernels run
// can also be compiled by g++ for the host-side emulation test
// (tests/emul/lane_emul.cpp); the product only ever runs it on the GPU.
//
// Differences from the reference's formulation (results are byte-identical):
//  * output bytes are written EAGERLY and a carry is propagated by walking back
//    over bytes already written, instead of the reference's bufferedByte /
//    numBufferedBytes deferral (CABAC_ArithmeticEncoder.cpp:380-412).  Both produce
//    "the lead bytes in order, with every carry added into the preceding bytes".
//  * the MPS and LPS paths of encodeBin/decodeBin are one branch-free sequence:
//    renorm shift = min(clz(r)-23, 6) of the selected sub-range r covers the
//    32-entry renorm table (Encoder.cpp:482-492), the MPS "range < 256" test
//    (:145-165) and the state-63 row in one expression.
//  * the LPS range table (Encoder.cpp:414-480) and both state-transition tables
//    (ContextModel.cpp:136-158) are fused into one 8-byte row per state byte.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#define CB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define CB_HD inline
#define CB_HD_NOINLINE static inline
struct uint2 { uint32_t x, y; };
#endif

namespace cabac {

// ---------------------------------------------------------------------------
// intrinsics with host fall-backs (host versions exist only for the emulation test)
// ---------------------------------------------------------------------------
CB_HD int cb_clz(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
// result byte i = byte sel_nibble[i] of the 8-byte pool {a (0..3), b (4..7)}
CB_HD uint32_t cb_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  uint64_t pool = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
#endif
}
// low 32 bits of ((hi:lo) >> s), 0 <= s < 32
CB_HD uint32_t cb_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> s);
#endif
}

// cb_perm without the selector sanitising of __byte_perm: every selector nibble must be 0..7
CB_HD uint32_t cb_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
#else
  return cb_perm(a, b, sel);
#endif
}
// makes a value opaque to the optimiser (keeps an address in registers instead of having it
// recomputed from the kernel parameters at every use)
template <class T>
CB_HD T* cb_keep(T* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+l"(p));
#endif
  return p;
}
// 32-bit store to global memory (keeps the STG form when the pointer came through cb_keep)
CB_HD void cb_stg32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  *p = v;
#endif
}
// (a ^ b) & 1 as ONE three-input logic op
CB_HD uint32_t cb_xor_and1(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, 1, 0x28;" : "=r"(d) : "r"(a), "r"(b));
  return d;
#else
  return (a ^ b) & 1u;
#endif
}
// (a ^ b) & mask as ONE three-input logic op whose result the optimiser cannot see through (it
// would otherwise turn the selects that test it into shift/mask arithmetic); ptxas folds the
// "!= 0" test of the caller into the predicate output of the same LOP3
template <uint32_t MASK>
CB_HD uint32_t cb_xor_and(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(d) : "r"(a), "r"(b), "n"(MASK));
  return d;
#else
  return (a ^ b) & MASK;
#endif
}
CB_HD uint32_t cb_keep32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(v));
#endif
  return v;
}
// acc + a * b with a 32x32 -> 64 bit product (one IMAD.WIDE)
CB_HD uint64_t cb_mad_wide(uint32_t a, uint32_t b, uint64_t acc) {
#if defined(__CUDA_ARCH__)
  uint64_t d;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(acc));
  return d;
#else
  return acc + (uint64_t)a * b;
#endif
}
// upper 32 bits of ((hi:lo) << s), 0 <= s < 32
CB_HD uint32_t cb_funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  return (uint32_t)((((((uint64_t)hi) << 32) | lo) << s) >> 32);
#endif
}

// ---------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------
// rangeTabLps of H.264/HEVC as kept in CABAC_ArithmeticEncoder.cpp:414-480; one
// little-endian word per state: byte q = LPS sub-range for range quartile q.
struct Tables {
  uint32_t lps[64];
  uint8_t trans_lps[64];
};
CB_HD constexpr uint32_t pk(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return a | (b << 8) | (c << 16) | (d << 24);
}
constexpr Tables kTables = {
    {pk(128, 176, 208, 240), pk(128, 167, 197, 227), pk(128, 158, 187, 216), pk(123, 150, 178, 205),
     pk(116, 142, 169, 195), pk(111, 135, 160, 185), pk(105, 128, 152, 175), pk(100, 122, 144, 166),
     pk(95, 116, 137, 158),  pk(90, 110, 130, 150),  pk(85, 104, 123, 142),  pk(81, 99, 117, 135),
     pk(77, 94, 111, 128),   pk(73, 89, 105, 122),   pk(69, 85, 100, 116),   pk(66, 80, 95, 110),
     pk(62, 76, 90, 104),    pk(59, 72, 86, 99),     pk(56, 69, 81, 94),     pk(53, 65, 77, 89),
     pk(51, 62, 73, 85),     pk(48, 59, 69, 80),     pk(46, 56, 66, 76),     pk(43, 53, 63, 72),
     pk(41, 50, 59, 69),     pk(39, 48, 56, 65),     pk(37, 45, 54, 62),     pk(35, 43, 51, 59),
     pk(33, 41, 48, 56),     pk(32, 39, 46, 53),     pk(30, 37, 43, 50),     pk(29, 35, 41, 48),
     pk(27, 33, 39, 45),     pk(26, 31, 37, 43),     pk(24, 30, 35, 41),     pk(23, 28, 33, 39),
     pk(22, 27, 32, 37),     pk(21, 26, 30, 35),     pk(20, 24, 29, 33),     pk(19, 23, 27, 31),
     pk(18, 22, 26, 30),     pk(17, 21, 25, 28),     pk(16, 20, 23, 27),     pk(15, 19, 22, 25),
     pk(14, 18, 21, 24),     pk(14, 17, 20, 23),     pk(13, 16, 19, 22),     pk(12, 15, 18, 21),
     pk(12, 14, 17, 20),     pk(11, 14, 16, 19),     pk(11, 13, 15, 18),     pk(10, 12, 15, 17),
     pk(10, 12, 14, 16),     pk(9, 11, 13, 15),      pk(9, 11, 12, 14),      pk(8, 10, 12, 14),
     pk(8, 9, 11, 13),       pk(7, 9, 11, 12),       pk(7, 9, 10, 12),       pk(7, 8, 10, 11),
     pk(6, 8, 9, 11),        pk(6, 7, 9, 10),        pk(6, 7, 8, 9),         pk(2, 2, 2, 2)},
    // transIdxLps; ContextModel.cpp:148-158 is this table expanded over (state<<1)+mps
    {0,  0,  1,  2,  2,  4,  4,  5,  6,  7,  8,  9,  9,  11, 11, 12, 13, 13, 15, 15, 16, 16,
     18, 18, 19, 19, 21, 21, 22, 22, 23, 24, 24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30,
     31, 32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 63}};

// Fused row for a context state byte st = (state<<1)|mps:
//   x = the four LPS sub-ranges of `state`
//   y = nextStateMPS(st) | nextStateLPS(st) << 8        (ContextModel.cpp:136-158)
CB_HD constexpr uint32_t next_mps(uint32_t st) { return st < 124 ? st + 2 : st; }
CB_HD constexpr uint32_t next_lps(uint32_t st) {
  return (st >> 1) == 0 ? 1u - (st & 1u) : (((uint32_t)kTables.trans_lps[st >> 1]) << 1) + (st & 1u);
}
CB_HD constexpr uint2 fused_row(uint32_t st) {
  return uint2{kTables.lps[st >> 1], next_mps(st) | (next_lps(st) << 8)};
}

// ---------------------------------------------------------------------------
// encoder lane
// ---------------------------------------------------------------------------
struct EncLane {
  uint32_t low, range;
  int32_t bits_left;
  uint32_t acc;       // up to 3 pending output bytes, newest in the top byte
  uint32_t nbytes;    // bytes produced so far (pending ones included)
  uint32_t nbuf;      // the reference's m_numBufferedBytes (kept for getNumBits semantics)
  uint32_t overflow;  // set when the slab is too small
  uint32_t cap;       // slab capacity in bytes (multiple of 4)
  uint8_t* out;       // slab base, 4-byte aligned
};

// CABAC_ArithmeticEncoder.cpp:54-61
CB_HD void enc_start(EncLane& L, uint8_t* out, uint32_t cap) {
  L.low = 0; L.range = 510; L.bits_left = 23;
  L.acc = 0; L.nbytes = 0; L.nbuf = 0; L.overflow = 0;
  L.cap = cap; L.out = out;
}

CB_HD void enc_put_byte(EncLane& L, uint32_t b) {
  L.acc = cb_funnel_r(L.acc, b, 8);
  L.nbytes++;
  if ((L.nbytes & 3u) == 0) {
    if (L.nbytes <= L.cap) *reinterpret_cast<uint32_t*>(L.out + (L.nbytes - 4)) = L.acc;
    else L.overflow = 1;
  }
}

// +1 on the big-endian number formed by all bytes produced so far: first through
// the pending bytes in `acc`, then (rarely) back over bytes already in the slab.
// Replaces the carry handling of writeOut()/finish() (Encoder.cpp:76-87,394-404).
// Takes and returns plain values so the lane state stays in registers around the call.
CB_HD_NOINLINE uint32_t enc_carry_acc(uint32_t acc, uint32_t nbytes, uint8_t* out, uint32_t overflow) {
  uint32_t p = nbytes & 3u;
  uint32_t r = cb_perm(acc, 0, 0x0123);  // byte-reverse: newest pending byte lowest
  uint32_t mask = p ? (0xffffffffu >> (32 - 8 * p)) : 0u;
  uint32_t pend = (r & mask) + 1u;
  uint32_t carry = pend >> (8 * p);
  r = (r & ~mask) | (pend & mask);
  if (carry && !overflow) {
    volatile uint8_t* o = out;
    for (int64_t i = (int64_t)(nbytes - p) - 1; i >= 0; --i) {
      uint32_t b = o[i];
      if (b == 0xff) { o[i] = 0; continue; }
      o[i] = (uint8_t)(b + 1);
      break;
    }
  }
  return cb_perm(r, 0, 0x0123);
}
CB_HD void enc_carry(EncLane& L) { L.acc = enc_carry_acc(L.acc, L.nbytes, L.out, L.overflow); }

// testAndWriteOut + writeOut, Encoder.cpp:369-412; caller checks bits_left < 12
template <bool TRACK>
CB_HD void enc_write_out(EncLane& L) {
  uint32_t lead = L.low >> (24 - L.bits_left);  // 9 bits, bit 8 = carry
  L.bits_left += 8;
  L.low &= 0xffffffffu >> L.bits_left;
  if (lead > 0xffu) enc_carry(L);
  if (TRACK) L.nbuf = (lead == 0xffu) ? L.nbuf + 1u : 1u;
  enc_put_byte(L, lead);
}

// encodeBin, Encoder.cpp:113-178.  st = context state byte, row = fused_row(st).
template <bool TRACK>
CB_HD void enc_bin_ctx(EncLane& L, uint32_t bin, uint32_t& st, uint2 row) {
  uint32_t lps = cb_perm(0, row.x, L.range >> 6);  // range in [256,510] -> selector 4..7 = row.x byte q
  uint32_t rmps = L.range - lps;
  uint32_t is_lps = (st ^ bin) & 1u;
  uint32_t rsel = is_lps ? lps : rmps;
  int n = cb_clz(rsel) - 23;
  n = n > 6 ? 6 : n;
  L.low = (L.low + (is_lps ? rmps : 0u)) << n;
  L.range = rsel << n;
  L.bits_left -= n;
  st = cb_perm(row.y, 0, is_lps | 0x4440u);
  if (L.bits_left < 12) enc_write_out<TRACK>(L);
}

// encodeBinEP, Encoder.cpp:250-270
template <bool TRACK>
CB_HD void enc_bin_ep(EncLane& L, uint32_t bin) {
  L.low = (L.low << 1) + (bin ? L.range : 0u);
  L.bits_left -= 1;
  if (L.bits_left < 12) enc_write_out<TRACK>(L);
}

// encodeBinTrm, Encoder.cpp:326-367
template <bool TRACK>
CB_HD void enc_bin_trm(EncLane& L, uint32_t bin) {
  L.range -= 2;
  if (bin) {
    L.low = (L.low + L.range) << 7;
    L.range = 256;
    L.bits_left -= 7;
  } else if (L.range < 256) {
    L.low <<= 1; L.range <<= 1; L.bits_left -= 1;
  }
  if (L.bits_left < 12) enc_write_out<TRACK>(L);
}

// finish, Encoder.cpp:70-105 (terminate bin, carry resolution, 24-bitsLeft tail
// bits, stop bit, zero padding) followed by flushing the pending bytes.
template <bool TRACK>
CB_HD void enc_finish(EncLane& L) {
  enc_bin_trm<TRACK>(L, 1);
  if (L.low >> (32 - L.bits_left)) {
    enc_carry(L);
    L.low -= 1u << (32 - L.bits_left);
  }
  int tb = 24 - L.bits_left + 1;             // tail bits incl. the stop bit: 6..13
  uint32_t v = ((L.low >> 8) << 1) | 1u;
  if (tb <= 8) {
    enc_put_byte(L, (v << (8 - tb)) & 0xffu);
  } else {
    uint32_t w = v << (16 - tb);
    enc_put_byte(L, (w >> 8) & 0xffu);
    enc_put_byte(L, w & 0xffu);
  }
  L.nbuf = 0;
}

// write the (<4) pending bytes; the lane can continue afterwards (acc is kept)
CB_HD void enc_flush_pending(EncLane& L) {
  uint32_t p = L.nbytes & 3u;
  for (uint32_t j = 0; j < p; ++j) {
    uint32_t idx = L.nbytes - p + j;
    if (idx < L.cap) L.out[idx] = (uint8_t)(L.acc >> (8 * (4 - p + j)));
    else L.overflow = 1;
  }
}

// getNumberOfWrittenBits() while coding: bytes already handed to the sink, i.e.
// everything except the reference's buffered bytes (CABAC_BitstreamFile.h:70)
CB_HD uint32_t enc_bits_written(const EncLane& L) { return 8u * (L.nbytes - L.nbuf); }

// ---------------------------------------------------------------------------
// decoder lane
// ---------------------------------------------------------------------------
struct DecLane {
  uint32_t value, range;
  int32_t bits_needed;
  uint32_t pos, len, last;
  const uint8_t* in;
};

// readByte, CABAC_BitstreamFile.cpp:153-158 (0xFF past the end)
CB_HD uint32_t dec_read(DecLane& D) {
  uint32_t b = D.pos < D.len ? (uint32_t)D.in[D.pos] : 0xffu;
  D.pos++;
  D.last = b;
  return b;
}

// start, CABAC_ArithmeticDecoder.cpp:54-60
CB_HD void dec_start(DecLane& D, const uint8_t* in, uint32_t len) {
  D.in = in; D.len = len; D.pos = 0; D.last = 0;
  D.range = 510; D.bits_needed = -8;
  D.value = dec_read(D) << 8;
  D.value |= dec_read(D);
}

// decodeBin, Decoder.cpp:87-190 (MPS and LPS paths merged, see file header)
CB_HD uint32_t dec_bin_ctx(DecLane& D, uint32_t& st, uint2 row) {
  uint32_t lps = cb_perm(0, row.x, D.range >> 6);
  uint32_t rmps = D.range - lps;
  uint32_t scaled = rmps << 7;
  uint32_t is_lps = D.value >= scaled ? 1u : 0u;
  uint32_t rsel = is_lps ? lps : rmps;
  int n = cb_clz(rsel) - 23;
  n = n > 6 ? 6 : n;
  D.value = (D.value - (is_lps ? scaled : 0u)) << n;
  D.range = rsel << n;
  uint32_t bin = (st ^ is_lps) & 1u;
  st = cb_perm(row.y, 0, is_lps | 0x4440u);
  D.bits_needed += n;
  if (D.bits_needed >= 0) {
    D.value += dec_read(D) << D.bits_needed;
    D.bits_needed -= 8;
  }
  return bin;
}

// decodeBinEP, Decoder.cpp:288-331
CB_HD uint32_t dec_bin_ep(DecLane& D) {
  D.value <<= 1;
  if (++D.bits_needed >= 0) {
    D.bits_needed = -8;
    D.value += dec_read(D);
  }
  uint32_t scaled = D.range << 7;
  uint32_t bin = D.value >= scaled ? 1u : 0u;
  D.value -= bin ? scaled : 0u;
  return bin;
}

// decodeBinTrm, Decoder.cpp:423-472
CB_HD uint32_t dec_bin_trm(DecLane& D) {
  D.range -= 2;
  uint32_t scaled = D.range << 7;
  if (D.value >= scaled) return 1u;
  if (scaled < (256u << 7)) {
    D.range = scaled >> 6;
    D.value <<= 1;
    if (++D.bits_needed == 0) {
      D.bits_needed = -8;
      D.value += dec_read(D);
    }
  }
  return 0u;
}

// finish, Decoder.cpp:73-85 with the two asserts as a result
CB_HD uint32_t dec_finish(DecLane& D) {
  uint32_t t = dec_bin_trm(D);
  uint32_t stop = ((D.last << (8 + D.bits_needed)) & 0xffu) == 0x80u;
  return (t == 1u && stop) ? 1u : 0u;
}

// ---------------------------------------------------------------------------
// binarization (cabacBinarizer.m:30-75) in closed form
// ---------------------------------------------------------------------------
enum { BIN_TU = 0, BIN_EG0 = 1, BIN_EG1 = 2, BIN_EG2 = 3, BIN_FL32 = 4, BIN_TR0 = 5, BIN_TR1 = 6, BIN_TR2 = 7 };
enum { PROFILE_DEMO = 0, PROFILE_ISS = 1, PROFILE_FLAT = 2, PROFILE_FLAT_EPSUF = 3 };
enum { CM_COND0 = 1, CM_COND1 = 2, CM_CONDBINLFT = 4, CM_CONDS0 = 8, CM_CONDS1 = 16 };

// A symbol's bin string without materialising it: bins 1..np-1 are 1, bin np is 0
// (np = len+1 when the string has no 0), bins np+1..len are the low len-np bits of suf.
struct SymCode {
  uint32_t len, np, suf;
};
CB_HD uint32_t sym_bin(const SymCode& c, uint32_t n) {
  return n < c.np ? 1u : (n == c.np ? 0u : ((c.suf >> (c.len - n)) & 1u));
}
CB_HD SymCode sym_code(uint32_t v, uint32_t Nq, int method) {
  SymCode c;
  if (method == BIN_TU) {              // cabacBinarizer.m:30-37
    bool top = (v == Nq - 1u);
    c.len = top ? v : v + 1u;
    c.np = v + 1u;                     // == len+1 when there is no terminating zero
    c.suf = 0;
  } else if (method == BIN_FL32) {     // cabacBinarizer.m:71-75
    c.len = 32;
    c.np = (uint32_t)cb_clz(~v) + 1u;  // 33 when v is all ones
    c.suf = v;
  } else if (method >= BIN_TR0) {      // truncated Rice, cabacBinarizer.m:39-54: v >> k ones, a zero, k suffix bits
    const uint32_t k = (uint32_t)(method - BIN_TR0);   // (encode only upstream: the decode loops have no case for it)
    c.np = (v >> k) + 1u;
    c.len = c.np + k;
    // the escape for v >= maxVal is a TODO upstream (:47-50): the suffix bits stay ones there
    c.suf = v >= Nq - 1u ? (1u << k) - 1u : v - ((c.np - 1u) << k);
  } else {                             // EG-k, cabacBinarizer.m:56-69
    uint32_t k = (uint32_t)(method - BIN_EG0);
    uint64_t t = ((uint64_t)v >> k) + 1u;
    uint32_t np = (t >> 32) ? 33u : (uint32_t)(32 - cb_clz((uint32_t)t));
    uint32_t ns = k + np - 1u;
    c.np = np;
    c.len = np + ns;
    c.suf = (uint32_t)((uint64_t)v - (((uint64_t)1 << k) * ((((uint64_t)1) << (np - 1)) - 1u)));
  }
  return c;
}

struct SymCfg {
  int profile, method;
  uint32_t Nq;
  int Nlbp;
  uint32_t types, rows;
};

// Context for bin n (1-based) of a symbol with code c; u = code of the up neighbour
// (ISS, cabacContextSelection.m:24-67) or of the previous symbol (DEMO,
// cabacDemo.m:113-121), has_up = it exists.  Returns the 0-based engine context
// (MATLAB ctxID-1) or -1 for a bypass bin.  `own_np` is the position of the first 0
// among the symbol's own bins (the decoder passes 0xffffffff while still in the prefix).
CB_HD int select_ctx(const SymCfg& cfg, uint32_t n, uint32_t own_np, const SymCode& u, bool has_up) {
  const int N = cfg.Nlbp;
  if (cfg.profile == PROFILE_DEMO) {
    if (n == 1 && has_up) return sym_bin(u, 1) ? 1 : 2;
    return 0;
  }
  bool in_prefix = n <= own_np;
  if (cfg.profile != PROFILE_ISS) {
    if (in_prefix) return (int)n <= N ? (int)n - 1 : N;
    if (cfg.profile == PROFILE_FLAT_EPSUF) return -1;
    int m = (int)(n - own_np);
    return m <= N ? N + m : 2 * N + 1;
  }
  bool up_has = has_up && u.len >= n;
  bool up_in_prefix = n <= u.np;
  int id;
  if (in_prefix) {
    if ((int)n <= N) {
      id = (int)n;
      if (up_has && up_in_prefix) {
        uint32_t ub = sym_bin(u, n);
        if (ub == 0 && (cfg.types & CM_COND0)) id = N + (int)n;
        else if (ub == 1 && (cfg.types & CM_COND1)) id = 2 * N + (int)n;
      } else if (n > 1 && (cfg.types & CM_CONDBINLFT)) {
        id = 3 * N + (int)n - 1;   // own bin n-1 is 1 by construction inside the prefix
      }
    } else {
      id = 7 * N + 1;
    }
  } else {
    int m = (int)(n - own_np);
    if (m <= N) {
      id = 4 * N + m;
      if (up_has && !up_in_prefix) {
        uint32_t ub = sym_bin(u, n);
        if (ub == 0 && (cfg.types & CM_CONDS0)) id = 5 * N + m;
        else if (ub == 1 && (cfg.types & CM_CONDS1)) id = 6 * N + m;
      }
    } else {
      id = 7 * N + 2;
    }
  }
  return id - 1;
}

CB_HD bool sym_has_up(const SymCfg& cfg, uint64_t i) {
  if (cfg.profile == PROFILE_ISS) return cfg.rows ? (i % cfg.rows) != 0 : i > 0;  // cabacEncode.m:52
  if (cfg.profile == PROFILE_DEMO) return i > 0;                                   // cabacDemo.m:105
  return false;
}

// Incremental symbol decoder state: finish detector (cabacDecodeSymbolFinished.m:10-32)
// and debinarizer (cabacDebinarizer.m:28-57) folded into one running state.
struct SymDec {
  uint32_t n;        // bins decoded so far
  uint32_t np;       // position of the first 0, 0xffffffff while in the prefix
  uint32_t ns_left;  // suffix bins still to come
  uint32_t suf;      // suffix value so far
};
CB_HD void symdec_reset(SymDec& s) { s.n = 0; s.np = 0xffffffffu; s.ns_left = 0; s.suf = 0; }
// feed one decoded bin; returns true when the symbol is complete, value in v
CB_HD bool symdec_push(SymDec& s, uint32_t bin, const SymCfg& cfg, uint32_t& v) {
  s.n++;
  if (cfg.method == BIN_TU) {
    if (bin == 0) { v = s.n - 1u; return true; }
    if (s.n == cfg.Nq - 1u) { v = cfg.Nq - 1u; return true; }
    return false;
  }
  if (cfg.method == BIN_FL32) {
    s.suf = (s.suf << 1) | bin;
    if (bin == 0 && s.np == 0xffffffffu) s.np = s.n;
    if (s.n == 32) { v = s.suf; return true; }
    return false;
  }
  uint32_t k = (uint32_t)(cfg.method - BIN_EG0);
  if (s.np == 0xffffffffu) {
    if (bin == 0) {
      s.np = s.n;
      s.ns_left = k + s.np - 1u;
      if (s.ns_left == 0) { v = 0; return true; }   // np == 1 and k == 0
    }
    return false;
  }
  s.suf = (s.suf << 1) | bin;
  if (--s.ns_left == 0) {
    v = (uint32_t)((((uint64_t)1 << k) * ((((uint64_t)1) << (s.np - 1)) - 1u)) + s.suf);
    return true;
  }
  return false;
}

}  // namespace cabac
