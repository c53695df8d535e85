// bin_emit.cuh -- the per-thread side of the symbol-parallel binarizer's emit phase (k_bin_emit8, symbols.cu), host +
// device so that tests/emul/lane_emul.cpp runs the same code on a whole tile without a GPU.
//
// What it restates: cabacBinarizer.m:30-75 (bin strings) + cabacContextSelection.m:24-67 / cabacDemo.m:113-121 (context of
// every bin) for a RUN of 8 consecutive symbols whose ops are a contiguous byte string of the op array.
//
// A thread owns 8 consecutive symbols, i.e. a byte range [pos, pos + total) of its tile's stage in shared memory, at any
// byte alignment.  It appends its symbols' op strings (<= 7 bytes each per append, out of a table entry or the closed
// form) to a running word in a register and stores every COMPLETED 32-bit word with one aligned store:
//   * the first word it stores has zeros in the bytes below its start -- they belong to the previous thread;
//   * its last, incomplete word is NOT stored in the loop: after the block's barrier the thread stores those (<= 3)
//     bytes one by one (bin_tail) -- byte stores, so neighbours never write the same byte, and after the barrier, so a
//     neighbour's zero bytes cannot land on top of them.
// No atomics (a spread shared-memory atomic costs 2 cycles per LANE on this part), one table load and ~17 straight-line
// instructions per symbol, 2 predicated stores.
#pragma once
#include "cabac_lane.cuh"

namespace cabac {

// Op strings by table: entry = 15 op bytes + the length in byte 15 (k_bin_lut); the emit kernel's fast table keeps the
// strings of <= 7 ops as 7 op bytes + the length in byte 7 (length byte 0: not in the fast table).
constexpr uint32_t LUT_MAX = 1024, LUT_ESC = 0xffffu;
struct LutGeom { uint32_t dom, entries; };
CB_HD LutGeom lut_geom(int profile, int method, uint32_t Nq) {
  if (method == BIN_FL32) return LutGeom{0u, 0u};
  const uint32_t nq = Nq ? Nq : 256u;
  if (profile == PROFILE_ISS) { const uint32_t d = nq < 31u ? nq : 31u; return LutGeom{d, d * (d + 1u)}; }
  const uint32_t d = nq < 256u ? nq : 256u;
  return LutGeom{d, profile == PROFILE_DEMO ? 3u * d : d};
}
// key of a symbol: v = its value, u = the value of its neighbour, has_up = the neighbour exists
CB_HD uint32_t lut_index(const SymCfg& cfg, uint32_t dom, uint32_t v, uint32_t u, bool has_up) {
  if (v >= dom) return LUT_ESC;
  if (cfg.profile == PROFILE_ISS) return (has_up && u >= dom) ? LUT_ESC : v * (dom + 1u) + (has_up ? u + 1u : 0u);
  if (cfg.profile == PROFILE_DEMO) return v + dom * (has_up ? (sym_code(u, cfg.Nq, cfg.method).np > 1u ? 1u : 2u) : 0u);
  return v;
}
// the op string of table entry e as four words (all zero when it has more than 15 ops)
CB_HD void lut_entry(const SymCfg& cfg, uint32_t e, uint32_t w[4]) {
  const LutGeom g = lut_geom(cfg.profile, cfg.method, cfg.Nq);
  uint32_t v, u = 0;
  bool has_up = false;
  SymCode uc = {0, 0, 0};
  if (cfg.profile == PROFILE_ISS) {
    v = e / (g.dom + 1u);
    const uint32_t r = e % (g.dom + 1u);
    has_up = r != 0;
    u = has_up ? r - 1u : 0u;
    uc = sym_code(u, cfg.Nq, cfg.method);
  } else if (cfg.profile == PROFILE_DEMO) {
    v = e % g.dom;
    const uint32_t t = e / g.dom;
    has_up = t != 0;
    uc = SymCode{1u, t == 1 ? 2u : 1u, 0u};      // only the neighbour's first bin matters: 1 (np > 1) or 0 (np == 1)
  } else {
    v = e;
  }
  const SymCode code = sym_code(v, cfg.Nq, cfg.method);
  w[0] = w[1] = w[2] = w[3] = 0u;
  if (code.len <= 15u) {
    for (uint32_t b = 1; b <= code.len; ++b) {
      const int cx = select_ctx(cfg, b, code.np, uc, has_up);
      const uint32_t cd = cx < 0 ? 126u : (uint32_t)cx;
      const uint32_t byte = (cd << 1) | sym_bin(code, b);
      const uint32_t at = b - 1u;
      if (at < 4) w[0] |= byte << (8 * at);
      else if (at < 8) w[1] |= byte << (8 * (at - 4));
      else if (at < 12) w[2] |= byte << (8 * (at - 8));
      else w[3] |= byte << (8 * (at - 12));
    }
    w[3] |= code.len << 24;
  }
}

// ---- the running word ---------------------------------------------------------------------------------------------
// `stage` is a byte address space (shared-window address on the device, an index into a byte array on the host); St
// provides word(addr, w) for an aligned 32-bit store, byte(addr, b) for a byte store and words() (see bin_append).
struct BinAcc {
  uint32_t a0;     // the current, incomplete word: bytes [0, fb / 8) are set
  uint32_t fb;     // bits of it in use: 0, 8, 16 or 24
  uint32_t wp;     // its address (4-byte aligned)
};
CB_HD void bin_acc_start(BinAcc& A, uint32_t stage_addr_of_first_byte) {
  A.a0 = 0u;
  A.fb = (stage_addr_of_first_byte & 3u) * 8u;
  A.wp = stage_addr_of_first_byte & ~3u;
}
CB_HD uint32_t bin_shl_hi(uint32_t lo, uint32_t hi, uint32_t s) {      // bits 32..63 of (hi:lo) << s, s in 0..24
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  return (uint32_t)(((((uint64_t)hi) << 32 | lo) << s) >> 32);
#endif
}
// append bits / 8 (0..7) bytes: d_lo = bytes 0..3, d_hi = bytes 4..6 (byte 3 of d_hi and every byte from the length on are
// zero); `bits` = 8 x the length
template <class St>
CB_HD void bin_append(BinAcc& A, St& st, uint32_t d_lo, uint32_t d_hi, uint32_t bits) {
  const uint32_t w0 = A.a0 | (d_lo << A.fb);
  const uint32_t w1 = bin_shl_hi(d_lo, d_hi, A.fb);
  const uint32_t w2 = bin_shl_hi(d_hi, 0u, A.fb);
  const uint32_t nfb = A.fb + bits;          // <= 24 + 56: 0, 1 or 2 words are complete
  st.words(A, w0, w1, w2, nfb);              // stores w0 if nfb >= 32 and w1 if nfb >= 64; new running word and its address
  A.fb = nfb & 31u;
}
// what St::words does, in plain C++ (the device's stage does it with predicated stores, selects and adds: the compiler
// turns the ifs into divergent branches, and a branch costs a warp more than the whole append)
template <class St>
CB_HD void bin_words_plain(St& st, BinAcc& A, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t nfb) {
  if (nfb >= 32u) st.word(A.wp, w0);
  if (nfb >= 64u) st.word(A.wp + 4u, w1);
  A.a0 = nfb >= 64u ? w2 : (nfb >= 32u ? w1 : w0);
  A.wp += (nfb >> 5) * 4u;
}
// a symbol that is not in the fast table: the 16-byte entry `q` (strings of 8..15 ops) in pieces of 7, or -- q.w's length
// byte 0 -- the closed form, 7 ops per append
// (by value in, by value out: a reference to the caller's running word would pin it in local memory for the hot path too)
template <class St>
CB_HD_NOINLINE BinAcc bin_append_long(BinAcc A, St st, SymCfg cfg, uint32_t qx, uint32_t qy, uint32_t qz, uint32_t qw,
                                      uint32_t v, uint32_t u, bool up) {
  const uint32_t tl = qw >> 24;
  if (tl) {
    bin_append(A, st, qx, qy & 0x00ffffffu, 8u * (tl < 7u ? tl : 7u));
    if (tl > 7u) bin_append(A, st, (qy >> 24) | (qz << 8), ((qz >> 24) | (qw << 8)) & 0x00ffffffu, 8u * (tl - 7u < 7u ? tl - 7u : 7u));
    if (tl > 14u) bin_append(A, st, (qw >> 16) & 0xffu, 0u, 8u);
    return A;
  }
  const SymCode code = sym_code(v, cfg.Nq, cfg.method), prev = sym_code(u, cfg.Nq, cfg.method);
  for (uint32_t j0 = 0; j0 < code.len; j0 += 7u) {
    const uint32_t m = code.len - j0 < 7u ? code.len - j0 : 7u;
    uint32_t d_lo = 0u, d_hi = 0u;
    for (uint32_t j = 0; j < m; ++j) {
      const int cx = select_ctx(cfg, j0 + j + 1u, code.np, prev, up);
      const uint32_t byte = ((cx < 0 ? 126u : (uint32_t)cx) << 1) | sym_bin(code, j0 + j + 1u);
      if (j < 4u) d_lo |= byte << (8u * j); else d_hi |= byte << (8u * (j - 4u));
    }
    bin_append(A, st, d_lo, d_hi, 8u * m);
  }
  return A;
}
// the bytes of the last, incomplete word; wp_first = the thread's first word address, fb_first = bits of that word that
// belong to the predecessor (a thread that never completed a word must not touch them)
template <class St>
CB_HD void bin_tail(const BinAcc& A, St& st, uint32_t wp_first, uint32_t fb_first) {
  const uint32_t from = A.wp == wp_first ? fb_first : 0u;
  if (from < 8u && A.fb > 0u) st.byte(A.wp, A.a0 & 0xffu);
  if (from < 16u && A.fb > 8u) st.byte(A.wp + 1u, (A.a0 >> 8) & 0xffu);
  if (from < 24u && A.fb > 16u) st.byte(A.wp + 2u, (A.a0 >> 16) & 0xffu);
}

// fast-table entry from the full one: 7 op bytes + 8 x the length in byte 7; (0, 0) when the string has 0 (not in any
// table) or more than 7 ops
CB_HD void lut8_from16(uint32_t qx, uint32_t qy, uint32_t qw, uint32_t& lo, uint32_t& hi) {
  const uint32_t tl = qw >> 24;
  const bool ok = tl >= 1u && tl <= 7u;
  lo = ok ? qx : 0u;
  hi = ok ? ((qy & 0x00ffffffu) | (tl << 27)) : 0u;
}

}  // namespace cabac
