/* TEST INFRASTRUCTURE ONLY -- see cabac_oracle.h.  CPU restatement of the
 * reference CABAC path; never used by the product. */
#define _GNU_SOURCE
#include "cabac_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== *
 *  Tables                                                                  *
 * ======================================================================== */

/* LPS sub-range by (probability state, range quartile): the H.264/HEVC
 * rangeTabLps as held in CABAC_ArithmeticEncoder.cpp:414-480 (identical copy in
 * CABAC_ArithmeticDecoder.cpp:474-540).  Row 63 is the terminate pseudo-state. */
static const uint8_t k_lps[64][4] = {
  {128,176,208,240},{128,167,197,227},{128,158,187,216},{123,150,178,205},{116,142,169,195},
  {111,135,160,185},{105,128,152,175},{100,122,144,166},{ 95,116,137,158},{ 90,110,130,150},
  { 85,104,123,142},{ 81, 99,117,135},{ 77, 94,111,128},{ 73, 89,105,122},{ 69, 85,100,116},
  { 66, 80, 95,110},{ 62, 76, 90,104},{ 59, 72, 86, 99},{ 56, 69, 81, 94},{ 53, 65, 77, 89},
  { 51, 62, 73, 85},{ 48, 59, 69, 80},{ 46, 56, 66, 76},{ 43, 53, 63, 72},{ 41, 50, 59, 69},
  { 39, 48, 56, 65},{ 37, 45, 54, 62},{ 35, 43, 51, 59},{ 33, 41, 48, 56},{ 32, 39, 46, 53},
  { 30, 37, 43, 50},{ 29, 35, 41, 48},{ 27, 33, 39, 45},{ 26, 31, 37, 43},{ 24, 30, 35, 41},
  { 23, 28, 33, 39},{ 22, 27, 32, 37},{ 21, 26, 30, 35},{ 20, 24, 29, 33},{ 19, 23, 27, 31},
  { 18, 22, 26, 30},{ 17, 21, 25, 28},{ 16, 20, 23, 27},{ 15, 19, 22, 25},{ 14, 18, 21, 24},
  { 14, 17, 20, 23},{ 13, 16, 19, 22},{ 12, 15, 18, 21},{ 12, 14, 17, 20},{ 11, 14, 16, 19},
  { 11, 13, 15, 18},{ 10, 12, 15, 17},{ 10, 12, 14, 16},{  9, 11, 13, 15},{  9, 11, 12, 14},
  {  8, 10, 12, 14},{  8,  9, 11, 13},{  7,  9, 11, 12},{  7,  9, 10, 12},{  7,  8, 10, 11},
  {  6,  8,  9, 11},{  6,  7,  9, 10},{  6,  7,  8,  9},{  2,  2,  2,  2}};

/* LPS state transition (transIdxLps of the standard).  ContextModel.cpp:148-158
 * holds it expanded over the (state<<1)+mps byte; the expansion is
 * next = (trans[state]<<1)+mps except at state 0 where the MPS flips. */
static const uint8_t k_trans_lps[64] = {
   0, 0, 1, 2, 2, 4, 4, 5, 6, 7, 8, 9, 9,11,11,12,13,13,15,15,16,16,18,18,19,19,21,21,22,22,23,24,
  24,25,26,26,27,27,28,29,29,30,30,30,31,32,32,33,33,33,34,34,35,35,35,36,36,36,37,37,37,38,38,63};

unsigned orc_lps_range(unsigned state, unsigned quartile) { return k_lps[state & 63][quartile & 3]; }

/* CABAC_ArithmeticEncoder.cpp:482-492: table indexed by lps>>3; equals
 * min(clz32(lps)-23, 6) -- NOT plain clz (lps=2 at state 63 gives 6, not 7). */
unsigned orc_renorm_shift(unsigned lps) {
  unsigned idx = lps >> 3, n = 1;
  if (idx == 0) return 6;
  while ((idx << n) < 32u) ++n; /* smallest n with idx<<n >= 32 ... */
  return n;                    /* ... 1 for idx>=16, 2 for 8..15, 3 for 4..7, 4 for 2..3, 5 for 1 */
}

/* ContextModel.h:78-80,101 / ContextModel.cpp:70-74 */
uint8_t orc_ctx_make(unsigned mps, unsigned state) { return (uint8_t)((state << 1) + mps); }

/* ContextModel.cpp:116-123,136-146: +1 state while state < 62 */
uint8_t orc_ctx_next_mps(uint8_t s) { return (uint8_t)(s < 124 ? s + 2 : s); }

/* ContextModel.cpp:107-114,148-158 */
uint8_t orc_ctx_next_lps(uint8_t s) {
  unsigned st = s >> 1, mps = s & 1u;
  if (st == 0) return (uint8_t)(1u - mps);
  return (uint8_t)((k_trans_lps[st] << 1) + mps);
}

/* ======================================================================== *
 *  Bit sink  (CABAC_BitstreamFile.cpp:95-151, .h:70)                       *
 * ======================================================================== */

static void sink_byte(orc_encoder* e, unsigned b) {
  if (e->n_out < e->cap) e->out[e->n_out] = (uint8_t)b;
  e->n_out++;
  e->bits_written += 8;
}

/* append the n low bits of v, MSB first; <8 leftover bits stay held */
static void sink_bits(orc_encoder* e, uint32_t v, unsigned n) {
  while (n > 0) {
    unsigned room = 8 - e->n_held, take = n < room ? n : room;
    uint32_t chunk = (n == 32 && take == 32) ? v : ((v >> (n - take)) & ((1u << take) - 1u));
    e->held = (e->held << take) | chunk;
    e->n_held += take;
    n -= take;
    if (e->n_held == 8) { sink_byte(e, e->held & 0xffu); e->held = 0; e->n_held = 0; }
  }
}

static void sink_align_zero(orc_encoder* e) {
  if (e->n_held == 0) return;
  sink_byte(e, (e->held << (8 - e->n_held)) & 0xffu);
  e->held = 0; e->n_held = 0;
}

uint64_t orc_enc_num_bits(const orc_encoder* e) { return e->bits_written + e->n_held; }

/* ======================================================================== *
 *  Encoder                                                                  *
 * ======================================================================== */

void orc_enc_attach(orc_encoder* e, uint8_t* out, uint64_t cap) {
  memset(e, 0, sizeof *e);
  e->out = out; e->cap = cap;
}

/* CABAC_ArithmeticEncoder.cpp:54-61 */
void orc_enc_start(orc_encoder* e) {
  e->low = 0; e->range = 510; e->bits_left = 23;
  e->num_buffered = 0; e->buffered_byte = 0xff; e->bins_coded = 0;
}

/* CABAC_ArithmeticEncoder.cpp:369-412 (testAndWriteOut + writeOut) */
static void enc_flush_if_needed(orc_encoder* e) {
  if (e->bits_left >= 12) return;
  uint32_t lead = e->low >> (24 - e->bits_left); /* 9 bits: bit 8 is the carry */
  e->bits_left += 8;
  e->low &= 0xffffffffu >> e->bits_left;
  if (lead == 0xff) { e->num_buffered++; return; }
  if (e->num_buffered == 0) { e->num_buffered = 1; e->buffered_byte = lead; return; }
  uint32_t carry = lead >> 8;
  sink_bits(e, e->buffered_byte + carry, 8);
  e->buffered_byte = lead & 0xff;
  for (; e->num_buffered > 1; e->num_buffered--) sink_bits(e, (0xff + carry) & 0xff, 8);
}

/* CABAC_ArithmeticEncoder.cpp:113-178 */
void orc_enc_bin(orc_encoder* e, unsigned bin, uint8_t* ctx) {
  e->bins_coded++;
  uint32_t lps = k_lps[*ctx >> 1][(e->range >> 6) & 3];
  e->range -= lps;
  if (bin != (*ctx & 1u)) {
    unsigned n = orc_renorm_shift(lps);
    e->low = (e->low + e->range) << n;
    e->range = lps << n;
    *ctx = orc_ctx_next_lps(*ctx);
    e->bits_left -= (int)n;
  } else {
    *ctx = orc_ctx_next_mps(*ctx);
    if (e->range >= 256) return; /* no renorm, no write-out test (:145-159) */
    e->low <<= 1; e->range <<= 1; e->bits_left--;
  }
  enc_flush_if_needed(e);
}

/* CABAC_ArithmeticEncoder.cpp:250-270 */
void orc_enc_ep(orc_encoder* e, unsigned bin) {
  e->bins_coded++;
  e->low <<= 1;
  if (bin) e->low += e->range;
  e->bits_left--;
  enc_flush_if_needed(e);
}

/* CABAC_ArithmeticEncoder.cpp:278-319 */
void orc_enc_bins_ep(orc_encoder* e, uint32_t bins, int n) {
  e->bins_coded += (uint64_t)n;
  while (n > 8) {
    n -= 8;
    uint32_t pat = bins >> n;
    e->low = (e->low << 8) + e->range * pat;
    bins -= pat << n;
    e->bits_left -= 8;
    enc_flush_if_needed(e);
  }
  e->low = (e->low << n) + e->range * bins;
  e->bits_left -= n;
  enc_flush_if_needed(e);
}

/* CABAC_ArithmeticEncoder.cpp:326-367 */
void orc_enc_trm(orc_encoder* e, unsigned bin) {
  e->bins_coded++;
  e->range -= 2;
  if (bin) {
    e->low = (e->low + e->range) << 7;
    e->range = 2u << 7;
    e->bits_left -= 7;
  } else if (e->range >= 256) {
    return;
  } else {
    e->low <<= 1; e->range <<= 1; e->bits_left--;
  }
  enc_flush_if_needed(e);
}

/* CABAC_ArithmeticEncoder.cpp:70-105 */
void orc_enc_finish(orc_encoder* e) {
  orc_enc_trm(e, 1);
  if (e->low >> (32 - e->bits_left)) {
    sink_bits(e, e->buffered_byte + 1, 8);
    for (; e->num_buffered > 1; e->num_buffered--) sink_bits(e, 0x00, 8);
    e->low -= 1u << (32 - e->bits_left);
  } else {
    if (e->num_buffered > 0) sink_bits(e, e->buffered_byte, 8);
    for (; e->num_buffered > 1; e->num_buffered--) sink_bits(e, 0xff, 8);
  }
  sink_bits(e, e->low >> 8, (unsigned)(24 - e->bits_left));
  sink_bits(e, 1, 1);
  sink_align_zero(e);
}

/* ======================================================================== *
 *  Decoder                                                                  *
 * ======================================================================== */

/* CABAC_BitstreamFile.cpp:153-158: EOF reads as 0xFF */
static uint32_t dec_byte(orc_decoder* d) {
  uint32_t b = d->pos < d->len ? d->in[d->pos] : 0xffu;
  d->pos++;
  d->last_byte = b;
  return b;
}

/* CABAC_ArithmeticDecoder.cpp:54-60 */
void orc_dec_start(orc_decoder* d, const uint8_t* in, uint64_t len) {
  d->in = in; d->len = len; d->pos = 0; d->last_byte = 0;
  d->range = 510; d->bits_needed = -8;
  d->value = dec_byte(d) << 8;
  d->value |= dec_byte(d);
}

/* CABAC_ArithmeticDecoder.cpp:87-190 */
unsigned orc_dec_bin(orc_decoder* d, uint8_t* ctx) {
  unsigned bin;
  uint32_t lps = k_lps[*ctx >> 1][((d->range >> 6) - 4) & 3]; /* &3: stay in bounds where the reference is UB (range<256) */
  d->range -= lps;
  uint32_t scaled = d->range << 7;
  if (d->value < scaled) {
    bin = *ctx & 1u;
    *ctx = orc_ctx_next_mps(*ctx);
    if (scaled >= (256u << 7)) return bin;
    d->range = scaled >> 6;
    d->value <<= 1;
    if (++d->bits_needed == 0) { d->bits_needed = -8; d->value += dec_byte(d); }
  } else {
    unsigned n = orc_renorm_shift(lps);
    d->value = (d->value - scaled) << n;
    d->range = lps << n;
    bin = 1u - (*ctx & 1u);
    *ctx = orc_ctx_next_lps(*ctx);
    d->bits_needed += (int)n;
    if (d->bits_needed >= 0) { d->value += dec_byte(d) << d->bits_needed; d->bits_needed -= 8; }
  }
  return bin;
}

/* CABAC_ArithmeticDecoder.cpp:288-331 */
unsigned orc_dec_ep(orc_decoder* d) {
  d->value <<= 1;
  if (++d->bits_needed >= 0) { d->bits_needed = -8; d->value += dec_byte(d); }
  uint32_t scaled = d->range << 7;
  if (d->value >= scaled) { d->value -= scaled; return 1; }
  return 0;
}

/* CABAC_ArithmeticDecoder.cpp:333-421 */
uint32_t orc_dec_bins_ep(orc_decoder* d, int n) {
  uint32_t bins = 0;
  while (n > 8) {
    d->value = (d->value << 8) + (dec_byte(d) << (8 + d->bits_needed));
    uint32_t scaled = d->range << 15;
    for (int i = 0; i < 8; ++i) {
      bins <<= 1; scaled >>= 1;
      if (d->value >= scaled) { bins |= 1; d->value -= scaled; }
    }
    n -= 8;
  }
  d->bits_needed += n;
  d->value <<= n;
  if (d->bits_needed >= 0) { d->value += dec_byte(d) << d->bits_needed; d->bits_needed -= 8; }
  uint32_t scaled = d->range << (n + 7);
  for (int i = 0; i < n; ++i) {
    bins <<= 1; scaled >>= 1;
    if (d->value >= scaled) { bins |= 1; d->value -= scaled; }
  }
  return bins;
}

/* CABAC_ArithmeticDecoder.cpp:423-472 */
unsigned orc_dec_trm(orc_decoder* d) {
  d->range -= 2;
  uint32_t scaled = d->range << 7;
  if (d->value >= scaled) return 1; /* no renorm, no read (:427-436) */
  if (scaled < (256u << 7)) {
    d->range = scaled >> 6;
    d->value <<= 1;
    if (++d->bits_needed == 0) { d->bits_needed = -8; d->value += dec_byte(d); }
  }
  return 0;
}

/* CABAC_ArithmeticDecoder.cpp:73-85 with the asserts turned into a result */
int orc_dec_finish(orc_decoder* d) {
  unsigned t = orc_dec_trm(d);
  int stop = ((d->last_byte << (8 + d->bits_needed)) & 0xffu) == 0x80u;
  return t == 1 && stop;
}

/* ======================================================================== *
 *  Context initialisation                                                   *
 * ======================================================================== */

/* CABAC_ContextModelsInit.cpp:124-148 */
void orc_map_prob_to_state(double p0, int* mps, int* state) {
  double plps;
  if (p0 >= 0.5) { plps = 1.0 - p0; *mps = 0; } else { plps = p0; *mps = 1; }
  if (plps < 0.01875) plps = 0.01875;
  int s = (int)round(62 * log10(2.0 * plps) / log10(2.0 * 0.01875));
  *state = s < 0 ? 0 : (s > 62 ? 62 : s);
}

uint8_t orc_ctx_from_p0(double p0) {
  int mps, st;
  orc_map_prob_to_state(p0, &mps, &st);
  return orc_ctx_make((unsigned)mps, (unsigned)st);
}

/* CABAC_ContextModelsInit.cpp:82-112 */
void orc_init_by_prob(const double* p0, int n, uint8_t* ctx) {
  for (int i = 0; i < n; ++i) ctx[i] = orc_ctx_from_p0(p0[i]);
}

/* CABAC_ContextModelsInit.cpp:51-80: triples [ctxIdx mps state], ctxIdx ignored */
void orc_init_by_state(const double* t, int n, uint8_t* ctx) {
  for (int i = 0; i < n; ++i)
    ctx[i] = orc_ctx_make((unsigned)t[3 * i + 1], (unsigned)t[3 * i + 2]);
}

/* MATLAB uint8(x): round half away from zero, saturate (cabacEncode.m:30, cabacDemo.m:77) */
uint8_t orc_matlab_uint8(double x) {
  if (!(x == x)) return 0;
  double r = floor(fabs(x) + 0.5);
  if (x < 0) r = -r;
  if (r <= 0) return 0;
  if (r >= 255) return 255;
  return (uint8_t)r;
}

/* ======================================================================== *
 *  Binarizer / debinarizer / finish detector                                *
 * ======================================================================== */

static int bit_length64(uint64_t x) { int n = 0; while (x) { ++n; x >>= 1; } return n; }

/* FLCode, cabacBinarizer.m:71-75 */
static int put_fixed(uint8_t* b, uint64_t v, int nbits) {
  for (int i = 0; i < nbits; ++i) b[i] = (uint8_t)((v >> (nbits - 1 - i)) & 1u);
  return nbits;
}

static int eg_order(int method) { return method - ORC_BIN_EG0; }
static int tr_order(int method) { return method - ORC_BIN_TR0; }

int orc_binarize(uint32_t v, uint32_t Nq, int method, uint8_t* b) {
  uint32_t maxv = Nq - 1;
  int n = 0;
  switch (method) {
    case ORC_BIN_TU: /* cabacBinarizer.m:30-37 */
      if (v == maxv) { for (uint32_t i = 0; i < maxv; ++i) b[n++] = 1; }
      else { for (uint32_t i = 0; i < v; ++i) b[n++] = 1; b[n++] = 0; }
      return n;
    case ORC_BIN_EG0: case ORC_BIN_EG1: case ORC_BIN_EG2: { /* cabacBinarizer.m:56-69 */
      int k = eg_order(method);
      /* n_p = floor(log2(v/2^k+1))+1 == bit length of (v>>k)+1 */
      int np = bit_length64(((uint64_t)v >> k) + 1);
      int ns = k + np - 1;
      uint64_t vs = (uint64_t)v - ((uint64_t)1 << k) * (((uint64_t)1 << (np - 1)) - 1);
      for (int i = 0; i < np - 1; ++i) b[n++] = 1;
      b[n++] = 0;
      n += put_fixed(b + n, vs, ns);
      return n;
    }
    case ORC_BIN_FL32: /* cabacBinarizer.m:26-27 */
      return put_fixed(b, v, 32);
    case ORC_BIN_TR0: case ORC_BIN_TR1: case ORC_BIN_TR2: { /* cabacBinarizer.m:39-54 */
      int k = tr_order(method);
      uint32_t np = (v >> k) + 1;
      uint32_t vs = v - ((np - 1) << k);
      for (uint32_t i = 0; i + 1 < np; ++i) b[n++] = 1;
      b[n++] = 0;
      if (v >= maxv) { for (int i = 0; i < k; ++i) b[n++] = 1; } /* escape is a TODO upstream */
      else n += put_fixed(b + n, vs, k);
      return n;
    }
  }
  return -1;
}

static int first_zero(const uint8_t* b, int n) { /* 1-based, 0 = none */
  for (int i = 0; i < n; ++i) if (!b[i]) return i + 1;
  return 0;
}
static uint64_t get_fixed(const uint8_t* b, int nbits) { /* rFLCode, cabacDebinarizer.m:55-57 */
  uint64_t v = 0;
  for (int i = 0; i < nbits; ++i) v = (v << 1) | b[i];
  return v;
}

uint32_t orc_debinarize(const uint8_t* b, int n, uint32_t Nq, int method) {
  uint32_t maxv = Nq - 1;
  switch (method) {
    case ORC_BIN_TU: { /* cabacDebinarizer.m:28-31 */
      int z = first_zero(b, n);
      return z ? (uint32_t)(z - 1) : maxv;
    }
    case ORC_BIN_EG0: case ORC_BIN_EG1: case ORC_BIN_EG2: { /* cabacDebinarizer.m:46-53 */
      int k = eg_order(method), np = first_zero(b, n);
      int ns = n - np;
      return (uint32_t)((((uint64_t)1 << k) * (((uint64_t)1 << (np - 1)) - 1)) + get_fixed(b + np, ns));
    }
    case ORC_BIN_FL32: return (uint32_t)get_fixed(b, 32);
    case ORC_BIN_TR0: case ORC_BIN_TR1: case ORC_BIN_TR2: { /* cabacDebinarizer.m:33-44 */
      int k = tr_order(method), np = first_zero(b, n);
      if (!np) return maxv;
      return (uint32_t)((((uint64_t)(np - 1)) << k) + get_fixed(b + np, k));
    }
  }
  return 0;
}

/* cabacDecodeSymbolFinished.m:10-32.  The reference has cases for TU and EG-k
 * only; FL32 (finished at bin 32) is this oracle's extension, TR returns -1. */
int orc_symbol_finished(const uint8_t* g, int n, uint32_t Nq, int method, int* n_p, int* n_s) {
  switch (method) {
    case ORC_BIN_TU:
      return g[n - 1] == 0 || (uint32_t)n == Nq - 1;
    case ORC_BIN_EG0: case ORC_BIN_EG1: case ORC_BIN_EG2: {
      int k = eg_order(method);
      if (*n_s == -1) {
        if (g[n - 1] == 0) {
          *n_p = *n_p + n;
          *n_s = k + *n_p - 1;
          if (*n_s == 0) return 1;
        }
        return 0;
      }
      if (*n_s == 1) return 1;
      *n_s = *n_s - 1;
      return 0;
    }
    case ORC_BIN_FL32:
      return n == 32;
  }
  return -1;
}

/* ======================================================================== *
 *  Context selection                                                        *
 * ======================================================================== */

int orc_profile_num_ctx(int profile, int Nlbp) {
  switch (profile) {
    case ORC_PROFILE_DEMO: return 3;
    case ORC_PROFILE_ISS: return 7 * Nlbp + 2;
    case ORC_PROFILE_FLAT: return 2 * Nlbp + 2;
    case ORC_PROFILE_FLAT_EPSUF: return Nlbp + 1;
  }
  return 0;
}

int orc_select_ctx(int profile, int n, const uint8_t* g, const uint8_t* up, int up_len,
                   int Nlbp, unsigned types) {
  if (profile == ORC_PROFILE_DEMO) { /* cabacDemo.m:113-121 */
    if (n == 1 && up_len > 0) return up[0] == 1 ? 1 : 2;
    return 0;
  }
  /* own prefix ends at the first 0 among the bins already coded */
  int np_prev = first_zero(g, n - 1);
  int in_prefix = !(np_prev && n > np_prev);
  if (profile == ORC_PROFILE_FLAT || profile == ORC_PROFILE_FLAT_EPSUF) {
    if (in_prefix) return n <= Nlbp ? n - 1 : Nlbp;
    if (profile == ORC_PROFILE_FLAT_EPSUF) return -1;
    int m = n - np_prev;
    return m <= Nlbp ? Nlbp + m : 2 * Nlbp + 1;
  }
  /* ISS: cabacContextSelection.m:24-67; ids below are 1-based like the MATLAB */
  int np_up = first_zero(up, up_len);
  int up_in_prefix = !(np_up && n > np_up);
  int id;
  if (in_prefix) {
    if (n <= Nlbp) {
      id = n;
      if (up_len >= n && up_in_prefix) {
        if (up[n - 1] == 0 && (types & ORC_CM_COND0)) id = Nlbp + n;
        else if (up[n - 1] == 1 && (types & ORC_CM_COND1)) id = 2 * Nlbp + n;
      } else if (n > 1 && g[n - 2] == 1 && (types & ORC_CM_CONDBINLFT)) {
        id = 3 * Nlbp + n - 1;
      }
    } else {
      id = 7 * Nlbp + 1;
    }
  } else {
    int m = n - np_prev;
    if (m <= Nlbp) {
      id = 4 * Nlbp + m;
      if (up_len >= n && !up_in_prefix) { /* absolute position n in the neighbour (:56-60) */
        if (up[n - 1] == 0 && (types & ORC_CM_CONDS0)) id = 5 * Nlbp + m;
        else if (up[n - 1] == 1 && (types & ORC_CM_CONDS1)) id = 6 * Nlbp + m;
      }
    } else {
      id = 7 * Nlbp + 2;
    }
  }
  return id - 1; /* cabacEncode.m:64, cabacDecode.m:42 */
}

/* ======================================================================== *
 *  Thread helper                                                            *
 * ======================================================================== */

typedef void (*stream_fn)(void* arg, uint32_t s);
typedef struct { stream_fn fn; void* arg; uint32_t n; uint32_t* next; pthread_mutex_t* mu; } pool_arg;

static void* pool_main(void* p) {
  pool_arg* a = (pool_arg*)p;
  for (;;) {
    pthread_mutex_lock(a->mu);
    uint32_t s0 = *a->next; *a->next += 16;
    pthread_mutex_unlock(a->mu);
    if (s0 >= a->n) break;
    uint32_t s1 = s0 + 16 < a->n ? s0 + 16 : a->n;
    for (uint32_t s = s0; s < s1; ++s) a->fn(a->arg, s);
  }
  return NULL;
}

static void run_streams(stream_fn fn, void* arg, uint32_t n, int n_threads) {
  if (n_threads <= 1 || n < 2) { for (uint32_t s = 0; s < n; ++s) fn(arg, s); return; }
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  uint32_t next = 0;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pool_arg a = {fn, arg, n, &next, &mu};
  for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], NULL, pool_main, &a);
  pool_main(&a);
  for (int t = 1; t < n_threads; ++t) pthread_join(th[t], NULL);
}

/* ======================================================================== *
 *  Batch drivers over op arrays                                             *
 * ======================================================================== */

typedef struct {
  const uint64_t *op_off, *byte_off;
  const uint8_t *ops8, *bytes, *ctx_init;
  const uint16_t* ops16;
  uint32_t n_ctx, ep, trm;
  int per_stream;
  uint8_t *out, *out_bins, *finish_ok;
  uint64_t out_stride;
  uint32_t* out_len;
} ops_job;

static void ops_job_init(ops_job* j, const void* ops, int width) {
  memset(j, 0, sizeof *j);
  if (width == 2) { j->ops16 = (const uint16_t*)ops; j->ep = ORC_OP16_EP; j->trm = ORC_OP16_TRM; }
  else { j->ops8 = (const uint8_t*)ops; j->ep = ORC_OP8_EP; j->trm = ORC_OP8_TRM; }
}
static inline uint32_t job_op(const ops_job* j, uint64_t i) { return j->ops8 ? j->ops8[i] : j->ops16[i]; }

static uint8_t* ctx_alloc(const ops_job* j, uint32_t s) {
  uint32_t n = j->n_ctx ? j->n_ctx : 1;
  uint8_t* c = (uint8_t*)malloc(n);
  if (j->n_ctx) memcpy(c, j->ctx_init + (j->per_stream ? (uint64_t)s * j->n_ctx : 0), j->n_ctx);
  return c;
}

static void encode_ops_stream(void* arg, uint32_t s) {
  ops_job* j = (ops_job*)arg;
  uint8_t* ctx = ctx_alloc(j, s);
  orc_encoder e;
  orc_enc_attach(&e, j->out + (uint64_t)s * j->out_stride, j->out_stride);
  orc_enc_start(&e);
  for (uint64_t i = j->op_off[s]; i < j->op_off[s + 1]; ++i) {
    uint32_t o = job_op(j, i), code = o >> 1, bin = o & 1u;
    /* op format (include/isscabac.h): a code >= n_ctx that is not the terminate code is a bypass bin */
    if (code == j->ep || (code != j->trm && code >= j->n_ctx)) orc_enc_ep(&e, bin);
    else if (code == j->trm) orc_enc_trm(&e, bin);
    else orc_enc_bin(&e, bin, &ctx[code]);
  }
  orc_enc_finish(&e);
  j->out_len[s] = (uint32_t)e.n_out;
  free(ctx);
}

int orc_encode_ops(uint32_t n_streams, const uint64_t* op_off, const void* ops, int op_width,
                   const uint8_t* ctx_init, uint32_t n_ctx, int per_stream_init,
                   uint8_t* out, uint64_t out_stride, uint32_t* out_len, int n_threads) {
  ops_job j;
  ops_job_init(&j, ops, op_width);
  j.op_off = op_off; j.ctx_init = ctx_init; j.n_ctx = n_ctx; j.per_stream = per_stream_init;
  j.out = out; j.out_stride = out_stride; j.out_len = out_len;
  run_streams(encode_ops_stream, &j, n_streams, n_threads);
  return 0;
}

static void decode_ops_stream(void* arg, uint32_t s) {
  ops_job* j = (ops_job*)arg;
  uint8_t* ctx = ctx_alloc(j, s);
  orc_decoder d;
  orc_dec_start(&d, j->bytes + j->byte_off[s], j->byte_off[s + 1] - j->byte_off[s]);
  for (uint64_t i = j->op_off[s]; i < j->op_off[s + 1]; ++i) {
    uint32_t code = job_op(j, i) >> 1;
    unsigned bin;
    if (code == j->ep || (code != j->trm && code >= j->n_ctx)) bin = orc_dec_ep(&d);
    else if (code == j->trm) bin = orc_dec_trm(&d);
    else bin = orc_dec_bin(&d, &ctx[code]);
    j->out_bins[i] = (uint8_t)bin;
  }
  if (j->finish_ok) j->finish_ok[s] = (uint8_t)orc_dec_finish(&d);
  free(ctx);
}

int orc_decode_ops(uint32_t n_streams, const uint64_t* byte_off, const uint8_t* bytes,
                   const uint64_t* op_off, const void* ops, int op_width,
                   const uint8_t* ctx_init, uint32_t n_ctx, int per_stream_init,
                   uint8_t* out_bins, uint8_t* finish_ok, int n_threads) {
  ops_job j;
  ops_job_init(&j, ops, op_width);
  j.op_off = op_off; j.byte_off = byte_off; j.bytes = bytes;
  j.ctx_init = ctx_init; j.n_ctx = n_ctx; j.per_stream = per_stream_init;
  j.out_bins = out_bins; j.finish_ok = finish_ok;
  run_streams(decode_ops_stream, &j, n_streams, n_threads);
  return 0;
}

/* ======================================================================== *
 *  Symbol-level drivers                                                     *
 * ======================================================================== */

#define ORC_MAX_SYM_BINS 4200

static int has_up(const orc_symcfg* c, uint64_t i) {
  if (c->profile == ORC_PROFILE_ISS) return c->rows ? (i % c->rows) != 0 : i > 0; /* cabacEncode.m:52 */
  if (c->profile == ORC_PROFILE_DEMO) return i > 0;                              /* cabacDemo.m:105 */
  return 0;
}

uint64_t orc_symbols_to_ops(const orc_symcfg* c, const uint32_t* sym, uint64_t n_sym,
                            uint8_t* ops, uint64_t cap) {
  uint8_t cur[ORC_MAX_SYM_BINS], prev[ORC_MAX_SYM_BINS];
  int prev_len = 0;
  uint64_t k = 0;
  for (uint64_t i = 0; i < n_sym; ++i) {
    int len = orc_binarize(sym[i], c->Nq, c->method, cur);
    int ul = has_up(c, i) ? prev_len : 0;
    for (int n = 1; n <= len; ++n) {
      int cx = orc_select_ctx(c->profile, n, cur, prev, ul, c->Nlbp, c->types);
      uint32_t code = cx < 0 ? ORC_OP8_EP : (uint32_t)cx;
      if (k < cap) ops[k] = (uint8_t)((code << 1) | cur[n - 1]);
      ++k;
    }
    memcpy(prev, cur, (size_t)len);
    prev_len = len;
  }
  return k;
}

typedef struct {
  const orc_symcfg* cfg;
  const uint64_t *sym_off, *byte_off;
  const uint32_t* symbols;
  const uint8_t *ctx_init, *bytes;
  uint32_t n_ctx;
  int per_stream;
  uint8_t *out, *finish_ok;
  uint64_t out_stride;
  uint32_t *out_len, *bits_after, *out_symbols;
} sym_job;

static void encode_sym_stream(void* arg, uint32_t s) {
  sym_job* j = (sym_job*)arg;
  const orc_symcfg* c = j->cfg;
  uint32_t nc = j->n_ctx ? j->n_ctx : 1;
  uint8_t* ctx = (uint8_t*)malloc(nc);
  if (j->n_ctx) memcpy(ctx, j->ctx_init + (j->per_stream ? (uint64_t)s * j->n_ctx : 0), j->n_ctx);
  uint8_t cur[ORC_MAX_SYM_BINS], prev[ORC_MAX_SYM_BINS];
  int prev_len = 0;
  orc_encoder e;
  orc_enc_attach(&e, j->out + (uint64_t)s * j->out_stride, j->out_stride);
  orc_enc_start(&e);
  uint64_t base = j->sym_off[s], n_sym = j->sym_off[s + 1] - base;
  for (uint64_t i = 0; i < n_sym; ++i) {
    int len = orc_binarize(j->symbols[base + i], c->Nq, c->method, cur);
    int ul = has_up(c, i) ? prev_len : 0;
    for (int n = 1; n <= len; ++n) {
      int cx = orc_select_ctx(c->profile, n, cur, prev, ul, c->Nlbp, c->types);
      if (cx < 0) orc_enc_ep(&e, cur[n - 1]);
      else orc_enc_bin(&e, cur[n - 1], &ctx[cx]);
    }
    if (j->bits_after) j->bits_after[base + i] = (uint32_t)orc_enc_num_bits(&e);
    memcpy(prev, cur, (size_t)len);
    prev_len = len;
  }
  orc_enc_finish(&e);
  j->out_len[s] = (uint32_t)e.n_out;
  free(ctx);
}

int orc_encode_symbols(const orc_symcfg* cfg, uint32_t n_streams, const uint64_t* sym_off,
                       const uint32_t* symbols, const uint8_t* ctx_init, uint32_t n_ctx,
                       int per_stream_init, uint8_t* out, uint64_t out_stride, uint32_t* out_len,
                       uint32_t* bits_after_symbol, int n_threads) {
  sym_job j;
  memset(&j, 0, sizeof j);
  j.cfg = cfg; j.sym_off = sym_off; j.symbols = symbols; j.ctx_init = ctx_init; j.n_ctx = n_ctx;
  j.per_stream = per_stream_init; j.out = out; j.out_stride = out_stride; j.out_len = out_len;
  j.bits_after = bits_after_symbol;
  run_streams(encode_sym_stream, &j, n_streams, n_threads);
  return 0;
}

static void decode_sym_stream(void* arg, uint32_t s) {
  sym_job* j = (sym_job*)arg;
  const orc_symcfg* c = j->cfg;
  uint32_t nc = j->n_ctx ? j->n_ctx : 1;
  uint8_t* ctx = (uint8_t*)malloc(nc);
  if (j->n_ctx) memcpy(ctx, j->ctx_init + (j->per_stream ? (uint64_t)s * j->n_ctx : 0), j->n_ctx);
  uint8_t cur[ORC_MAX_SYM_BINS], prev[ORC_MAX_SYM_BINS];
  int prev_len = 0;
  orc_decoder d;
  orc_dec_start(&d, j->bytes + j->byte_off[s], j->byte_off[s + 1] - j->byte_off[s]);
  uint64_t base = j->sym_off[s], n_sym = j->sym_off[s + 1] - base;
  for (uint64_t i = 0; i < n_sym; ++i) {
    int ul = has_up(c, i) ? prev_len : 0;
    int n = 1, n_p = 0, n_s = -1, fin = 0; /* cabacDecode.m:35 */
    while (!fin && n < ORC_MAX_SYM_BINS) {
      int cx = orc_select_ctx(c->profile, n, cur, prev, ul, c->Nlbp, c->types);
      cur[n - 1] = (uint8_t)(cx < 0 ? orc_dec_ep(&d) : orc_dec_bin(&d, &ctx[cx]));
      fin = orc_symbol_finished(cur, n, c->Nq, c->method, &n_p, &n_s);
      if (fin < 0) fin = 1;
      ++n;
    }
    int len = n - 1;
    j->out_symbols[base + i] = orc_debinarize(cur, len, c->Nq, c->method);
    memcpy(prev, cur, (size_t)len);
    prev_len = len;
  }
  if (j->finish_ok) j->finish_ok[s] = (uint8_t)orc_dec_finish(&d);
  free(ctx);
}

int orc_decode_symbols(const orc_symcfg* cfg, uint32_t n_streams, const uint64_t* byte_off,
                       const uint8_t* bytes, const uint64_t* sym_off, const uint8_t* ctx_init,
                       uint32_t n_ctx, int per_stream_init, uint32_t* out_symbols,
                       uint8_t* finish_ok, int n_threads) {
  sym_job j;
  memset(&j, 0, sizeof j);
  j.cfg = cfg; j.sym_off = sym_off; j.byte_off = byte_off; j.bytes = bytes;
  j.ctx_init = ctx_init; j.n_ctx = n_ctx; j.per_stream = per_stream_init;
  j.out_symbols = out_symbols; j.finish_ok = finish_ok;
  run_streams(decode_sym_stream, &j, n_streams, n_threads);
  return 0;
}

/* ======================================================================== *
 *  ISS context-init statistics  (cabacInitContextModel.m:15-129)            *
 * ======================================================================== */

typedef struct { double hit, tot; } frac;
static double frac_val(frac f) { return f.tot > 0 ? f.hit / f.tot : 0.0; }

void orc_iss_ctx_init(const uint32_t* G, uint32_t rows, uint32_t cols, uint32_t Nq, int method,
                      int Nlbp, unsigned types, double* p0) {
  uint64_t N = (uint64_t)rows * cols;
  /* binarise everything once */
  uint8_t* bins = (uint8_t*)malloc(N * 80);
  int* L = (int*)malloc(N * sizeof(int));
  int* np = (int*)malloc(N * sizeof(int));
  for (uint64_t i = 0; i < N; ++i) {
    uint8_t tmp[ORC_MAX_SYM_BINS];
    int len = orc_binarize(G[i], Nq, method, tmp);
    if (len > 80) len = 80;
    memcpy(bins + i * 80, tmp, (size_t)len);
    L[i] = len;
    int z = first_zero(tmp, len);
    np[i] = z ? z : len; /* :18-20 */
  }
#define B(i, n) bins[(i) * 80 + ((n) - 1)] /* 1-based bin n of symbol i */
  int nctx = 7 * Nlbp + 2;
  for (int i = 0; i < nctx; ++i) p0[i] = 0.0;
  for (int n = 1; n <= Nlbp; ++n) {
    frac pre = {0, 0}, c0 = {0, 0}, c0n = {0, 0}, c1 = {0, 0}, c1n = {0, 0};
    frac bl = {0, 0}, bln = {0, 0}, suf = {0, 0}, s0 = {0, 0}, s0n = {0, 0}, s1 = {0, 0}, s1n = {0, 0};
    for (uint64_t i = 0; i < N; ++i) {
      int hasup = (i % rows) != 0;
      uint64_t u = i - 1;
      int np_up = hasup ? np[u] : 0; /* :21 zeros in the first row */
      int L_up = hasup ? L[u] : 1;   /* :23 length of the NaN placeholder cell */
      /* prefix :31-33 */
      if (n <= np[i]) { pre.tot++; pre.hit += B(i, n) == 0; }
      /* cond0 / cond1 :37-58 (n <= np_up implies a real neighbour) */
      if (n <= np[i] && n <= np_up) {
        c0.tot++; c0.hit += (B(i, n) == 0 && B(u, n) == 0); c0n.tot++; c0n.hit += B(u, n) == 0;
        c1.tot++; c1.hit += (B(i, n) == 0 && B(u, n) == 1); c1n.tot++; c1n.hit += B(u, n) == 1;
      }
      /* condbinlft :62-72 */
      if (n + 1 <= np[i] && np_up < n + 1) {
        bl.tot++; bl.hit += (B(i, n + 1) == 0 && B(i, n) == 1);
        bln.tot++; bln.hit += B(i, n) == 1;
      }
      /* suffix :77-80 */
      if (n + np[i] <= L[i]) { suf.tot++; suf.hit += B(i, n + np[i]) == 0; }
      /* conds0 / conds1 :84-106; note the normaliser looks at y(n), not y(n+np_up) (:90,:102).
       * In the first row Gbin_up1 is a NaN cell of length 1 with np_up1=0, so
       * n+np_up1<=L_up1 only for n==1, where NaN==0 and NaN==1 are both false. */
      if (n + np[i] <= L[i] && n + np_up <= L_up) {
        int y_suf0 = hasup ? B(u, n + np_up) == 0 : 0, y_suf1 = hasup ? B(u, n + np_up) == 1 : 0;
        int y_n0 = hasup ? B(u, n) == 0 : 0, y_n1 = hasup ? B(u, n) == 1 : 0;
        s0.tot++; s0.hit += (B(i, n + np[i]) == 0 && y_suf0); s0n.tot++; s0n.hit += y_n0;
        s1.tot++; s1.hit += (B(i, n + np[i]) == 0 && y_suf1); s1n.tot++; s1n.hit += y_n1;
      }
    }
    p0[n - 1] = frac_val(pre);
    if (types & ORC_CM_COND0) { double nr = c0n.hit > 0 ? frac_val(c0n) : 1.0; p0[Nlbp + n - 1] = frac_val(c0) / nr; }
    if (types & ORC_CM_COND1) { double nr = c1n.hit > 0 ? frac_val(c1n) : 1.0; p0[2 * Nlbp + n - 1] = frac_val(c1) / nr; }
    if (types & ORC_CM_CONDBINLFT) { double nr = bln.hit > 0 ? frac_val(bln) : 1.0; p0[3 * Nlbp + n - 1] = frac_val(bl) / nr; }
    p0[4 * Nlbp + n - 1] = frac_val(suf);
    if (types & ORC_CM_CONDS0) { double nr = s0n.hit > 0 ? frac_val(s0n) : 1.0; p0[5 * Nlbp + n - 1] = frac_val(s0) / nr; }
    if (types & ORC_CM_CONDS1) { double nr = s1n.hit > 0 ? frac_val(s1n) : 1.0; p0[6 * Nlbp + n - 1] = frac_val(s1) / nr; }
  }
  /* rest contexts :111-126 */
  frac rp = {0, 0}, rs = {0, 0};
  int n0 = Nlbp + 1;
  for (uint64_t i = 0; i < N; ++i) {
    if (n0 <= np[i]) for (int n = n0; n <= np[i]; ++n) { rp.tot++; rp.hit += B(i, n) == 0; }
    if (n0 > np[i] && n0 <= L[i]) for (int n = n0; n <= L[i]; ++n) { rs.tot++; rs.hit += B(i, n) == 0; }
  }
  p0[7 * Nlbp] = frac_val(rp);
  p0[7 * Nlbp + 1] = frac_val(rs);
#undef B
  free(bins); free(L); free(np);
}
