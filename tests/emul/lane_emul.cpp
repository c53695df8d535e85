// TEST INFRASTRUCTURE: compiles the product's per-lane coder logic
// (isscabac_b200/csrc/cabac_lane.cuh, the exact code the CUDA kernels run) with g++
// and drives it one stream at a time, so the algorithmic restructuring (eager byte
// output + walk-back carries, fused MPS/LPS step, closed-form binarizer) can be
// checked against the oracle on a machine without a GPU.  Never part of the product.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../isscabac_b200/csrc/cabac_lane.cuh"
#include "../../isscabac_b200/csrc/bin_emit.cuh"

using namespace cabac;

extern "C" {

int emul_encode_ops(uint32_t n_streams, const uint64_t* op_off, const void* ops, int width,
                    const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                    uint8_t* slab, uint64_t stride, uint32_t* lens, uint32_t* bits_trace) {
  const uint32_t ep = width == 1 ? 126u : 0x7FFEu, trm = width == 1 ? 125u : 0x7FFDu;
  std::vector<uint32_t> ctx(n_ctx ? n_ctx : 1);
  int ovf = 0;
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) ctx[c] = ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c];
    EncLane L;
    enc_start(L, slab + s * stride, (uint32_t)stride);
    for (uint64_t i = op_off[s]; i < op_off[s + 1]; ++i) {
      uint32_t o = width == 1 ? ((const uint8_t*)ops)[i] : ((const uint16_t*)ops)[i];
      uint32_t code = o >> 1, bin = o & 1u;
      if (code == ep || (code != trm && code >= n_ctx)) enc_bin_ep<true>(L, bin);   // op format: non-context codes are bypass bins
      else if (code == trm) enc_bin_trm<true>(L, bin);
      else enc_bin_ctx<true>(L, bin, ctx[code], fused_row(ctx[code]));
      if (bits_trace) bits_trace[i] = enc_bits_written(L);
    }
    enc_finish<true>(L);
    enc_flush_pending(L);
    lens[s] = L.nbytes;
    ovf |= L.overflow;
  }
  return ovf;
}

int emul_decode_ops(uint32_t n_streams, const uint64_t* byte_off, const uint8_t* bytes,
                    const uint64_t* op_off, const void* ops, int width,
                    const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                    uint8_t* bins, uint8_t* ok) {
  const uint32_t ep = width == 1 ? 126u : 0x7FFEu, trm = width == 1 ? 125u : 0x7FFDu;
  std::vector<uint32_t> ctx(n_ctx ? n_ctx : 1);
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) ctx[c] = ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c];
    DecLane D;
    dec_start(D, bytes + byte_off[s], (uint32_t)(byte_off[s + 1] - byte_off[s]));
    for (uint64_t i = op_off[s]; i < op_off[s + 1]; ++i) {
      uint32_t o = width == 1 ? ((const uint8_t*)ops)[i] : ((const uint16_t*)ops)[i];
      uint32_t code = o >> 1;
      if (code == ep || (code != trm && code >= n_ctx)) bins[i] = (uint8_t)dec_bin_ep(D);
      else if (code == trm) bins[i] = (uint8_t)dec_bin_trm(D);
      else bins[i] = (uint8_t)dec_bin_ctx(D, ctx[code], fused_row(ctx[code]));
    }
    ok[s] = (uint8_t)dec_finish(D);
  }
  return 0;
}

// symbols -> ops through the closed-form binarizer + context selection
uint64_t emul_symbols_to_ops(const int32_t* cfgv, const uint32_t* sym, uint64_t n, uint8_t* ops) {
  SymCfg cfg{cfgv[0], cfgv[1], (uint32_t)cfgv[2], cfgv[3], (uint32_t)cfgv[4], (uint32_t)cfgv[5]};
  uint64_t k = 0;
  SymCode prev{0, 0, 0};
  for (uint64_t i = 0; i < n; ++i) {
    SymCode c = sym_code(sym[i], cfg.Nq, cfg.method);
    bool up = sym_has_up(cfg, i);
    for (uint32_t b = 1; b <= c.len; ++b) {
      int cx = select_ctx(cfg, b, c.np, prev, up);
      uint32_t code = cx < 0 ? 126u : (uint32_t)cx;
      ops[k++] = (uint8_t)((code << 1) | sym_bin(c, b));
    }
    prev = c;
  }
  return k;
}

// bytes -> symbols through the incremental symbol decoder
int emul_decode_symbols(const int32_t* cfgv, const uint8_t* bytes, uint32_t len, uint64_t n_sym,
                        const uint8_t* ctx_init, uint32_t n_ctx, uint32_t* out, uint8_t* ok) {
  SymCfg cfg{cfgv[0], cfgv[1], (uint32_t)cfgv[2], cfgv[3], (uint32_t)cfgv[4], (uint32_t)cfgv[5]};
  std::vector<uint32_t> ctx(n_ctx ? n_ctx : 1);
  for (uint32_t c = 0; c < n_ctx; ++c) ctx[c] = ctx_init[c];
  DecLane D;
  dec_start(D, bytes, len);
  SymCode prev{0, 0, 0};
  for (uint64_t i = 0; i < n_sym; ++i) {
    SymDec sd;
    symdec_reset(sd);
    bool up = sym_has_up(cfg, i);
    uint32_t v = 0;
    for (;;) {
      int cx = select_ctx(cfg, sd.n + 1, sd.np, prev, up);
      uint32_t bin = cx < 0 ? dec_bin_ep(D) : dec_bin_ctx(D, ctx[cx], fused_row(ctx[cx]));
      if (symdec_push(sd, bin, cfg, v)) break;
    }
    out[i] = v;
    prev = sym_code(v, cfg.Nq, cfg.method);
  }
  *ok = (uint8_t)dec_finish(D);
  return 0;
}

// The emit phase of the u8 binarizer (k_bin_emit8, symbols.cu) on the host: the same per-thread code (bin_emit.cuh) for
// every "thread" of every tile of ONE stream of u8 symbols -- table set-up, scan, word-wise append into a stage of
// `stage_bytes` (a small stage forces the rounds of an over-full tile), tail bytes after the "barrier", pieces out.
// The threads of a phase run in the order `order` says (0: ascending, 1: descending, 2: odd lanes first): what they
// store between two barriers must not depend on it.  `skew0` = the op array's byte alignment.
struct EmulStage {
  uint8_t* stage;
  uint32_t w0, size;
  void word(uint32_t a, uint32_t w) { if (a - w0 < size) memcpy(stage + (a - w0), &w, 4); }
  void byte(uint32_t a, uint32_t b) { if (a - w0 < size) stage[a - w0] = (uint8_t)b; }
  void words(BinAcc& A, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t nfb) { bin_words_plain(*this, A, x0, x1, x2, nfb); }
};
uint64_t emul_bin_emit8(const int32_t* cfgv, const uint8_t* sym, uint64_t n, uint8_t* ops, uint32_t skew0,
                        uint32_t stage_bytes, int order) {
  SymCfg cfg{cfgv[0], cfgv[1], (uint32_t)cfgv[2], cfgv[3], (uint32_t)cfgv[4], (uint32_t)cfgv[5]};
  const uint32_t T = 128, ITEMS = 16, TILE = T * ITEMS;      // the kernel's geometry
  const LutGeom geom = lut_geom(cfg.profile, cfg.method, cfg.Nq);
  std::vector<uint32_t> lut16(4 * (LUT_MAX + 1), 0u), lut8(2 * (LUT_MAX + 1), 0u);
  for (uint32_t e = 0; e < geom.entries; ++e) {
    lut_entry(cfg, e, &lut16[4 * e]);
    lut8_from16(lut16[4 * e], lut16[4 * e + 1], lut16[4 * e + 3], lut8[2 * e], lut8[2 * e + 1]);
  }
  uint32_t len_tab[256];
  for (uint32_t v = 0; v < 256; ++v) len_tab[v] = sym_code(v, cfg.Nq, cfg.method).len;
  std::vector<uint8_t> stage(stage_bytes + 16, 0xEE);
  std::vector<uint32_t> perm(T);
  for (uint32_t t = 0; t < T; ++t) perm[t] = order == 0 ? t : (order == 1 ? T - 1 - t : (t < T / 2 ? 2 * t + 1 : 2 * (t - T / 2)));
  uint64_t tile_base = 0;
  for (uint64_t t0 = 0; t0 < n; t0 += TILE) {
    uint32_t lo[128], tot[128], nvalid[128];
    uint32_t block_total = 0;
    for (uint32_t t = 0; t < T; ++t) {
      const uint64_t i0 = t0 + (uint64_t)t * ITEMS;
      nvalid[t] = i0 >= n ? 0u : (n - i0 < ITEMS ? (uint32_t)(n - i0) : ITEMS);
      tot[t] = 0;
      for (uint32_t k = 0; k < nvalid[t]; ++k) tot[t] += len_tab[sym[i0 + k]];
      lo[t] = block_total;
      block_total += tot[t];
    }
    const uint32_t skew = (uint32_t)((skew0 + tile_base) & 15u), span = block_total + skew;
    for (uint32_t w0 = 0; w0 < span; w0 += stage_bytes) {
      std::fill(stage.begin(), stage.end(), (uint8_t)0xEE);
      BinAcc acc[128];
      uint32_t wpf[128], fbf[128];
      EmulStage st{stage.data(), w0, stage_bytes};
      for (uint32_t tt = 0; tt < T; ++tt) {
        const uint32_t t = perm[tt];
        const uint64_t i0 = t0 + (uint64_t)t * ITEMS;
        const uint32_t pos = lo[t] + skew;
        BinAcc& A = acc[t];
        bin_acc_start(A, pos);
        wpf[t] = A.wp;
        fbf[t] = A.fb;
        if (pos < w0 + stage_bytes && pos + tot[t] + 4u > w0) {
          for (uint32_t k = 0; k < nvalid[t]; ++k) {
            const uint64_t i = i0 + k;
            const uint32_t v = sym[i], u = i ? sym[i - 1] : 0u;
            const bool up = sym_has_up(cfg, i);
            uint32_t idx = lut_index(cfg, geom.dom, v, u, up);
            if (idx == LUT_ESC) idx = LUT_MAX;
            const uint32_t ex = lut8[2 * idx], ey = lut8[2 * idx + 1];
            if (ey >> 24) bin_append(A, st, ex, ey & 0x00ffffffu, ey >> 24);
            else A = bin_append_long(A, st, cfg, lut16[4 * idx], lut16[4 * idx + 1], lut16[4 * idx + 2], lut16[4 * idx + 3], v, u, up);
          }
        } else {
          A.a0 = 0u;
          A.fb = 0u;
        }
      }
      for (uint32_t tt = 0; tt < T; ++tt) bin_tail(acc[perm[tt]], st, wpf[perm[tt]], fbf[perm[tt]]);
      const uint32_t wend = w0 + stage_bytes < span ? w0 + stage_bytes : span;
      for (uint32_t b = w0; b < wend; ++b)
        if (b >= skew) ops[tile_base + b - skew] = stage[b - w0];
    }
    tile_base += block_total;
  }
  return tile_base;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// wide-window formulation (cabac_wide.cuh), driven with the kernels' schedule:
// general path up to the first 16-byte boundary of the op array, 16-op blocks, general tail
// ---------------------------------------------------------------------------
#include "../../isscabac_b200/csrc/cabac_wide.cuh"

namespace {
struct HostCtx {
  uint32_t* p;
  uint32_t load(uint32_t c) const { return p[c]; }
  void store(uint32_t c, uint32_t v) const { p[c] = v; }
  void store_sel(uint32_t c, uint32_t sel, uint32_t a, uint32_t b) const { p[c] = sel ? a : b; }
};
struct HostTab {   // host tokens are the state bytes themselves
  uint32_t token(uint32_t st) const { return st; }
  WRow row(uint32_t tok) const { return wide_row(tok); }
};
inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
}  // namespace

extern "C" {

int emul_encode_ops_wide(uint32_t n_streams, const uint64_t* op_off, const uint8_t* ops,
                         const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                         uint8_t* slab, uint64_t stride, uint32_t* lens) {
  std::vector<uint32_t> cs(n_ctx + 1);
  int ovf = 0;
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) cs[c] = ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c] & 127u;
    cs[n_ctx] = kEpState;
    HostCtx ctx{cs.data()};
    HostTab tab;
    EncWide E;
    encw_start(E, slab + s * stride, (uint32_t)stride);
    const uint8_t* p = ops + op_off[s];
    uint64_t n = op_off[s + 1] - op_off[s], i = 0;
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > n) head = n;
    for (; i < head; ++i) encw_general(E, p[i], ctx, tab, n_ctx);
    for (; i + 16 <= n; i += 16) {
      const uint32_t w[4] = {ld32(p + i), ld32(p + i + 4), ld32(p + i + 8), ld32(p + i + 12)};
      const uint32_t cw[4] = {op_codes4(w[0]), op_codes4(w[1]), op_codes4(w[2]), op_codes4(w[3])};
      if (block_has_trm(cw)) {
        for (int k = 0; k < 16; ++k) encw_general(E, p[i + k], ctx, tab, n_ctx);
      } else {
        encw_block16<false>(E, w, cw, ctx, tab, n_ctx);
      }
    }
    for (; i < n; ++i) encw_general(E, p[i], ctx, tab, n_ctx);
    lens[s] = encw_finish(E);
    ovf |= lens[s] > stride;
  }
  return ovf;
}

int emul_decode_ops_wide(uint32_t n_streams, const uint64_t* byte_off, const uint8_t* bytes,
                         const uint64_t* op_off, const uint8_t* ops,
                         const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                         uint8_t* bins, uint8_t* ok) {
  std::vector<uint32_t> cs(n_ctx + 1);
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) cs[c] = ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c] & 127u;
    cs[n_ctx] = kEpState;
    HostCtx ctx{cs.data()};
    HostTab tab;
    DecWide D;
    decw_start(D, bytes + byte_off[s], (uint32_t)(byte_off[s + 1] - byte_off[s]));
    const uint8_t* p = ops + op_off[s];
    uint8_t* q = bins + op_off[s];
    uint64_t n = op_off[s + 1] - op_off[s], i = 0;
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > n) head = n;
    for (; i < head; ++i) q[i] = (uint8_t)decw_general(D, p[i], ctx, tab, n_ctx);
    for (; i + 16 <= n; i += 16) {
      const uint32_t cw[4] = {op_codes4(ld32(p + i)), op_codes4(ld32(p + i + 4)), op_codes4(ld32(p + i + 8)),
                              op_codes4(ld32(p + i + 12))};
      if (block_has_trm(cw)) {
        for (int k = 0; k < 16; ++k) q[i + k] = (uint8_t)decw_general(D, p[i + k], ctx, tab, n_ctx);
      } else {
        uint32_t r[4];
        decw_block16<false>(D, cw, r, ctx, tab, n_ctx);
        memcpy(q + i, r, 16);
      }
    }
    for (; i < n; ++i) q[i] = (uint8_t)decw_general(D, p[i], ctx, tab, n_ctx);
    ok[s] = (uint8_t)decw_finish(D);
  }
  return 0;
}

// the same op arrays with every run of bypass ops coded by the multi-bin steps (encw_ep_run /
// decw_ep_run, chunks of up to kEpRunMax bins, emission / refill before and after a run as in the
// fused symbol kernels); everything else goes down the general path
int emul_encode_ops_wide_runs(uint32_t n_streams, const uint64_t* op_off, const uint8_t* ops,
                              const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                              uint8_t* slab, uint64_t stride, uint32_t* lens, uint32_t max_run) {
  std::vector<uint32_t> cs(n_ctx + 1);
  int ovf = 0;
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) cs[c] = ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c] & 127u;
    cs[n_ctx] = kEpState;
    HostCtx ctx{cs.data()};
    HostTab tab;
    EncWide E;
    encw_start(E, slab + s * stride, (uint32_t)stride);
    const uint8_t* p = ops + op_off[s];
    const uint64_t n = op_off[s + 1] - op_off[s];
    for (uint64_t i = 0; i < n;) {
      if ((p[i] >> 1) > kOpTrmCode) {
        uint32_t c = 0, bits = 0;
        while (i < n && (p[i] >> 1) > kOpTrmCode && c < max_run) { bits = (bits << 1) | (p[i] & 1u); ++c; ++i; }
        encw_emit(E);
        encw_ep_run(E, bits, c);
        encw_emit(E);
      } else {
        encw_general(E, p[i++], ctx, tab, n_ctx);
      }
    }
    lens[s] = encw_finish(E);
    ovf |= lens[s] > stride;
  }
  return ovf;
}

int emul_decode_ops_wide_runs(uint32_t n_streams, const uint64_t* byte_off, const uint8_t* bytes,
                              const uint64_t* op_off, const uint8_t* ops,
                              const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                              uint8_t* bins, uint8_t* ok, uint32_t max_run) {
  std::vector<uint32_t> cs(n_ctx + 1);
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) cs[c] = ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c] & 127u;
    cs[n_ctx] = kEpState;
    HostCtx ctx{cs.data()};
    HostTab tab;
    DecWide D;
    decw_start(D, bytes + byte_off[s], (uint32_t)(byte_off[s + 1] - byte_off[s]));
    const uint8_t* p = ops + op_off[s];
    uint8_t* q = bins + op_off[s];
    const uint64_t n = op_off[s + 1] - op_off[s];
    for (uint64_t i = 0; i < n;) {
      if ((p[i] >> 1) > kOpTrmCode) {
        uint32_t c = 0;
        while (i + c < n && (p[i + c] >> 1) > kOpTrmCode && c < max_run) ++c;
        decw_refill(D);
        const uint32_t v = decw_ep_run(D, c);
        decw_refill(D);
        for (uint32_t k = 0; k < c; ++k) q[i + k] = (uint8_t)((v >> (c - 1 - k)) & 1u);
        i += c;
      } else {
        q[i] = (uint8_t)decw_general(D, p[i], ctx, tab, n_ctx);
        ++i;
      }
    }
    ok[s] = (uint8_t)decw_finish(D);
  }
  return 0;
}

// the bypass runs as the tree decoder codes them (k_decode_symbols_tree): restoring division on the upper word
// (decw_ep_bits), the window topped up before the run only when the run would read unfilled bits, and after it
int emul_decode_ops_wide_runs_loop(uint32_t n_streams, const uint64_t* byte_off, const uint8_t* bytes,
                                   const uint64_t* op_off, const uint8_t* ops,
                                   const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                                   uint8_t* bins, uint8_t* ok, uint32_t max_run) {
  std::vector<uint32_t> cs(n_ctx + 1);
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) cs[c] = ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c] & 127u;
    cs[n_ctx] = kEpState;
    HostCtx ctx{cs.data()};
    HostTab tab;
    DecWide D;
    decw_start(D, bytes + byte_off[s], (uint32_t)(byte_off[s + 1] - byte_off[s]));
    const uint8_t* p = ops + op_off[s];
    uint8_t* q = bins + op_off[s];
    const uint64_t n = op_off[s + 1] - op_off[s];
    uint32_t since = 0;     // context bins since the window was last known to hold f <= 35
    for (uint64_t i = 0; i < n;) {
      if ((p[i] >> 1) > kOpTrmCode) {
        uint32_t c = 0;
        while (i + c < n && (p[i + c] >> 1) > kOpTrmCode && c < max_run) ++c;
        if (D.f + (int32_t)c > 54) decw_refill(D);
        const uint32_t v = decw_ep_bits(D, c);
        decw_refill(D);
        since = 0;
        for (uint32_t k = 0; k < c; ++k) q[i + k] = (uint8_t)((v >> (c - 1 - k)) & 1u);
        i += c;
      } else if ((p[i] >> 1) == kOpTrmCode) {
        q[i] = (uint8_t)decw_general(D, p[i], ctx, tab, n_ctx);
        ++i;
        since = 0;
      } else {
        // the kernel's group schedule: up to four context bins between two (voted) refills at f >= kLazyDec
        q[i] = (uint8_t)decw_op<0>(D, p[i] >> 1, ctx, tab, n_ctx);
        ++i;
        if (++since == 4u) { if (D.f >= kLazyDec) decw_refill(D); since = 0; }
      }
    }
    ok[s] = (uint8_t)decw_finish(D);
  }
  return 0;
}

// direct hook for the (practically unreachable) carry walk past a 0xFFFFFFFF pending word
void emul_carry_walk(uint32_t* row, uint32_t wp, uint32_t cap_words) { encw_carry_walk(wp, cap_words, row + wp); }

}  // extern "C"

// ---------------------------------------------------------------------------
// latency formulation (cabac_spec.cuh): rows in the context slots, both successor rows loaded ahead,
// LPS arm by table; driven with the schedule of k_encode_ops_lat / k_decode_ops_lat
// ---------------------------------------------------------------------------
#include "../../isscabac_b200/csrc/cabac_spec.cuh"

namespace {
struct HostSpecMem {
  SRow* slots;   // n_ctx + 1
  SRow ldrow(uint32_t tok) const { return spec_row(tok); }
  SRow ldctx(uint32_t c) const { return slots[c]; }
  void stctx(uint32_t c, const SRow& r) const { slots[c] = r; }
};
}  // namespace

extern "C" {

int emul_encode_ops_spec(uint32_t n_streams, const uint64_t* op_off, const uint8_t* ops,
                         const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                         uint8_t* slab, uint64_t stride, uint32_t* lens) {
  std::vector<SRow> cs(n_ctx + 1);
  int ovf = 0;
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) cs[c] = spec_row(ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c] & 127u);
    cs[n_ctx] = spec_row(0);
    HostSpecMem mem{cs.data()};
    EncWide E;
    encw_start(E, slab + s * stride, (uint32_t)stride);
    const uint8_t* p = ops + op_off[s];
    uint64_t n = op_off[s + 1] - op_off[s], i = 0;
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > n) head = n;
    for (; i < head; ++i) encs_general(E, p[i], mem, n_ctx);
    for (; i + 16 <= n; i += 16) {
      const uint32_t w[4] = {ld32(p + i), ld32(p + i + 4), ld32(p + i + 8), ld32(p + i + 12)};
      const uint32_t cw[4] = {op_codes4(w[0]), op_codes4(w[1]), op_codes4(w[2]), op_codes4(w[3])};
      if (block_has_trm(cw)) {
        for (int k = 0; k < 16; ++k) encs_general(E, p[i + k], mem, n_ctx);
      } else {
        encs_block16<false>(E, w, cw, mem, n_ctx);
      }
    }
    for (; i < n; ++i) encs_general(E, p[i], mem, n_ctx);
    lens[s] = encw_finish(E);
    ovf |= lens[s] > stride;
  }
  return ovf;
}

int emul_decode_ops_spec(uint32_t n_streams, const uint64_t* byte_off, const uint8_t* bytes,
                         const uint64_t* op_off, const uint8_t* ops,
                         const uint8_t* ctx_init, uint32_t n_ctx, int per_stream,
                         uint8_t* bins, uint8_t* ok) {
  std::vector<SRow> cs(n_ctx + 1);
  for (uint32_t s = 0; s < n_streams; ++s) {
    for (uint32_t c = 0; c < n_ctx; ++c) cs[c] = spec_row(ctx_init[(per_stream ? (uint64_t)s * n_ctx : 0) + c] & 127u);
    cs[n_ctx] = spec_row(0);
    HostSpecMem mem{cs.data()};
    DecWide D;
    decw_start(D, bytes + byte_off[s], (uint32_t)(byte_off[s + 1] - byte_off[s]));
    const uint8_t* p = ops + op_off[s];
    uint8_t* q = bins + op_off[s];
    uint64_t n = op_off[s + 1] - op_off[s], i = 0;
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > n) head = n;
    for (; i < head; ++i) q[i] = (uint8_t)decs_general(D, p[i], mem, n_ctx);
    for (; i + 16 <= n; i += 16) {
      const uint32_t cw[4] = {op_codes4(ld32(p + i)), op_codes4(ld32(p + i + 4)), op_codes4(ld32(p + i + 8)),
                              op_codes4(ld32(p + i + 12))};
      if (block_has_trm(cw)) {
        for (int k = 0; k < 16; ++k) q[i + k] = (uint8_t)decs_general(D, p[i + k], mem, n_ctx);
      } else {
        uint32_t r[4];
        decs_block16<false>(D, cw, r, mem, n_ctx);
        memcpy(q + i, r, 16);
      }
    }
    for (; i < n; ++i) q[i] = (uint8_t)decs_general(D, p[i], mem, n_ctx);
    ok[s] = (uint8_t)decw_finish(D);
  }
  return 0;
}

}  // extern "C"
