// symbols.cu -- symbol-level kernels: the binarizer + context selection of the MATLAB
// layer (cabacBinarizer.m, cabacContextSelection.m, cabacDemo.m:113-121,
// cabacDecodeSymbolFinished.m, cabacDebinarizer.m) moved onto the device.
//
//   cabac_binarize_symbols   symbol-parallel: every symbol's bin count is known in closed
//                            form, so ops are produced by count -> device-wide scan -> emit;
//                            context selection only looks at the symbol itself and its up
//                            neighbour (cabacContextSelection.m:24-67 never uses g_lft/g_up2).
//   cabac_encode_symbols     lane per stream, binarize + select + encode fused (no op array
//                            in HBM); optional getNumBits() trace per symbol.
//   cabac_decode_symbols     lane per stream: decodeBin with the finish detector and the
//                            context rule in the loop (inherently serial per stream), then
//                            the debinarized symbol is stored.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/isscabac.h"
#include "cabac_lane.cuh"
#include "internal.h"

using namespace cabac;
using namespace isscabac_internal;

namespace {

constexpr int NT = 128;
constexpr int TAB_WORDS = 128 * 32;

struct RowTable {
  uint2 r[128];
  constexpr RowTable() : r{} {
    for (uint32_t i = 0; i < 128; ++i) r[i] = fused_row(i);
  }
};
__constant__ RowTable c_rows_sym = RowTable();

__device__ __forceinline__ uint32_t load_sym(const void* p, int width, uint64_t i) {
  if (width == 1) return static_cast<const uint8_t*>(p)[i];
  if (width == 2) return static_cast<const uint16_t*>(p)[i];
  return static_cast<const uint32_t*>(p)[i];
}
__device__ __forceinline__ void store_sym(void* p, int width, uint64_t i, uint32_t v) {
  if (width == 1) static_cast<uint8_t*>(p)[i] = (uint8_t)v;
  else if (width == 2) static_cast<uint16_t*>(p)[i] = (uint16_t)v;
  else static_cast<uint32_t*>(p)[i] = v;
}

__device__ __forceinline__ SymCfg to_cfg(const isscabac_symcfg& c) {
  return SymCfg{c.profile, c.method, c.Nq, c.Nlbp, c.types, c.rows};
}

// ---- symbol-parallel binarizer ------------------------------------------------
__global__ void k_sym_count(isscabac_symcfg c, const void* sym, int width, uint64_t n, uint32_t* counts) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  counts[i] = sym_code(load_sym(sym, width, i), c.Nq, c.method).len;
}

__global__ void k_stream_op_off(const uint64_t* sym_off, const uint64_t* sym_op_off, uint64_t* op_off, uint32_t n_streams) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s <= n_streams) op_off[s] = sym_op_off[sym_off[s]];
}

__global__ void k_sym_emit(isscabac_symcfg c, const void* sym, int width, uint64_t n, const uint64_t* sym_off,
                           uint32_t n_streams, const uint64_t* sym_op_off, uint8_t* ops, uint64_t cap) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SymCfg cfg = to_cfg(c);
  // stream of symbol i: last s with sym_off[s] <= i
  uint32_t lo = 0, hi = n_streams;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (sym_off[mid] <= i) lo = mid; else hi = mid;
  }
  const uint64_t in_stream = i - sym_off[lo];
  const bool up = sym_has_up(cfg, in_stream);
  const SymCode code = sym_code(load_sym(sym, width, i), cfg.Nq, cfg.method);
  SymCode u = {0, 0, 0};
  if (up) u = sym_code(load_sym(sym, width, i - 1), cfg.Nq, cfg.method);
  uint64_t o = sym_op_off[i];
  for (uint32_t b = 1; b <= code.len; ++b, ++o) {
    int cx = select_ctx(cfg, b, code.np, u, up);
    uint32_t cd = cx < 0 ? ISSCABAC_OP8_EP : (uint32_t)cx;
    if (o < cap) ops[o] = (uint8_t)((cd << 1) | sym_bin(code, b));
  }
}

// ---- fused symbol encode / decode ----------------------------------------------
struct SymParams {
  isscabac_symcfg cfg;
  uint32_t n_streams, n_ctx;
  int per_stream_init, sym_width;
  const uint64_t* sym_off;
  const void* symbols;
  const uint8_t* ctx_init;
  uint8_t* slab;
  uint64_t slab_stride;
  uint32_t* lengths;
  uint32_t* bits_after;
  uint32_t* overflow;
  const uint64_t* byte_off;
  const uint8_t* bytes;
  void* out_symbols;
  uint8_t* finish_ok;
};

__device__ __forceinline__ uint32_t* setup_smem(const SymParams& P, uint8_t* smem, uint32_t s, bool valid) {
  uint2* tab = reinterpret_cast<uint2*>(smem);
  for (int i = threadIdx.x; i < TAB_WORDS; i += blockDim.x) tab[i] = c_rows_sym.r[i >> 5];
  uint32_t* ctx = reinterpret_cast<uint32_t*>(smem + TAB_WORDS * sizeof(uint2));
  const uint8_t* init = P.ctx_init + (P.per_stream_init && valid ? (uint64_t)s * P.n_ctx : 0);
  for (uint32_t c = 0; c < P.n_ctx; ++c) ctx[c * NT + threadIdx.x] = init[c];
  __syncthreads();
  return ctx + threadIdx.x;
}

template <bool TRACK>
__global__ void __launch_bounds__(NT) k_encode_symbols(SymParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t s = blockIdx.x * NT + threadIdx.x;
  const bool valid = s < P.n_streams;
  uint32_t* myctx = setup_smem(P, smem, valid ? s : 0, valid);
  if (!valid) return;
  const uint2* mytab = reinterpret_cast<const uint2*>(smem) + (threadIdx.x & 31);
  const SymCfg cfg = to_cfg(P.cfg);
  const uint32_t ctx_max = P.n_ctx ? P.n_ctx - 1 : 0;
  const uint64_t s0 = P.sym_off[s], s1 = P.sym_off[s + 1];
  EncLane L;
  uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  enc_start(L, P.slab + (uint64_t)s * P.slab_stride, cap);
  SymCode prev = {0, 0, 0};
  for (uint64_t i = s0; i < s1; ++i) {
    const SymCode code = sym_code(load_sym(P.symbols, P.sym_width, i), cfg.Nq, cfg.method);
    const bool up = sym_has_up(cfg, i - s0);
    for (uint32_t b = 1; b <= code.len; ++b) {
      const int cx = select_ctx(cfg, b, code.np, prev, up);
      const uint32_t bin = sym_bin(code, b);
      if (cx < 0) {
        enc_bin_ep<TRACK>(L, bin);
      } else {
        const uint32_t c = min((uint32_t)cx, ctx_max);
        uint32_t st = myctx[c * NT];
        enc_bin_ctx<TRACK>(L, bin, st, mytab[st * 32]);
        myctx[c * NT] = st;
      }
    }
    if (TRACK) P.bits_after[i] = enc_bits_written(L);
    prev = code;
  }
  enc_finish<TRACK>(L);
  enc_flush_pending(L);
  P.lengths[s] = L.nbytes;
  if ((L.overflow || L.nbytes > cap) && P.overflow) atomicOr(P.overflow, 1u);
}

__global__ void __launch_bounds__(NT) k_decode_symbols(SymParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t s = blockIdx.x * NT + threadIdx.x;
  const bool valid = s < P.n_streams;
  uint32_t* myctx = setup_smem(P, smem, valid ? s : 0, valid);
  if (!valid) return;
  const uint2* mytab = reinterpret_cast<const uint2*>(smem) + (threadIdx.x & 31);
  const SymCfg cfg = to_cfg(P.cfg);
  const uint32_t ctx_max = P.n_ctx ? P.n_ctx - 1 : 0;
  const uint64_t s0 = P.sym_off[s], s1 = P.sym_off[s + 1];
  const uint64_t b0 = P.byte_off[s], b1 = P.byte_off[s + 1];
  DecLane D;
  dec_start(D, P.bytes + b0, (uint32_t)(b1 - b0));
  SymCode prev = {0, 0, 0};
  for (uint64_t i = s0; i < s1; ++i) {
    const bool up = sym_has_up(cfg, i - s0);
    SymDec sd;
    symdec_reset(sd);
    uint32_t v = 0;
    for (;;) {
      const int cx = select_ctx(cfg, sd.n + 1, sd.np, prev, up);
      uint32_t bin;
      if (cx < 0) {
        bin = dec_bin_ep(D);
      } else {
        const uint32_t c = min((uint32_t)cx, ctx_max);
        uint32_t st = myctx[c * NT];
        bin = dec_bin_ctx(D, st, mytab[st * 32]);
        myctx[c * NT] = st;
      }
      if (symdec_push(sd, bin, cfg, v)) break;
    }
    store_sym(P.out_symbols, P.sym_width, i, v);
    prev = sym_code(v, cfg.Nq, cfg.method);
  }
  if (P.finish_ok) P.finish_ok[s] = (uint8_t)dec_finish(D);
}

int check_cfg(const isscabac_symcfg* cfg, uint32_t n_ctx, int sym_width, bool need_ctx) {
  if (!cfg) { set_error("symcfg is NULL"); return ISSCABAC_ERR_INVALID; }
  if (cfg->profile < 0 || cfg->profile > ISSCABAC_PROFILE_FLAT_EPSUF) { set_error("unknown profile %d", cfg->profile); return ISSCABAC_ERR_INVALID; }
  if (cfg->method < 0 || cfg->method > ISSCABAC_BIN_FL32) {
    // the truncated-Rice codes of cabacBinarizer.m:39-54 are incomplete upstream (escape is a TODO)
    // and cannot be decoded by the reference loops (cabacDecodeSymbolFinished.m has no case)
    set_error("binarization method %d not supported", cfg->method);
    return ISSCABAC_ERR_UNSUPPORTED;
  }
  if (cfg->Nlbp < 1 || cfg->Nlbp > 32) { set_error("Nlbp out of range"); return ISSCABAC_ERR_INVALID; }
  if (cfg->method == ISSCABAC_BIN_TU && cfg->Nq < 2) { set_error("TU needs Nq >= 2"); return ISSCABAC_ERR_INVALID; }
  if (sym_width != 1 && sym_width != 2 && sym_width != 4) { set_error("sym_width must be 1, 2 or 4"); return ISSCABAC_ERR_INVALID; }
  if (need_ctx) {
    int want = cabac_profile_num_ctx(cfg->profile, cfg->Nlbp);
    if ((int)n_ctx < want) { set_error("profile needs %d contexts, got %u", want, n_ctx); return ISSCABAC_ERR_INVALID; }
    if ((size_t)n_ctx * NT * 4 + TAB_WORDS * sizeof(uint2) > smem_limit()) { set_error("too many contexts for the symbol kernels"); return ISSCABAC_ERR_UNSUPPORTED; }
  }
  return ISSCABAC_OK;
}

template <class K>
int launch_sym(K kernel, const SymParams& P, cudaStream_t st, const char* name) {
  size_t smem = TAB_WORDS * sizeof(uint2) + (size_t)P.n_ctx * NT * 4;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<(P.n_streams + NT - 1) / NT, NT, smem, st>>>(P);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, name);
}

}  // namespace

extern "C" {

size_t cabac_binarize_scratch_bytes(uint64_t n_symbols, uint32_t n_streams) {
  (void)n_streams;
  size_t counts = ((size_t)n_symbols * 4 + 255) & ~(size_t)255;
  size_t offs = (((size_t)n_symbols + 1) * 8 + 255) & ~(size_t)255;
  uint64_t tiles = (n_symbols + 2047) / 2048;
  return counts + offs + (tiles + 1) * 8 + 512;
}

int cabac_binarize_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_sym_off,
                           const void* d_symbols, int sym_width, uint64_t n_symbols,
                           uint64_t* d_op_off, uint8_t* d_ops, uint64_t ops_cap,
                           void* d_scratch, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cfg(cfg, 0, sym_width, false);
  if (rc) return rc;
  if (!d_sym_off || !d_op_off || !d_scratch || (n_symbols && !d_symbols)) { set_error("cabac_binarize_symbols: null pointer"); return ISSCABAC_ERR_INVALID; }
  uint8_t* scr = static_cast<uint8_t*>(d_scratch);
  size_t counts_b = ((size_t)n_symbols * 4 + 255) & ~(size_t)255;
  size_t offs_b = (((size_t)n_symbols + 1) * 8 + 255) & ~(size_t)255;
  uint32_t* counts = reinterpret_cast<uint32_t*>(scr);
  uint64_t* sym_op_off = reinterpret_cast<uint64_t*>(scr + counts_b);
  void* scan_scr = scr + counts_b + offs_b;
  const uint32_t blocks = (uint32_t)((n_symbols + 255) / 256);
  if (n_symbols) k_sym_count<<<blocks, 256, 0, st>>>(*cfg, d_symbols, sym_width, n_symbols, counts);
  if ((rc = exclusive_scan_u32_u64(counts, sym_op_off, n_symbols, scan_scr, st))) return rc;
  k_stream_op_off<<<(n_streams + 256) / 256, 256, 0, st>>>(d_sym_off, sym_op_off, d_op_off, n_streams);
  if (d_ops && n_symbols)
    k_sym_emit<<<blocks, 256, 0, st>>>(*cfg, d_symbols, sym_width, n_symbols, d_sym_off, n_streams, sym_op_off, d_ops, ops_cap);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "cabac_binarize_symbols");
}

int cabac_encode_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_sym_off,
                         const void* d_symbols, int sym_width,
                         const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                         uint8_t* d_slab, uint64_t slab_stride, uint32_t* d_lengths,
                         uint32_t* d_bits_after_symbol, uint32_t* d_overflow, void* stream) {
  int rc = check_cfg(cfg, n_ctx, sym_width, true);
  if (rc) return rc;
  if (n_streams == 0) return ISSCABAC_OK;
  if (!d_sym_off || !d_slab || !d_lengths || !d_ctx_init) { set_error("cabac_encode_symbols: null pointer"); return ISSCABAC_ERR_INVALID; }
  if ((slab_stride & 15u) || slab_stride < 16 || (reinterpret_cast<uintptr_t>(d_slab) & 15u)) {
    set_error("cabac_encode_symbols: slab and slab_stride must be 16-byte aligned");
    return ISSCABAC_ERR_INVALID;
  }
  SymParams P;
  memset(&P, 0, sizeof P);
  P.cfg = *cfg; P.n_streams = n_streams; P.n_ctx = n_ctx; P.per_stream_init = per_stream_init; P.sym_width = sym_width;
  P.sym_off = d_sym_off; P.symbols = d_symbols; P.ctx_init = d_ctx_init;
  P.slab = d_slab; P.slab_stride = slab_stride; P.lengths = d_lengths; P.bits_after = d_bits_after_symbol; P.overflow = d_overflow;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return d_bits_after_symbol ? launch_sym(k_encode_symbols<true>, P, st, "k_encode_symbols")
                             : launch_sym(k_encode_symbols<false>, P, st, "k_encode_symbols");
}

int cabac_decode_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_byte_off,
                         const uint8_t* d_bytes, const uint64_t* d_sym_off,
                         const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                         void* d_symbols, int sym_width, uint8_t* d_finish_ok, void* stream) {
  int rc = check_cfg(cfg, n_ctx, sym_width, true);
  if (rc) return rc;
  if (n_streams == 0) return ISSCABAC_OK;
  if (!d_sym_off || !d_byte_off || !d_bytes || !d_symbols || !d_ctx_init) { set_error("cabac_decode_symbols: null pointer"); return ISSCABAC_ERR_INVALID; }
  SymParams P;
  memset(&P, 0, sizeof P);
  P.cfg = *cfg; P.n_streams = n_streams; P.n_ctx = n_ctx; P.per_stream_init = per_stream_init; P.sym_width = sym_width;
  P.sym_off = d_sym_off; P.ctx_init = d_ctx_init; P.byte_off = d_byte_off; P.bytes = d_bytes;
  P.out_symbols = d_symbols; P.finish_ok = d_finish_ok;
  return launch_sym(k_decode_symbols, P, static_cast<cudaStream_t>(stream), "k_decode_symbols");
}

}  // extern "C"
