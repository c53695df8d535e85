// iss_stats.cu -- the step right before encode in the ISS flow: empirical p(0) per context
// from the binarised matrix (ISS/+coder/cabacInitContextModel.m:15-129), as a device-side
// reduction + a host-side finalisation in double precision (cabacEncode.m:23-31: statistics,
// `equalProb`, uint8(ctxInit*255) side-info quantisation, /255, initByProb's state mapping).
//
// Every counter of the MATLAB code is a sum over symbols of a 0/1 term that depends only on
// the symbol's own bin string and on the string of the symbol above it in the same column
// (Gbin_up1, :16), so a thread evaluates the terms of four consecutive symbols from the closed-form codes
// (sym_code / sym_bin), keeps them as packed 8-bit counters in registers, and the warp adds them up with REDUX
// before touching global memory.
// Quirks kept on purpose: the first row's neighbour is a NaN cell of length 1 with np = 0
// (:16,:21,:23), so it takes part in the conds0/conds1 denominators for n = 1; the conds
// normalisers index the neighbour at the absolute position n, not n + np_up1 (:90,:102); the
// "rest" suffix statistic starts at the absolute bin position Nlbp+1 (:121-126).
#include <cuda_runtime.h>
#include <mutex>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "../../include/isscabac.h"
#include "cabac_lane.cuh"
#include "internal.h"

using namespace cabac;
using namespace isscabac_internal;

namespace {

// counters per modelled bin position n (1..Nlbp), then 4 for the two "rest" contexts
enum { PRE_HIT, PRE_TOT, C_TOT, C0_HIT, C0N_HIT, C1_HIT, C1N_HIT, BL_TOT, BL_HIT, BLN_HIT,
       SUF_HIT, SUF_TOT, CS_TOT, S0_HIT, S0N_HIT, S1_HIT, S1N_HIT, PER_N };
enum { RP_HIT, RP_TOT, RS_HIT, RS_TOT, N_REST };
constexpr int MAX_NLBP = 8;

__device__ __forceinline__ uint32_t load_sym(const void* p, int width, uint64_t i) {
  if (width == 1) return static_cast<const uint8_t*>(p)[i];
  if (width == 2) return static_cast<const uint16_t*>(p)[i];
  return static_cast<const uint32_t*>(p)[i];
}

// The 0/1 terms of one symbol (x = its code, u = the code of the symbol above it, hasup = there is one): add(k, v) adds v
// to counter k of the symbol's group.  The loop over the modelled positions is unrolled with constant counter indices, so
// that a caller may keep the counters in registers.
template <class F>
__device__ __forceinline__ void iss_terms(const SymCode& x, const SymCode& u, bool hasup, bool live, int N, F&& add) {
  // np = position of the first 0, or the length when there is none (:18-20)
  const uint32_t L = x.len, np = x.np <= x.len ? x.np : x.len;
  const uint32_t L_up = hasup ? u.len : 1u;                                  // :23
  const uint32_t np_up = hasup ? (u.np <= u.len ? u.np : u.len) : 0u;        // :21
  const uint32_t lv = live ? 1u : 0u;
#pragma unroll
  for (int n = 1; n <= MAX_NLBP; ++n) {
    if (n > N) break;
    const int k0 = PER_N * (n - 1);
    const uint32_t un = (uint32_t)n;
    const bool in_pre = lv && un <= np;
    const uint32_t xn = in_pre ? sym_bin(x, un) : 1u;
    add(k0 + PRE_TOT, (uint32_t)in_pre);                                      // :31-33
    add(k0 + PRE_HIT, (uint32_t)(in_pre && xn == 0));
    const bool csel = in_pre && un <= np_up;                                  // :38,:50
    const uint32_t yn = (lv && hasup && un <= L_up) ? sym_bin(u, un) : 2u;    // 2 = NaN / out of range
    add(k0 + C_TOT, (uint32_t)csel);
    add(k0 + C0_HIT, (uint32_t)(csel && xn == 0 && yn == 0));
    add(k0 + C0N_HIT, (uint32_t)(csel && yn == 0));
    add(k0 + C1_HIT, (uint32_t)(csel && xn == 0 && yn == 1));
    add(k0 + C1N_HIT, (uint32_t)(csel && yn == 1));
    const bool bsel = lv && un + 1 <= np && np_up < un + 1;                   // :62-72
    add(k0 + BL_TOT, (uint32_t)bsel);
    add(k0 + BL_HIT, (uint32_t)(bsel && sym_bin(x, un + 1) == 0 && xn == 1));
    add(k0 + BLN_HIT, (uint32_t)(bsel && xn == 1));
    const bool ssel = lv && un + np <= L;                                     // :77-80
    const uint32_t xs = ssel ? sym_bin(x, un + np) : 1u;
    add(k0 + SUF_TOT, (uint32_t)ssel);
    add(k0 + SUF_HIT, (uint32_t)(ssel && xs == 0));
    const bool cssel = ssel && un + np_up <= L_up;                            // :84-106
    const uint32_t ys = (cssel && hasup) ? sym_bin(u, un + np_up) : 2u;
    add(k0 + CS_TOT, (uint32_t)cssel);
    add(k0 + S0_HIT, (uint32_t)(cssel && xs == 0 && ys == 0));
    add(k0 + S0N_HIT, (uint32_t)(cssel && yn == 0));
    add(k0 + S1_HIT, (uint32_t)(cssel && xs == 0 && ys == 1));
    add(k0 + S1N_HIT, (uint32_t)(cssel && yn == 1));
  }
}
// the two "rest" contexts (:111-126): bins from absolute position Nlbp+1 on -> {RP_HIT, RP_TOT, RS_HIT, RS_TOT}
__device__ __forceinline__ void iss_rest(const SymCode& x, bool live, int N, uint32_t r[N_REST]) {
  const uint32_t L = x.len, np = x.np <= x.len ? x.np : x.len;
  const uint32_t n0 = (uint32_t)N + 1u;
  r[RP_HIT] = r[RP_TOT] = r[RS_HIT] = r[RS_TOT] = 0;
  if (live && n0 <= np) {
    r[RP_TOT] = np - (uint32_t)N;
    r[RP_HIT] = sym_bin(x, np) == 0 ? 1u : 0u;   // bins n0..np-1 are prefix ones
  } else if (live && n0 <= L) {
    r[RS_TOT] = L - (uint32_t)N;
    for (uint32_t b = n0; b <= L; ++b) r[RS_HIT] += sym_bin(x, b) == 0;
  }
}

// Four consecutive symbols per thread; the 0/1 counters are kept in registers as 8-bit fields (four counters per word: a
// thread adds at most 4, a warp at most 128 per field), so a warp of 128 symbols needs one REDUX per word and one global
// atomic per counter -- a thirteenth of the reductions and a quarter of the atomics of one symbol per thread.  A warp whose
// symbols span two groups (one in 8,000 symbols for an ISS matrix) adds its non-zero terms one by one.
constexpr int ISS_SPT = 4;
__global__ void __launch_bounds__(256) k_iss_ctx_stats(isscabac_symcfg c, const void* sym, int width, uint64_t n_sym,
                                                        const uint64_t* sym_off, uint32_t n_streams, uint32_t per_group,
                                                        unsigned long long* counters) {
  const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * ISS_SPT;
  const int N = c.Nlbp;
  const uint32_t K = (uint32_t)(PER_N * N + N_REST);
  uint32_t s = 0;
  uint64_t s_begin = 0, s_end = 0;
  if (i0 < n_sym) {
    uint32_t lo = 0, hi = n_streams;   // stream of symbol i0: last s with sym_off[s] <= i0
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (sym_off[mid] <= i0) lo = mid; else hi = mid;
    }
    s = lo;
    s_begin = sym_off[s];
    s_end = sym_off[s + 1];
  }
  // groups of this thread's first and last live symbol
  uint32_t g_first = 0xffffffffu, g_last = 0xffffffffu;
  if (i0 < n_sym) {
    g_first = s / per_group;
    uint64_t il = i0 + ISS_SPT - 1 < n_sym ? i0 + ISS_SPT - 1 : n_sym - 1;
    uint32_t sl = s;
    while (sl + 1 < n_streams && sym_off[sl + 1] <= il) ++sl;
    g_last = sl / per_group;
  }
  const uint32_t g0 = __shfl_sync(0xffffffffu, g_first, 0);
  const bool uniform = g0 != 0xffffffffu && __all_sync(0xffffffffu, i0 >= n_sym || (g_first == g0 && g_last == g0));
  uint32_t acc[PER_N * MAX_NLBP / 4 + 1];
#pragma unroll
  for (int w = 0; w < PER_N * MAX_NLBP / 4 + 1; ++w) acc[w] = 0;
  uint32_t rest[N_REST] = {0, 0, 0, 0};
  uint32_t prev_v = 0;
  if (i0 < n_sym && i0 > 0) prev_v = load_sym(sym, width, i0 - 1);
#pragma unroll
  for (int j = 0; j < ISS_SPT; ++j) {
    const uint64_t i = i0 + j;
    const bool live = i < n_sym;
    uint32_t v = 0;
    bool hasup = false;
    if (live) {
      while (i >= s_end && s + 1 < n_streams) { ++s; s_begin = s_end; s_end = sym_off[s + 1]; }
      const uint64_t in_stream = i - s_begin;
      hasup = c.rows ? (in_stream % c.rows) != 0 : in_stream > 0;
      v = load_sym(sym, width, i);
    }
    const SymCode x = live ? sym_code(v, c.Nq, c.method) : SymCode{0, 0, 0};
    const SymCode u = hasup ? sym_code(prev_v, c.Nq, c.method) : SymCode{0, 0, 0};
    prev_v = v;
    uint32_t r[N_REST];
    iss_rest(x, live, N, r);
    if (uniform) {
      iss_terms(x, u, hasup, live, N, [&](int k, uint32_t t) { acc[k >> 2] += t << (8 * (k & 3)); });
#pragma unroll
      for (int q = 0; q < N_REST; ++q) rest[q] += r[q];
    } else if (live) {
      unsigned long long* cnt = counters + (size_t)(s / per_group) * K;
      iss_terms(x, u, hasup, live, N, [&](int k, uint32_t t) { if (t) atomicAdd(cnt + k, (unsigned long long)t); });
      for (int q = 0; q < N_REST; ++q) if (r[q]) atomicAdd(cnt + PER_N * N + q, (unsigned long long)r[q]);
    }
  }
  if (!uniform) return;
  unsigned long long* cnt = counters + (size_t)g0 * K;
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int w = 0; w < PER_N * MAX_NLBP / 4 + 1; ++w) {
    if (4 * w >= PER_N * N) break;
    // fields cannot carry into each other: <= 128 per field
    const uint32_t sum = __reduce_add_sync(0xffffffffu, acc[w]);
    if (lane < 4) {
      const uint32_t f = (sum >> (8 * lane)) & 0xffu;
      const int k = 4 * w + (int)lane;
      if (f && k < PER_N * N) atomicAdd(cnt + k, (unsigned long long)f);
    }
  }
#pragma unroll
  for (int q = 0; q < N_REST; ++q) {
    const uint32_t sum = __reduce_add_sync(0xffffffffu, rest[q]);
    if (lane == 0 && sum) atomicAdd(cnt + PER_N * N + q, (unsigned long long)sum);
  }
}

__host__ __device__ inline double frac(unsigned long long hit, unsigned long long tot) { return tot ? (double)hit / (double)tot : 0.0; }

// counters of one group -> p(0) of its 7*Nlbp+2 contexts (cabacInitContextModel.m:128); host and device run this same code:
// two IEEE double divisions per context, so the results are identical
__host__ __device__ inline void group_p0(const isscabac_symcfg& cfg, const unsigned long long* c, double* p) {
  const int N = cfg.Nlbp;
  const int nctx = 7 * N + 2;
  for (int i = 0; i < nctx; ++i) p[i] = 0.0;
  for (int n = 0; n < N; ++n) {
    const unsigned long long* cn = c + PER_N * n;
    p[n] = frac(cn[PRE_HIT], cn[PRE_TOT]);
    // joint / marginal, the marginal replaced by 1 when it is empty or zero (:41-45 etc.)
    auto cond = [&](unsigned long long hit, unsigned long long norm_hit, unsigned long long tot) {
      const double nr = norm_hit > 0 ? frac(norm_hit, tot) : 1.0;
      return frac(hit, tot) / nr;
    };
    if (cfg.types & ISSCABAC_CM_COND0) p[N + n] = cond(cn[C0_HIT], cn[C0N_HIT], cn[C_TOT]);
    if (cfg.types & ISSCABAC_CM_COND1) p[2 * N + n] = cond(cn[C1_HIT], cn[C1N_HIT], cn[C_TOT]);
    if (cfg.types & ISSCABAC_CM_CONDBINLFT) p[3 * N + n] = cond(cn[BL_HIT], cn[BLN_HIT], cn[BL_TOT]);
    p[4 * N + n] = frac(cn[SUF_HIT], cn[SUF_TOT]);
    if (cfg.types & ISSCABAC_CM_CONDS0) p[5 * N + n] = cond(cn[S0_HIT], cn[S0N_HIT], cn[CS_TOT]);
    if (cfg.types & ISSCABAC_CM_CONDS1) p[6 * N + n] = cond(cn[S1_HIT], cn[S1N_HIT], cn[CS_TOT]);
  }
  const unsigned long long* cr = c + PER_N * N;
  p[7 * N] = frac(cr[RP_HIT], cr[RP_TOT]);
  p[7 * N + 1] = frac(cr[RS_HIT], cr[RS_TOT]);
}

// uint8(v * 255) as MATLAB does it: round half away from zero (v >= 0), saturate; one expression for host and device
// (NaN from degenerate counters -> 0 on both)
__host__ __device__ inline double quant255(double v) {
  double q = floor(v * 255.0 + 0.5);
  q = (255.0 < q) ? 255.0 : q;
  return (0.0 < q) ? q : 0.0;
}

// state byte of every uint8 side-information value q (p0 = q / 255): filled once per process from the host function that is
// pinned to the reference (cabac_ctx_from_prob), so the device never evaluates log10 itself
__constant__ uint8_t c_state_of_q[256];

__global__ void k_iss_ctx_finalise(isscabac_symcfg cfg, const unsigned long long* counters, uint32_t n_groups, int equal_prob,
                                   double* p0, uint8_t* ctx_quant, uint8_t* ctx_state) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const int nctx = 7 * cfg.Nlbp + 2;
  const int K = PER_N * cfg.Nlbp + N_REST;
  double p[7 * MAX_NLBP + 2];
  group_p0(cfg, counters + (size_t)g * K, p);
  for (int i = 0; i < nctx; ++i) {
    const double v = equal_prob ? 0.5 : p[i];
    const double q = quant255(v);
    if (p0) p0[(size_t)g * nctx + i] = v;
    if (ctx_quant) ctx_quant[(size_t)g * nctx + i] = (uint8_t)q;
    if (ctx_state) ctx_state[(size_t)g * nctx + i] = c_state_of_q[(int)q];
  }
}

}  // namespace

extern "C" {

int cabac_iss_num_counters(int Nlbp) {
  if (Nlbp < 1 || Nlbp > MAX_NLBP) return ISSCABAC_ERR_INVALID;
  return PER_N * Nlbp + N_REST;
}

int cabac_iss_ctx_stats(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_sym_off,
                        const void* d_symbols, int sym_width, uint64_t n_symbols, uint32_t streams_per_group,
                        uint64_t* d_counters, void* stream) {
  if (!cfg || !d_sym_off || !d_counters || (n_symbols && !d_symbols)) { set_error("cabac_iss_ctx_stats: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (cfg->Nlbp < 1 || cfg->Nlbp > MAX_NLBP) { set_error("Nlbp must be 1..%d", MAX_NLBP); return ISSCABAC_ERR_INVALID; }
  if (cfg->method < 0 || cfg->method > ISSCABAC_BIN_FL32) { set_error("binarization method %d not supported", cfg->method); return ISSCABAC_ERR_UNSUPPORTED; }
  if (sym_width != 1 && sym_width != 2 && sym_width != 4) { set_error("sym_width must be 1, 2 or 4"); return ISSCABAC_ERR_INVALID; }
  if (streams_per_group == 0) streams_per_group = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t groups = (n_streams + streams_per_group - 1) / streams_per_group;
  const size_t K = (size_t)cabac_iss_num_counters(cfg->Nlbp);
  CK(cudaMemsetAsync(d_counters, 0, groups * K * sizeof(uint64_t), st));
  if (n_symbols == 0 || n_streams == 0) return ISSCABAC_OK;
  const uint32_t blocks = (uint32_t)((n_symbols + 256ull * ISS_SPT - 1) / (256ull * ISS_SPT));
  k_iss_ctx_stats<<<blocks, 256, 0, st>>>(*cfg, d_symbols, sym_width, n_symbols, d_sym_off, n_streams, streams_per_group,
                                          reinterpret_cast<unsigned long long*>(d_counters));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_iss_ctx_stats");
}

// counters -> p(0) per context (7*Nlbp+2 doubles per group, cabacInitContextModel.m:128), then
// cabacEncode.m:26-31: equalProb override, uint8(p*255) (MATLAB rounding: half away from zero,
// saturating), p = q/255, and the state bytes initByProb would derive from those p
// (CABAC_ContextModelsInit.cpp:124-148).  Any output pointer may be NULL.
int cabac_iss_ctx_from_counters(const isscabac_symcfg* cfg, const uint64_t* h_counters, uint32_t n_groups,
                                int equal_prob, double* h_p0, uint8_t* h_ctx_quant, uint8_t* h_ctx_state) {
  if (!cfg || (n_groups && !h_counters)) { set_error("cabac_iss_ctx_from_counters: null pointer"); return ISSCABAC_ERR_INVALID; }
  const int N = cfg->Nlbp;
  const int K = cabac_iss_num_counters(N);
  if (K < 0) { set_error("Nlbp must be 1..%d", MAX_NLBP); return ISSCABAC_ERR_INVALID; }
  const int nctx = 7 * N + 2;
  for (uint32_t g = 0; g < n_groups; ++g) {
    const unsigned long long* c = reinterpret_cast<const unsigned long long*>(h_counters) + (size_t)g * K;
    double p[7 * MAX_NLBP + 2];
    group_p0(*cfg, c, p);
    double pq[7 * MAX_NLBP + 2];
    for (int i = 0; i < nctx; ++i) {
      const double v = equal_prob ? 0.5 : p[i];
      const double q = quant255(v);
      if (h_p0) h_p0[(size_t)g * nctx + i] = v;
      if (h_ctx_quant) h_ctx_quant[(size_t)g * nctx + i] = (uint8_t)q;
      pq[i] = q / 255.0;
    }
    if (h_ctx_state) {
      int rc = cabac_ctx_from_prob(pq, (uint32_t)nctx, h_ctx_state + (size_t)g * nctx);
      if (rc) return rc;
    }
  }
  return ISSCABAC_OK;
}

// The same on the device (no host round trip between the statistics and the encoder): d_counters as cabac_iss_ctx_stats
// leaves them; outputs are device arrays [n_groups][7*Nlbp+2], any of them may be NULL.
int cabac_iss_ctx_from_counters_device(const isscabac_symcfg* cfg, const uint64_t* d_counters, uint32_t n_groups,
                                       int equal_prob, double* d_p0, uint8_t* d_ctx_quant, uint8_t* d_ctx_state, void* stream) {
  if (!cfg || (n_groups && !d_counters)) { set_error("cabac_iss_ctx_from_counters_device: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (cabac_iss_num_counters(cfg->Nlbp) < 0) { set_error("Nlbp must be 1..%d", MAX_NLBP); return ISSCABAC_ERR_INVALID; }
  if (n_groups == 0) return ISSCABAC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool lut_ready[64] = {};
  static std::mutex lut_mutex;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device index out of range"); return ISSCABAC_ERR_UNSUPPORTED; }
  std::lock_guard<std::mutex> lut_lock(lut_mutex);
  if (!lut_ready[dev]) {
    double pq[256];
    uint8_t lut[256];
    for (int q = 0; q < 256; ++q) pq[q] = q / 255.0;
    int rc = cabac_ctx_from_prob(pq, 256, lut);
    if (rc) return rc;
    CK(cudaMemcpyToSymbol(c_state_of_q, lut, sizeof lut));
    lut_ready[dev] = true;
  }
  k_iss_ctx_finalise<<<(n_groups + 127) / 128, 128, 0, st>>>(*cfg, reinterpret_cast<const unsigned long long*>(d_counters), n_groups,
                                                             equal_prob, d_p0, d_ctx_quant, d_ctx_state);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_iss_ctx_finalise");
}

}  // extern "C"
