// host_api.cu -- library housekeeping, context initialisation and the HOST-buffer
// entry points (what a reference-side caller binds: plain host arrays in, bytes out).
// Host<->device copies happen inside these calls; the stream set is chunked over a few
// CUDA streams so that the H2D copy of chunk k+1 overlaps the coding of chunk k.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/isscabac.h"
#include "internal.h"

using namespace isscabac_internal;

extern "C" {

int isscabac_version(void) { return ISSCABAC_VERSION; }

const char* isscabac_strerror(int code) {
  switch (code) {
    case ISSCABAC_OK: return "ok";
    case ISSCABAC_ERR_INVALID: return "invalid argument";
    case ISSCABAC_ERR_CUDA: return "CUDA error (no device, or a runtime failure)";
    case ISSCABAC_ERR_OVERFLOW: return "output buffer too small";
    case ISSCABAC_ERR_NOMEM: return "out of memory";
    case ISSCABAC_ERR_UNSUPPORTED: return "unsupported";
    case ISSCABAC_ERR_STATE: return "handle used out of order";
    case ISSCABAC_ERR_IO: return "bitstream file access error";
    case ISSCABAC_ERR_CORRUPT: return "bitstream not terminated properly";
  }
  return "unknown error";
}

const char* isscabac_last_error(void) { return g_err; }

int isscabac_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return ISSCABAC_OK;
}

// p(0) -> context state byte.  Restates xMapProbabilityToState
// (CABAC/CABAC_ContextModelsInit.cpp:124-148) in double precision on the host, so that
// the result is bit-identical to the reference: pLPS = min(p0, 1-p0) floored at 0.01875,
// MPS = (p0 < 0.5), state = clip(round(62*log10(2 pLPS)/log10(2*0.01875)), 0, 62).
int cabac_ctx_from_prob(const double* p0, uint32_t n, uint8_t* ctx_out) {
  if (n && (!p0 || !ctx_out)) { set_error("cabac_ctx_from_prob: null pointer"); return ISSCABAC_ERR_INVALID; }
  const double denom = log10(2.0 * 0.01875);
  for (uint32_t i = 0; i < n; ++i) {
    const double p = p0[i];
    if (!(p >= 0.0 && p <= 1.0)) { set_error("cabac_ctx_from_prob: p0[%u] outside [0,1]", i); return ISSCABAC_ERR_INVALID; }
    const unsigned mps = p >= 0.5 ? 0u : 1u;
    double plps = mps ? p : 1.0 - p;
    if (plps < 0.01875) plps = 0.01875;
    int st = (int)round(62 * log10(2.0 * plps) / denom);
    st = std::max(0, std::min(st, 62));
    ctx_out[i] = (uint8_t)((st << 1) | mps);
  }
  return ISSCABAC_OK;
}

// [ctxIdx mps state] triples (CABAC_ContextModelsInit.cpp:51-80; ctxIdx is ignored there too)
int cabac_ctx_from_state(const double* t, uint32_t n, uint8_t* ctx_out) {
  if (n && (!t || !ctx_out)) { set_error("cabac_ctx_from_state: null pointer"); return ISSCABAC_ERR_INVALID; }
  for (uint32_t i = 0; i < n; ++i) {
    const unsigned mps = (unsigned)t[3 * i + 1], st = (unsigned)t[3 * i + 2];
    ctx_out[i] = (uint8_t)((st << 1) + mps);
  }
  return ISSCABAC_OK;
}

int cabac_profile_num_ctx(int profile, int Nlbp) {
  switch (profile) {
    case ISSCABAC_PROFILE_DEMO: return 3;
    case ISSCABAC_PROFILE_ISS: return 7 * Nlbp + 2;
    case ISSCABAC_PROFILE_FLAT: return 2 * Nlbp + 2;
    case ISSCABAC_PROFILE_FLAT_EPSUF: return Nlbp + 1;
  }
  return ISSCABAC_ERR_INVALID;
}

int cabac_host_alloc(void** p, size_t bytes) {
  if (!p) return ISSCABAC_ERR_INVALID;
  CK(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
  return ISSCABAC_OK;
}
int cabac_host_free(void* p) {
  if (p) CK(cudaFreeHost(p));
  return ISSCABAC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// host-buffer pipelines
// ---------------------------------------------------------------------------
namespace {

constexpr int kChunks = 16;  // stream groups in flight
constexpr int kLanes = 4;    // CUDA streams

struct Pipeline {
  cudaStream_t s[kLanes] = {};
  cudaEvent_t done[kLanes] = {};
  bool ok = false;
  int init() {
    if (ok) return ISSCABAC_OK;
    for (int i = 0; i < kLanes; ++i) {
      CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    // keep freed blocks cached in the stream-ordered pool between calls
    int dev = 0;
    CK(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = ~0ull;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    ok = true;
    return ISSCABAC_OK;
  }
};
// one pipeline per (host thread, device): streams and events belong to the device that was current when they were made
constexpr int kMaxDevs = 64;
thread_local Pipeline g_pipes[kMaxDevs];
Pipeline* current_pipe() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevs) {
    set_error("no current CUDA device (or device index out of range)");
    return nullptr;
  }
  return &g_pipes[dev];
}
// Declared AFTER the DevBufs of a host entry point, so it is destroyed BEFORE them: whatever path leaves the
// function -- an error in the middle of the chunk loop included -- every lane is idle before a buffer is handed
// back to the pool on lane 0 (no stream-ordered use-after-free on error paths).
struct Drain {
  Pipeline& p;
  ~Drain() {
    for (int l = 0; l < kLanes; ++l)
      if (p.s[l]) cudaStreamSynchronize(p.s[l]);
  }
};

struct DevBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  int alloc(size_t bytes, cudaStream_t s) {
    st = s;
    CK(cudaMallocAsync(&p, bytes ? bytes : 16, s));
    return ISSCABAC_OK;
  }
  ~DevBuf() { if (p) cudaFreeAsync(p, st); }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

// chunk boundaries in stream index space, balanced by op count
void make_chunks(uint32_t n_streams, const uint64_t* off, int want, std::vector<uint32_t>& b) {
  b.clear();
  b.push_back(0);
  const uint64_t total = off[n_streams] - off[0];
  for (int k = 1; k < want; ++k) {
    uint64_t target = off[0] + total * k / want;
    uint32_t s = (uint32_t)(std::lower_bound(off, off + n_streams + 1, target) - off);
    s = (s + 127u) & ~127u;  // whole CTAs
    if (s > n_streams) s = n_streams;
    if (s > b.back()) b.push_back(s);
  }
  if (b.back() != n_streams) b.push_back(n_streams);
}

}  // namespace

extern "C" {

int cabac_encode_ops_host(uint32_t n_streams, const uint64_t* h_op_off, const void* h_ops, int op_width,
                          const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                          uint8_t* h_payload, uint64_t payload_cap, uint64_t* h_byte_off) {
  if (!h_op_off || !h_byte_off || (n_streams && !h_payload)) { set_error("cabac_encode_ops_host: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (op_width != 1 && op_width != 2) { set_error("op_width must be 1 or 2"); return ISSCABAC_ERR_INVALID; }
  if (n_streams == 0) { h_byte_off[0] = 0; return ISSCABAC_OK; }
  Pipeline* pipe = current_pipe();
  if (!pipe) return ISSCABAC_ERR_CUDA;
  Pipeline& g_pipe = *pipe;
  int rc = g_pipe.init();
  if (rc) return rc;
  cudaStream_t s0 = g_pipe.s[0];
  const uint64_t base = h_op_off[0], total = h_op_off[n_streams] - base;
  uint64_t max_ops = 0;
  for (uint32_t s = 0; s < n_streams; ++s) max_ops = std::max(max_ops, h_op_off[s + 1] - h_op_off[s]);
  // practical stride first (2 bits/op; adaptive CABAC never exceeds ~1 bit/bin), proof-level bound on retry
  uint64_t stride = ((max_ops / 4 + 64) + 15) & ~15ull;
  const size_t ctx_bytes = (size_t)n_ctx * (per_stream_init ? n_streams : 1);

  DevBuf d_ops, d_off, d_ctx, d_len, d_boff, d_scr, d_flag;
  Drain drain{g_pipe};
  if ((rc = d_ops.alloc(total * op_width, s0)) || (rc = d_off.alloc((n_streams + 1ull) * 8, s0)) ||
      (rc = d_ctx.alloc(ctx_bytes, s0)) || (rc = d_len.alloc(n_streams * 4ull, s0)) ||
      (rc = d_boff.alloc((n_streams + 1ull) * 8, s0)) ||
      (rc = d_scr.alloc(cabac_compact_scratch_bytes(n_streams), s0)) || (rc = d_flag.alloc(16, s0)))
    return rc;
  // op offsets are rebased so that they index the device op buffer
  std::vector<uint64_t> off_rel;
  const uint64_t* off_src = h_op_off;
  if (base) {
    off_rel.resize(n_streams + 1ull);
    for (uint32_t s = 0; s <= n_streams; ++s) off_rel[s] = h_op_off[s] - base;
    off_src = off_rel.data();
  }
  CK(cudaMemcpyAsync(d_off.p, off_src, (n_streams + 1ull) * 8, cudaMemcpyHostToDevice, s0));
  if (ctx_bytes) CK(cudaMemcpyAsync(d_ctx.p, h_ctx_init, ctx_bytes, cudaMemcpyHostToDevice, s0));
  CK(cudaEventRecord(g_pipe.done[0], s0));
  std::vector<uint32_t> cb;
  make_chunks(n_streams, off_src, kChunks, cb);

  for (int attempt = 0; attempt < 2; ++attempt) {
    DevBuf d_slab, d_payload;
    Drain drain_inner{g_pipe};
    if ((rc = d_slab.alloc((size_t)n_streams * stride, s0))) return rc;
    CK(cudaMemsetAsync(d_flag.p, 0, 16, s0));
    CK(cudaEventRecord(g_pipe.done[0], s0));
    for (int l = 1; l < kLanes; ++l) CK(cudaStreamWaitEvent(g_pipe.s[l], g_pipe.done[0], 0));
    for (size_t k = 0; k + 1 < cb.size(); ++k) {
      cudaStream_t st = g_pipe.s[k % kLanes];
      const uint32_t a = cb[k], b = cb[k + 1];
      const uint64_t oa = off_src[a], ob = off_src[b];
      if (attempt == 0 && ob > oa)
        CK(cudaMemcpyAsync(d_ops.as<uint8_t>() + oa * op_width, (const uint8_t*)h_ops + (base + oa) * op_width,
                           (ob - oa) * op_width, cudaMemcpyHostToDevice, st));
      rc = cabac_encode_ops(b - a, d_off.as<uint64_t>() + a, d_ops.p, op_width,
                            d_ctx.as<uint8_t>() + (per_stream_init ? (size_t)a * n_ctx : 0), n_ctx, per_stream_init,
                            d_slab.as<uint8_t>() + (size_t)a * stride, stride, d_len.as<uint32_t>() + a,
                            d_flag.as<uint32_t>(), st);
      if (rc) return rc;
    }
    for (int l = 1; l < kLanes; ++l) {
      CK(cudaEventRecord(g_pipe.done[l], g_pipe.s[l]));
      CK(cudaStreamWaitEvent(s0, g_pipe.done[l], 0));
    }
    // offsets first (the total sizes the payload buffer), then the copy
    if ((rc = exclusive_scan_u32_u64(d_len.as<uint32_t>(), d_boff.as<uint64_t>(), n_streams, d_scr.p, s0))) return rc;
    uint32_t flag = 0;
    uint64_t total_bytes = 0;
    CK(cudaMemcpyAsync(&flag, d_flag.p, 4, cudaMemcpyDeviceToHost, s0));
    CK(cudaMemcpyAsync(&total_bytes, d_boff.as<uint64_t>() + n_streams, 8, cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    if (flag & 1u) {
      if (attempt == 1) { set_error("slab overflow with the proof-level stride"); return ISSCABAC_ERR_OVERFLOW; }
      stride = cabac_slab_stride_bound(max_ops);
      continue;
    }
    if (total_bytes > payload_cap) {
      set_error("payload needs %llu bytes, capacity %llu", (unsigned long long)total_bytes, (unsigned long long)payload_cap);
      return ISSCABAC_ERR_OVERFLOW;
    }
    if ((rc = d_payload.alloc((total_bytes + 3) & ~(size_t)3, s0))) return rc;
    rc = cabac_compact(n_streams, d_slab.as<uint8_t>(), stride, d_len.as<uint32_t>(), d_payload.as<uint8_t>(), total_bytes,
                       d_boff.as<uint64_t>(), d_scr.p, d_flag.as<uint32_t>(), s0);
    if (rc) return rc;
    if (total_bytes) CK(cudaMemcpyAsync(h_payload, d_payload.p, total_bytes, cudaMemcpyDeviceToHost, s0));
    CK(cudaMemcpyAsync(h_byte_off, d_boff.p, (n_streams + 1ull) * 8, cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    return ISSCABAC_OK;
  }
  return ISSCABAC_ERR_OVERFLOW;
}

static int decode_ops_host_impl(uint32_t n_streams, const uint64_t* h_byte_off, const uint8_t* h_bytes,
                          const uint64_t* h_op_off, const void* h_ops, int op_width,
                          const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                          uint8_t* h_bins, uint8_t* h_finish_ok, bool packed) {
  if (!h_op_off || !h_byte_off) { set_error("cabac_decode_ops_host: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (op_width != 1 && op_width != 2) { set_error("op_width must be 1 or 2"); return ISSCABAC_ERR_INVALID; }
  if (n_streams == 0) return ISSCABAC_OK;
  Pipeline* pipe = current_pipe();
  if (!pipe) return ISSCABAC_ERR_CUDA;
  Pipeline& g_pipe = *pipe;
  int rc = g_pipe.init();
  if (rc) return rc;
  cudaStream_t s0 = g_pipe.s[0];
  const uint64_t obase = h_op_off[0], total = h_op_off[n_streams] - obase;
  const uint64_t bbase = h_byte_off[0], nbytes = h_byte_off[n_streams] - bbase;
  const size_t ctx_bytes = (size_t)n_ctx * (per_stream_init ? n_streams : 1);
  DevBuf d_ops, d_off, d_boff, d_bytes, d_ctx, d_bins, d_ok, d_packed;
  Drain drain{g_pipe};
  if ((rc = d_ops.alloc(total * op_width, s0)) || (rc = d_off.alloc((n_streams + 1ull) * 8, s0)) ||
      (rc = d_boff.alloc((n_streams + 1ull) * 8, s0)) || (rc = d_bytes.alloc(nbytes + 16, s0)) ||
      (rc = d_ctx.alloc(ctx_bytes, s0)) || (rc = d_bins.alloc(total, s0)) || (rc = d_ok.alloc(n_streams, s0)) ||
      (rc = d_packed.alloc(packed ? (total + 7) / 8 + 16 : 16, s0)))
    return rc;
  std::vector<uint64_t> off_rel, boff_rel;
  const uint64_t *off_src = h_op_off, *boff_src = h_byte_off;
  if (obase) {
    off_rel.resize(n_streams + 1ull);
    for (uint32_t s = 0; s <= n_streams; ++s) off_rel[s] = h_op_off[s] - obase;
    off_src = off_rel.data();
  }
  if (bbase) {
    boff_rel.resize(n_streams + 1ull);
    for (uint32_t s = 0; s <= n_streams; ++s) boff_rel[s] = h_byte_off[s] - bbase;
    boff_src = boff_rel.data();
  }
  CK(cudaMemcpyAsync(d_off.p, off_src, (n_streams + 1ull) * 8, cudaMemcpyHostToDevice, s0));
  CK(cudaMemcpyAsync(d_boff.p, boff_src, (n_streams + 1ull) * 8, cudaMemcpyHostToDevice, s0));
  if (nbytes) CK(cudaMemcpyAsync(d_bytes.p, h_bytes + bbase, nbytes, cudaMemcpyHostToDevice, s0));
  if (ctx_bytes) CK(cudaMemcpyAsync(d_ctx.p, h_ctx_init, ctx_bytes, cudaMemcpyHostToDevice, s0));
  CK(cudaEventRecord(g_pipe.done[0], s0));
  for (int l = 1; l < kLanes; ++l) CK(cudaStreamWaitEvent(g_pipe.s[l], g_pipe.done[0], 0));
  std::vector<uint32_t> cb;
  make_chunks(n_streams, off_src, kChunks, cb);
  bool chunks_byte_aligned = true;
  for (size_t k = 1; k + 1 < cb.size(); ++k) chunks_byte_aligned = chunks_byte_aligned && (off_src[cb[k]] & 7u) == 0;
  for (size_t k = 0; k + 1 < cb.size(); ++k) {
    cudaStream_t st = g_pipe.s[k % kLanes];
    const uint32_t a = cb[k], b = cb[k + 1];
    const uint64_t oa = off_src[a], ob = off_src[b];
    if (ob > oa)
      CK(cudaMemcpyAsync(d_ops.as<uint8_t>() + oa * op_width, (const uint8_t*)h_ops + (obase + oa) * op_width,
                         (ob - oa) * op_width, cudaMemcpyHostToDevice, st));
    rc = cabac_decode_ops(b - a, d_boff.as<uint64_t>() + a, d_bytes.as<uint8_t>(), d_off.as<uint64_t>() + a, d_ops.p,
                          op_width, d_ctx.as<uint8_t>() + (per_stream_init ? (size_t)a * n_ctx : 0), n_ctx,
                          per_stream_init, d_bins.as<uint8_t>(), d_ok.as<uint8_t>() + a, st);
    if (rc) return rc;
    if (h_bins && ob > oa && !packed)
      CK(cudaMemcpyAsync(h_bins + oa, d_bins.as<uint8_t>() + oa, ob - oa, cudaMemcpyDeviceToHost, st));
    // packed: a chunk whose op range starts and ends on byte boundaries of the packed array is packed and sent home
    // on its own lane while later chunks are still coded; otherwise everything is packed after the last chunk
    if (h_bins && ob > oa && packed && chunks_byte_aligned) {
      if ((rc = cabac_pack_bins(d_bins.as<uint8_t>(), oa, ob, d_packed.as<uint8_t>(), st))) return rc;
      CK(cudaMemcpyAsync(h_bins + (oa >> 3), d_packed.as<uint8_t>() + (oa >> 3), (ob - oa + 7) >> 3, cudaMemcpyDeviceToHost, st));
    }
  }
  for (int l = 1; l < kLanes; ++l) {
    CK(cudaEventRecord(g_pipe.done[l], g_pipe.s[l]));
    CK(cudaStreamWaitEvent(s0, g_pipe.done[l], 0));
  }
  if (h_bins && packed && !chunks_byte_aligned && total) {
    if ((rc = cabac_pack_bins(d_bins.as<uint8_t>(), 0, total, d_packed.as<uint8_t>(), s0))) return rc;
    CK(cudaMemcpyAsync(h_bins, d_packed.p, (total + 7) >> 3, cudaMemcpyDeviceToHost, s0));
  }
  if (h_finish_ok) CK(cudaMemcpyAsync(h_finish_ok, d_ok.p, n_streams, cudaMemcpyDeviceToHost, s0));
  CK(cudaStreamSynchronize(s0));
  return ISSCABAC_OK;
}

int cabac_decode_ops_host(uint32_t n_streams, const uint64_t* h_byte_off, const uint8_t* h_bytes,
                          const uint64_t* h_op_off, const void* h_ops, int op_width,
                          const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                          uint8_t* h_bins, uint8_t* h_finish_ok) {
  return decode_ops_host_impl(n_streams, h_byte_off, h_bytes, h_op_off, h_ops, op_width, h_ctx_init, n_ctx, per_stream_init,
                              h_bins, h_finish_ok, false);
}

int cabac_decode_ops_host_packed(uint32_t n_streams, const uint64_t* h_byte_off, const uint8_t* h_bytes,
                                 const uint64_t* h_op_off, const void* h_ops, int op_width,
                                 const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                                 uint8_t* h_bins_packed, uint8_t* h_finish_ok) {
  return decode_ops_host_impl(n_streams, h_byte_off, h_bytes, h_op_off, h_ops, op_width, h_ctx_init, n_ctx, per_stream_init,
                              h_bins_packed, h_finish_ok, true);
}


}  // extern "C"

// ---------------------------------------------------------------------------
// symbol-level host-buffer entry points (small inputs: ISS matrices, demo sequences)
// ---------------------------------------------------------------------------
extern "C" {

int cabac_encode_symbols_host(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* h_sym_off,
                              const void* h_symbols, int sym_width,
                              const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                              uint8_t* h_payload, uint64_t payload_cap, uint64_t* h_byte_off,
                              uint32_t* h_bits_after_symbol) {
  if (!cfg || !h_sym_off || !h_byte_off) { set_error("cabac_encode_symbols_host: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (sym_width != 1 && sym_width != 2 && sym_width != 4) { set_error("sym_width must be 1, 2 or 4"); return ISSCABAC_ERR_INVALID; }
  if (n_streams == 0) { h_byte_off[0] = 0; return ISSCABAC_OK; }
  if (h_sym_off[0] != 0) { set_error("sym_off[0] must be 0"); return ISSCABAC_ERR_INVALID; }
  Pipeline* pipe = current_pipe();
  if (!pipe) return ISSCABAC_ERR_CUDA;
  Pipeline& g_pipe = *pipe;
  int rc = g_pipe.init();
  if (rc) return rc;
  cudaStream_t s0 = g_pipe.s[0];
  const uint64_t n_sym = h_sym_off[n_streams];
  uint64_t max_sym = 0;
  for (uint32_t s = 0; s < n_streams; ++s) max_sym = std::max(max_sym, h_sym_off[s + 1] - h_sym_off[s]);
  const size_t ctx_bytes = (size_t)n_ctx * (per_stream_init ? n_streams : 1);
  DevBuf d_sym, d_off, d_ctx, d_len, d_boff, d_scr, d_flag, d_bits;
  Drain drain{g_pipe};
  if ((rc = d_sym.alloc(n_sym * sym_width, s0)) || (rc = d_off.alloc((n_streams + 1ull) * 8, s0)) ||
      (rc = d_ctx.alloc(ctx_bytes, s0)) || (rc = d_len.alloc(n_streams * 4ull, s0)) ||
      (rc = d_boff.alloc((n_streams + 1ull) * 8, s0)) ||
      (rc = d_scr.alloc(cabac_compact_scratch_bytes(n_streams), s0)) || (rc = d_flag.alloc(16, s0)) ||
      (rc = d_bits.alloc(h_bits_after_symbol ? n_sym * 4 : 16, s0)))
    return rc;
  CK(cudaMemcpyAsync(d_off.p, h_sym_off, (n_streams + 1ull) * 8, cudaMemcpyHostToDevice, s0));
  if (ctx_bytes) CK(cudaMemcpyAsync(d_ctx.p, h_ctx_init, ctx_bytes, cudaMemcpyHostToDevice, s0));
  // bins per symbol: EG-k of a b-bit value has at most 2b+3 bins, TU at most Nq-1
  uint64_t bins_per_sym = cfg->method == ISSCABAC_BIN_TU ? (cfg->Nq > 1 ? cfg->Nq - 1 : 1)
                        : cfg->method == ISSCABAC_BIN_FL32 ? 32 : (uint64_t)(2 * 8 * sym_width + 3);
  if (cfg->method >= ISSCABAC_BIN_TR0) {   // truncated Rice: (v >> k) + 1 + k bins, v < Nq (or whatever the symbol type holds)
    const uint32_t k = (uint32_t)(cfg->method - ISSCABAC_BIN_TR0);
    const uint64_t vmax = cfg->Nq ? cfg->Nq - 1 : (sym_width == 4 ? 0xfffffull : (1ull << (8 * sym_width)) - 1);
    bins_per_sym = (vmax >> k) + 1 + k;
  }
  // practical stride first (one byte per symbol: adaptive CABAC of quantised data stays far below), proof-level bound on retry
  const uint64_t stride_bound = cabac_slab_stride_bound(max_sym * bins_per_sym);
  uint64_t stride = std::min<uint64_t>(stride_bound, ((max_sym + 64) + 15) & ~15ull);
  std::vector<uint32_t> cb;
  // (the per-symbol bit trace runs on the one-kernel path: its output is indexed by the global symbol position)
  make_chunks(n_streams, h_sym_off, h_bits_after_symbol ? 1 : kChunks, cb);
  for (int attempt = 0; attempt < 2; ++attempt) {
    DevBuf d_slab, d_payload;
    Drain drain_inner{g_pipe};
    if ((rc = d_slab.alloc((size_t)n_streams * stride, s0))) return rc;
    CK(cudaMemsetAsync(d_flag.p, 0, 16, s0));
    CK(cudaEventRecord(g_pipe.done[0], s0));
    for (int l = 1; l < kLanes; ++l) CK(cudaStreamWaitEvent(g_pipe.s[l], g_pipe.done[0], 0));
    // stream groups over the lanes: the symbols of group k+1 travel while group k is coded
    for (size_t k = 0; k + 1 < cb.size(); ++k) {
      cudaStream_t st = g_pipe.s[k % kLanes];
      const uint32_t a = cb[k], b = cb[k + 1];
      const uint64_t sa = h_sym_off[a], sb = h_sym_off[b];
      if (attempt == 0 && sb > sa)
        CK(cudaMemcpyAsync(d_sym.as<uint8_t>() + sa * sym_width, (const uint8_t*)h_symbols + sa * sym_width, (sb - sa) * sym_width,
                           cudaMemcpyHostToDevice, st));
      // the chunk's offsets index the whole symbol buffer: the kernels get the buffer base and this chunk's slice of the table
      rc = cabac_encode_symbols(cfg, b - a, d_off.as<uint64_t>() + a, d_sym.p, sym_width,
                                d_ctx.as<uint8_t>() + (per_stream_init ? (size_t)a * n_ctx : 0), n_ctx, per_stream_init,
                                d_slab.as<uint8_t>() + (size_t)a * stride, stride, d_len.as<uint32_t>() + a,
                                h_bits_after_symbol ? d_bits.as<uint32_t>() : nullptr, d_flag.as<uint32_t>(), st);
      if (rc) return rc;
    }
    for (int l = 1; l < kLanes; ++l) {
      CK(cudaEventRecord(g_pipe.done[l], g_pipe.s[l]));
      CK(cudaStreamWaitEvent(s0, g_pipe.done[l], 0));
    }
    if ((rc = exclusive_scan_u32_u64(d_len.as<uint32_t>(), d_boff.as<uint64_t>(), n_streams, d_scr.p, s0))) return rc;
    uint32_t flag = 0;
    uint64_t total_bytes = 0;
    CK(cudaMemcpyAsync(&flag, d_flag.p, 4, cudaMemcpyDeviceToHost, s0));
    CK(cudaMemcpyAsync(&total_bytes, d_boff.as<uint64_t>() + n_streams, 8, cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    if (flag & 1u) {
      if (attempt == 1 || stride >= stride_bound) { set_error("slab overflow with the proof-level stride"); return ISSCABAC_ERR_OVERFLOW; }
      stride = stride_bound;
      continue;
    }
    if (total_bytes > payload_cap) {
      set_error("payload needs %llu bytes, capacity %llu", (unsigned long long)total_bytes, (unsigned long long)payload_cap);
      return ISSCABAC_ERR_OVERFLOW;
    }
    if ((rc = d_payload.alloc((total_bytes + 3) & ~(size_t)3, s0))) return rc;
    rc = cabac_compact(n_streams, d_slab.as<uint8_t>(), stride, d_len.as<uint32_t>(), d_payload.as<uint8_t>(), total_bytes,
                       d_boff.as<uint64_t>(), d_scr.p, d_flag.as<uint32_t>(), s0);
    if (rc) return rc;
    if (total_bytes) CK(cudaMemcpyAsync(h_payload, d_payload.p, total_bytes, cudaMemcpyDeviceToHost, s0));
    CK(cudaMemcpyAsync(h_byte_off, d_boff.p, (n_streams + 1ull) * 8, cudaMemcpyDeviceToHost, s0));
    if (h_bits_after_symbol && n_sym) CK(cudaMemcpyAsync(h_bits_after_symbol, d_bits.p, n_sym * 4, cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    return ISSCABAC_OK;
  }
  return ISSCABAC_ERR_OVERFLOW;
}

int cabac_decode_symbols_host(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* h_byte_off,
                              const uint8_t* h_bytes, const uint64_t* h_sym_off,
                              const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                              void* h_symbols, int sym_width, uint8_t* h_finish_ok) {
  if (!cfg || !h_sym_off || !h_byte_off) { set_error("cabac_decode_symbols_host: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (sym_width != 1 && sym_width != 2 && sym_width != 4) { set_error("sym_width must be 1, 2 or 4"); return ISSCABAC_ERR_INVALID; }
  if (n_streams == 0) return ISSCABAC_OK;
  if (h_sym_off[0] != 0 || h_byte_off[0] != 0) { set_error("offset tables must start at 0"); return ISSCABAC_ERR_INVALID; }
  Pipeline* pipe = current_pipe();
  if (!pipe) return ISSCABAC_ERR_CUDA;
  Pipeline& g_pipe = *pipe;
  int rc = g_pipe.init();
  if (rc) return rc;
  cudaStream_t s0 = g_pipe.s[0];
  const uint64_t n_sym = h_sym_off[n_streams], nbytes = h_byte_off[n_streams];
  const size_t ctx_bytes = (size_t)n_ctx * (per_stream_init ? n_streams : 1);
  DevBuf d_sym, d_off, d_boff, d_bytes, d_ctx, d_ok;
  Drain drain{g_pipe};
  if ((rc = d_sym.alloc(n_sym * sym_width, s0)) || (rc = d_off.alloc((n_streams + 1ull) * 8, s0)) ||
      (rc = d_boff.alloc((n_streams + 1ull) * 8, s0)) || (rc = d_bytes.alloc(nbytes + 16, s0)) ||
      (rc = d_ctx.alloc(ctx_bytes, s0)) || (rc = d_ok.alloc(n_streams, s0)))
    return rc;
  CK(cudaMemcpyAsync(d_off.p, h_sym_off, (n_streams + 1ull) * 8, cudaMemcpyHostToDevice, s0));
  CK(cudaMemcpyAsync(d_boff.p, h_byte_off, (n_streams + 1ull) * 8, cudaMemcpyHostToDevice, s0));
  if (ctx_bytes) CK(cudaMemcpyAsync(d_ctx.p, h_ctx_init, ctx_bytes, cudaMemcpyHostToDevice, s0));
  CK(cudaEventRecord(g_pipe.done[0], s0));
  for (int l = 1; l < kLanes; ++l) CK(cudaStreamWaitEvent(g_pipe.s[l], g_pipe.done[0], 0));
  // stream groups over the lanes: the bytes of group k+1 arrive and the symbols of group k-1 leave while group k is decoded
  std::vector<uint32_t> cb;
  make_chunks(n_streams, h_sym_off, kChunks, cb);
  for (size_t k = 0; k + 1 < cb.size(); ++k) {
    cudaStream_t st = g_pipe.s[k % kLanes];
    const uint32_t a = cb[k], b = cb[k + 1];
    const uint64_t ba = h_byte_off[a], bb = h_byte_off[b], sa = h_sym_off[a], sb = h_sym_off[b];
    if (bb > ba) CK(cudaMemcpyAsync(d_bytes.as<uint8_t>() + ba, h_bytes + ba, bb - ba, cudaMemcpyHostToDevice, st));
    rc = cabac_decode_symbols(cfg, b - a, d_boff.as<uint64_t>() + a, d_bytes.as<uint8_t>(), d_off.as<uint64_t>() + a,
                              d_ctx.as<uint8_t>() + (per_stream_init ? (size_t)a * n_ctx : 0), n_ctx, per_stream_init, d_sym.p,
                              sym_width, d_ok.as<uint8_t>() + a, st);
    if (rc) return rc;
    if (h_symbols && sb > sa)
      CK(cudaMemcpyAsync((uint8_t*)h_symbols + sa * sym_width, d_sym.as<uint8_t>() + sa * sym_width, (sb - sa) * sym_width,
                         cudaMemcpyDeviceToHost, st));
  }
  for (int l = 1; l < kLanes; ++l) {
    CK(cudaEventRecord(g_pipe.done[l], g_pipe.s[l]));
    CK(cudaStreamWaitEvent(s0, g_pipe.done[l], 0));
  }
  if (h_finish_ok) CK(cudaMemcpyAsync(h_finish_ok, d_ok.p, n_streams, cudaMemcpyDeviceToHost, s0));
  CK(cudaStreamSynchronize(s0));
  return ISSCABAC_OK;
}

}  // extern "C"
