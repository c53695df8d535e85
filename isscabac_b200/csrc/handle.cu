// handle.cu -- single-stream engine handle: the reference's `class CABAC` aggregate
// (CABAC/SimpleCABACMex.cpp:69-80: bitstream + encoder context set + decoder context set
// + encoder + decoder) with the coder state resident on the GPU.
//
// * encode: engine calls (encodeBin / encodeBinEP / encodeBinsEP / encodeBinTrm) are queued
//   as u16 ops on the host and executed by k_stream_encode -- which RESUMES from the saved
//   device state -- whenever the caller needs an answer (getNumBits) or finishes the stream.
//   The per-bin call therefore costs a vector push, and a whole stream costs one launch.
// * decode: decodeBin(ctx) must return a value the caller uses to choose the next context
//   (cabacDecode.m:38-45), so each call is one tiny launch of k_stream_decode with the bin
//   returned through mapped pinned memory.  This is the conformance path; throughput lives
//   in the batch entry points (kernels.cu / symbols.cu).
// * like the reference, context models persist across encodeStart()/decodeStart() calls and
//   are only (re)set by initByProb/initByState (SimpleCABACMex.cpp:186-209 does not touch
//   them); indices up to 999 are usable even beyond the initialised count, with the
//   constructor default (mps=1,state=0) (ContextModel.cpp:51-55).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/isscabac.h"
#include "cabac_lane.cuh"
#include "internal.h"

using namespace cabac;
using namespace isscabac_internal;

namespace {

constexpr uint32_t kMaxCtx = 1000;  // RWTH_MAX_NUM_CONTEXTS, CommonDef.h:56

struct DevState {
  // encoder registers (EncLane without its pointers)
  uint32_t low, range;
  int32_t bits_left;
  uint32_t acc, nbytes, nbuf, overflow;
  uint32_t enc_bins;
  // decoder registers
  uint32_t value, drange;
  int32_t bits_needed;
  uint32_t pos, last, finish_ok;
  uint8_t enc_ctx[kMaxCtx];
  uint8_t dec_ctx[kMaxCtx];
};

struct RowTable {
  uint2 r[128];
  constexpr RowTable() : r{} {
    for (uint32_t i = 0; i < 128; ++i) r[i] = fused_row(i);
  }
};
__constant__ RowTable c_rows_h = RowTable();

__global__ void k_stream_encode(DevState* S, const uint16_t* ops, uint32_t n, uint8_t* out, uint32_t cap,
                                int do_start, int do_finish) {
  if (threadIdx.x || blockIdx.x) return;
  EncLane L;
  if (do_start) {
    enc_start(L, out, cap);
    S->enc_bins = 0;
  } else {
    L.low = S->low; L.range = S->range; L.bits_left = S->bits_left;
    L.acc = S->acc; L.nbytes = S->nbytes; L.nbuf = S->nbuf; L.overflow = S->overflow;
    L.cap = cap; L.out = out;
  }
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t o = ops[i], code = o >> 1, bin = o & 1u;
    if (code == ISSCABAC_OP16_EP) enc_bin_ep<true>(L, bin);
    else if (code == ISSCABAC_OP16_TRM) enc_bin_trm<true>(L, bin);
    else if (code < kMaxCtx) {
      uint32_t st = S->enc_ctx[code];
      enc_bin_ctx<true>(L, bin, st, c_rows_h.r[st & 127u]);
      S->enc_ctx[code] = (uint8_t)st;
    }
  }
  S->enc_bins += n;
  if (do_finish) { enc_finish<true>(L); S->enc_bins += 1; }
  enc_flush_pending(L);
  S->low = L.low; S->range = L.range; S->bits_left = L.bits_left;
  S->acc = L.acc; S->nbytes = L.nbytes; S->nbuf = L.nbuf; S->overflow = L.overflow;
}

// mode 0: start (reads two bytes), 1: decode n ops, 2: finish check
__global__ void k_stream_decode(DevState* S, int mode, const uint16_t* ops, uint32_t n, const uint8_t* in,
                                uint32_t len, uint8_t* bins) {
  if (threadIdx.x || blockIdx.x) return;
  DecLane D;
  if (mode == 0) {
    dec_start(D, in, len);
  } else {
    D.value = S->value; D.range = S->drange; D.bits_needed = S->bits_needed;
    D.pos = S->pos; D.len = len; D.last = S->last; D.in = in;
  }
  if (mode == 1) {
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t code = ops[i] >> 1;
      uint32_t bin = 0;
      if (code == ISSCABAC_OP16_EP) bin = dec_bin_ep(D);
      else if (code == ISSCABAC_OP16_TRM) bin = dec_bin_trm(D);
      else if (code < kMaxCtx) {
        uint32_t st = S->dec_ctx[code];
        bin = dec_bin_ctx(D, st, c_rows_h.r[st & 127u]);
        S->dec_ctx[code] = (uint8_t)st;
      }
      bins[i] = (uint8_t)bin;
    }
  } else if (mode == 2) {
    S->finish_ok = dec_finish(D);
  }
  S->value = D.value; S->drange = D.range; S->bits_needed = D.bits_needed;
  S->pos = D.pos; S->last = D.last;
}

}  // namespace

struct simplecabac {
  std::string fn;
  bool has_fn = false;
  cudaStream_t st = nullptr;
  DevState* d_state = nullptr;
  // encode side
  bool encoding = false, enc_started_on_device = false;
  std::vector<uint16_t> pending;
  uint16_t* d_ops = nullptr; size_t d_ops_cap = 0;
  uint8_t* d_out = nullptr; size_t d_out_cap = 0;
  uint64_t bytes_ub = 0;          // upper bound of bytes produced so far in this stream
  uint64_t bits_base = 0;         // bits of all earlier streams on this handle (the reference never resets its counter)
  uint64_t cur_bits = 0, total_bins = 0;
  uint32_t cur_nbytes = 0;
  std::vector<uint8_t> out_bytes; // finished stream
  // decode side
  bool decoding = false;
  std::vector<uint8_t> in_bytes; bool in_set = false;
  uint8_t* d_in = nullptr; size_t d_in_cap = 0; uint32_t in_len = 0;
  uint8_t* h_bins = nullptr; uint8_t* d_bins_mapped = nullptr; size_t bins_cap = 0;
  // trace (RWTH_TRACE_CABAC_STATES): the (bin, ctx) history since initBy*, per context set
  bool trace = false;
  std::vector<uint16_t> enc_log, dec_log;
  std::vector<uint8_t> init_bytes = std::vector<uint8_t>(kMaxCtx, (uint8_t)1);
};

namespace {

int ensure_dev(void** p, size_t* cap, size_t need, cudaStream_t st, bool keep) {
  if (need <= *cap) return ISSCABAC_OK;
  size_t ncap = need * 2 + 256;
  void* q = nullptr;
  CK(cudaMalloc(&q, ncap));
  if (*p) {
    if (keep) CK(cudaMemcpyAsync(q, *p, *cap, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(*p));
  }
  *p = q; *cap = ncap;
  return ISSCABAC_OK;
}

int ensure_bins(simplecabac* h, size_t n) {
  if (n <= h->bins_cap) return ISSCABAC_OK;
  if (h->h_bins) CK(cudaFreeHost(h->h_bins));
  h->bins_cap = n * 2 + 64;
  CK(cudaHostAlloc((void**)&h->h_bins, h->bins_cap, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&h->d_bins_mapped, h->h_bins, 0));
  return ISSCABAC_OK;
}

// run the queued ops (and optionally finish) on the device, refresh the host view
int enc_flush(simplecabac* h, bool finish) {
  const uint32_t n = (uint32_t)h->pending.size();
  int rc;
  if ((rc = ensure_dev((void**)&h->d_ops, &h->d_ops_cap, (size_t)(n + 1) * 2, h->st, false))) return rc;
  h->bytes_ub += n + 4;
  if ((rc = ensure_dev((void**)&h->d_out, &h->d_out_cap, ((size_t)h->bytes_ub + 19) & ~(size_t)15, h->st, true))) return rc;
  if (n) CK(cudaMemcpyAsync(h->d_ops, h->pending.data(), (size_t)n * 2, cudaMemcpyHostToDevice, h->st));
  uint32_t cap = (uint32_t)(h->d_out_cap & ~(size_t)3);
  k_stream_encode<<<1, 32, 0, h->st>>>(h->d_state, h->d_ops, n, h->d_out, cap, h->enc_started_on_device ? 0 : 1, finish ? 1 : 0);
  CK(cudaGetLastError());
  h->enc_started_on_device = true;
  h->pending.clear();
  uint32_t hdr[8];
  CK(cudaMemcpyAsync(hdr, h->d_state, sizeof hdr, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  const uint32_t nbytes = hdr[4], nbuf = hdr[5], ovf = hdr[6];
  if (ovf) { set_error("internal: stream output buffer overflow"); return ISSCABAC_ERR_OVERFLOW; }
  h->cur_nbytes = nbytes;
  h->cur_bits = finish ? 8ull * nbytes : 8ull * (nbytes - nbuf);
  h->total_bins = hdr[7];
  return ISSCABAC_OK;
}

int init_ctx(simplecabac* h, const uint8_t* ctx, uint32_t n) {
  // both sets, like initByProb/initByState (SimpleCABACMex.cpp:165-176); everything beyond n
  // gets the ContextModel constructor default (mps=1, state=0)
  std::vector<uint8_t> full(kMaxCtx, (uint8_t)1);
  if (n > ISSCABAC_MAX_CTX) { set_error("at most %u contexts", ISSCABAC_MAX_CTX); return ISSCABAC_ERR_INVALID; }
  memcpy(full.data(), ctx, n);
  h->init_bytes = full;       // a fresh `class CABAC` in the reference: its trace members start empty
  h->enc_log.clear();
  h->dec_log.clear();
  CK(cudaMemcpyAsync(h->d_state->enc_ctx, full.data(), kMaxCtx, cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(h->d_state->dec_ctx, full.data(), kMaxCtx, cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));
  return ISSCABAC_OK;
}

}  // namespace

extern "C" {

int simplecabac_create(simplecabac** out, const char* filename) {
  if (!out) return ISSCABAC_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return cuda_fail(e != cudaSuccess ? e : cudaErrorNoDevice, "simplecabac_create: no CUDA device");
  simplecabac* h = new simplecabac;
  if (filename) { h->fn = filename; h->has_fn = true; }
  CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  CK(cudaMalloc((void**)&h->d_state, sizeof(DevState)));
  CK(cudaMemsetAsync(h->d_state, 0, sizeof(DevState), h->st));
  uint8_t def = 1;
  int rc = init_ctx(h, &def, 0);
  if (rc) return rc;
  *out = h;
  return ISSCABAC_OK;
}

int simplecabac_destroy(simplecabac* h) {
  if (!h) return ISSCABAC_OK;
  cudaStreamSynchronize(h->st);
  cudaFree(h->d_state); cudaFree(h->d_ops); cudaFree(h->d_out); cudaFree(h->d_in);
  if (h->h_bins) cudaFreeHost(h->h_bins);
  cudaStreamDestroy(h->st);
  delete h;
  return ISSCABAC_OK;
}

int simplecabac_init_by_prob(simplecabac* h, const double* p0, uint32_t n) {
  if (!h) return ISSCABAC_ERR_INVALID;
  std::vector<uint8_t> ctx(n ? n : 1);
  int rc = cabac_ctx_from_prob(p0, n, ctx.data());
  return rc ? rc : init_ctx(h, ctx.data(), n);
}

int simplecabac_init_by_state(simplecabac* h, const double* triples, uint32_t n) {
  if (!h) return ISSCABAC_ERR_INVALID;
  std::vector<uint8_t> ctx(n ? n : 1);
  int rc = cabac_ctx_from_state(triples, n, ctx.data());
  return rc ? rc : init_ctx(h, ctx.data(), n);
}

// encodeStart (SimpleCABACMex.cpp:186-209): opens the sink, Encoder::start(); contexts untouched
int simplecabac_encode_start(simplecabac* h) {
  if (!h) return ISSCABAC_ERR_INVALID;
  if (h->has_fn) {
    FILE* f = fopen(h->fn.c_str(), "wb");
    if (!f) { set_error("filename %s cannot be opened for writing", h->fn.c_str()); return ISSCABAC_ERR_IO; }
    fclose(f);
  }
  h->encoding = true;
  h->enc_started_on_device = false;
  h->pending.clear();
  h->bytes_ub = 0; h->cur_bits = 0; h->cur_nbytes = 0; h->total_bins = 0;
  h->out_bytes.clear();
  return ISSCABAC_OK;
}

static int queue_op(simplecabac* h, uint32_t op) {
  if (!h) return ISSCABAC_ERR_INVALID;
  if (!h->encoding) { set_error("encode call outside encodeStart()...encodeFinish()"); return ISSCABAC_ERR_STATE; }
  h->pending.push_back((uint16_t)op);
  if (h->trace) h->enc_log.push_back((uint16_t)op);
  if (h->pending.size() >= (1u << 20)) return enc_flush(h, false);
  return ISSCABAC_OK;
}

int simplecabac_encode_bin(simplecabac* h, unsigned bin, unsigned ctx_idx) {
  if (bin > 1) { set_error("bin must be 0 or 1"); return ISSCABAC_ERR_INVALID; }
  if (ctx_idx > ISSCABAC_MAX_CTX) { set_error("context index %u > %u", ctx_idx, ISSCABAC_MAX_CTX); return ISSCABAC_ERR_INVALID; }
  return queue_op(h, (ctx_idx << 1) | bin);
}
int simplecabac_encode_bin_ep(simplecabac* h, unsigned bin) {
  if (bin > 1) { set_error("bin must be 0 or 1"); return ISSCABAC_ERR_INVALID; }
  return queue_op(h, (ISSCABAC_OP16_EP << 1) | bin);
}
// encodeBinsEP(values, n) is byte-identical to n single bypass bins, MSB first
// (Encoder.cpp:278-319; SURVEY.md row a7, K3 == K4), so it is queued as such.
int simplecabac_encode_bins_ep(simplecabac* h, unsigned bins, int n) {
  if (n < 0 || n > 32) { set_error("encodeBinsEP: 0 <= numBins <= 32"); return ISSCABAC_ERR_INVALID; }
  for (int i = n - 1; i >= 0; --i) {
    int rc = queue_op(h, (ISSCABAC_OP16_EP << 1) | ((bins >> i) & 1u));
    if (rc) return rc;
  }
  return ISSCABAC_OK;
}
int simplecabac_encode_bin_trm(simplecabac* h, unsigned bin) {
  if (bin > 1) { set_error("bin must be 0 or 1"); return ISSCABAC_ERR_INVALID; }
  return queue_op(h, (ISSCABAC_OP16_TRM << 1) | bin);
}

// getNumBits (SimpleCABACMex.cpp:248-264 -> CABAC_BitstreamFile.h:70): bits already handed
// to the sink; lags the coder by the buffered byte(s); never reset between streams.
int simplecabac_get_num_bits(simplecabac* h, uint64_t* bits) {
  if (!h || !bits) return ISSCABAC_ERR_INVALID;
  if (h->encoding && (!h->pending.empty() || !h->enc_started_on_device)) {
    int rc = enc_flush(h, false);
    if (rc) return rc;
  }
  *bits = h->bits_base + h->cur_bits;
  return ISSCABAC_OK;
}

int simplecabac_get_bins_coded(simplecabac* h, uint64_t* bins) {
  if (!h || !bins) return ISSCABAC_ERR_INVALID;
  if (h->encoding && !h->pending.empty()) {
    int rc = enc_flush(h, false);
    if (rc) return rc;
  }
  *bins = h->total_bins;
  return ISSCABAC_OK;
}

int simplecabac_encode_finish(simplecabac* h) {
  if (!h) return ISSCABAC_ERR_INVALID;
  if (!h->encoding) { set_error("encodeFinish without encodeStart"); return ISSCABAC_ERR_STATE; }
  int rc = enc_flush(h, true);
  if (rc) return rc;
  h->out_bytes.resize(h->cur_nbytes);
  if (h->cur_nbytes) CK(cudaMemcpyAsync(h->out_bytes.data(), h->d_out, h->cur_nbytes, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  h->encoding = false;
  h->bits_base += h->cur_bits;
  h->cur_bits = 0;
  if (h->has_fn) {
    FILE* f = fopen(h->fn.c_str(), "wb");
    if (!f) { set_error("filename %s cannot be opened for writing", h->fn.c_str()); return ISSCABAC_ERR_IO; }
    size_t w = h->out_bytes.empty() ? 0 : fwrite(h->out_bytes.data(), 1, h->out_bytes.size(), f);
    fclose(f);
    if (w != h->out_bytes.size()) { set_error("short write to %s", h->fn.c_str()); return ISSCABAC_ERR_IO; }
  }
  return ISSCABAC_OK;
}

int simplecabac_get_bytes(simplecabac* h, const uint8_t** bytes, uint64_t* n) {
  if (!h || !bytes || !n) return ISSCABAC_ERR_INVALID;
  *bytes = h->out_bytes.data();
  *n = h->out_bytes.size();
  return ISSCABAC_OK;
}

int simplecabac_set_bytes(simplecabac* h, const uint8_t* bytes, uint64_t n) {
  if (!h || (n && !bytes)) return ISSCABAC_ERR_INVALID;
  h->in_bytes.assign(bytes, bytes + n);
  h->in_set = true;
  return ISSCABAC_OK;
}

// decodeStart (SimpleCABACMex.cpp:280-300): opens the source, Decoder::start() (reads 2 bytes)
int simplecabac_decode_start(simplecabac* h) {
  if (!h) return ISSCABAC_ERR_INVALID;
  if (h->has_fn && !h->in_set) {
    FILE* f = fopen(h->fn.c_str(), "rb");
    if (!f) { set_error("filename %s cannot be opened for reading", h->fn.c_str()); return ISSCABAC_ERR_IO; }
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    h->in_bytes.resize(len > 0 ? (size_t)len : 0);
    size_t r = len > 0 ? fread(h->in_bytes.data(), 1, (size_t)len, f) : 0;
    fclose(f);
    if (r != h->in_bytes.size()) { set_error("short read from %s", h->fn.c_str()); return ISSCABAC_ERR_IO; }
  } else if (!h->in_set && !h->has_fn) {
    h->in_bytes = h->out_bytes;  // memory sink -> memory source hand-off
  }
  h->in_set = false;
  h->in_len = (uint32_t)h->in_bytes.size();
  int rc;
  if ((rc = ensure_dev((void**)&h->d_in, &h->d_in_cap, h->in_bytes.size() + 16, h->st, false))) return rc;
  if (h->in_len) CK(cudaMemcpyAsync(h->d_in, h->in_bytes.data(), h->in_len, cudaMemcpyHostToDevice, h->st));
  k_stream_decode<<<1, 32, 0, h->st>>>(h->d_state, 0, nullptr, 0, h->d_in, h->in_len, nullptr);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->st));
  h->decoding = true;
  return ISSCABAC_OK;
}

int simplecabac_decode_ops(simplecabac* h, const uint16_t* ops, uint32_t n, uint8_t* bins) {
  if (!h || (n && (!ops || !bins))) return ISSCABAC_ERR_INVALID;
  if (!h->decoding) { set_error("decode call outside decodeStart()...decodeFinish()"); return ISSCABAC_ERR_STATE; }
  if (n == 0) return ISSCABAC_OK;
  int rc;
  if ((rc = ensure_dev((void**)&h->d_ops, &h->d_ops_cap, (size_t)n * 2, h->st, false))) return rc;
  if ((rc = ensure_bins(h, n))) return rc;
  CK(cudaMemcpyAsync(h->d_ops, ops, (size_t)n * 2, cudaMemcpyHostToDevice, h->st));
  k_stream_decode<<<1, 32, 0, h->st>>>(h->d_state, 1, h->d_ops, n, h->d_in, h->in_len, h->d_bins_mapped);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->st));
  memcpy(bins, h->h_bins, n);
  if (h->trace)
    for (uint32_t i = 0; i < n; ++i) h->dec_log.push_back((uint16_t)((ops[i] & ~1u) | (bins[i] & 1u)));
  return ISSCABAC_OK;
}

static int decode_one_op(simplecabac* h, uint32_t code, unsigned* bin) {
  if (!bin) return ISSCABAC_ERR_INVALID;
  uint16_t op = (uint16_t)(code << 1);
  uint8_t b = 0;
  int rc = simplecabac_decode_ops(h, &op, 1, &b);
  *bin = b;
  return rc;
}

int simplecabac_decode_bin(simplecabac* h, unsigned ctx_idx, unsigned* bin) {
  if (ctx_idx > ISSCABAC_MAX_CTX) { set_error("context index %u > %u", ctx_idx, ISSCABAC_MAX_CTX); return ISSCABAC_ERR_INVALID; }
  return decode_one_op(h, ctx_idx, bin);
}
int simplecabac_decode_bin_ep(simplecabac* h, unsigned* bin) { return decode_one_op(h, ISSCABAC_OP16_EP, bin); }
int simplecabac_decode_bin_trm(simplecabac* h, unsigned* bin) { return decode_one_op(h, ISSCABAC_OP16_TRM, bin); }

// decodeBinsEP(n) == n single bypass decodes, MSB first (Decoder.cpp:333-421)
int simplecabac_decode_bins_ep(simplecabac* h, int n, unsigned* bins) {
  if (!bins || n < 0 || n > 32) { set_error("decodeBinsEP: 0 <= numBins <= 32"); return ISSCABAC_ERR_INVALID; }
  uint16_t ops[32];
  uint8_t out[32];
  for (int i = 0; i < n; ++i) ops[i] = (uint16_t)(ISSCABAC_OP16_EP << 1);
  int rc = simplecabac_decode_ops(h, ops, (uint32_t)n, out);
  if (rc) return rc;
  unsigned v = 0;
  for (int i = 0; i < n; ++i) v = (v << 1) | out[i];
  *bins = v;
  return ISSCABAC_OK;
}

// decodeFinish: the two asserts of Decoder::finish() (Decoder.cpp:75-81) as an error code
int simplecabac_decode_finish(simplecabac* h) {
  if (!h) return ISSCABAC_ERR_INVALID;
  if (!h->decoding) { set_error("decodeFinish without decodeStart"); return ISSCABAC_ERR_STATE; }
  k_stream_decode<<<1, 32, 0, h->st>>>(h->d_state, 2, nullptr, 0, h->d_in, h->in_len, nullptr);
  CK(cudaGetLastError());
  uint32_t ok = 0;
  CK(cudaMemcpyAsync(&ok, &h->d_state->finish_ok, 4, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  h->decoding = false;
  if (!ok) { set_error("terminate bin / stop bit check failed (Decoder::finish asserts)"); return ISSCABAC_ERR_CORRUPT; }
  return ISSCABAC_OK;
}

int simplecabac_set_trace(simplecabac* h, int on) {
  if (!h) return ISSCABAC_ERR_INVALID;
  h->trace = on != 0;
  return ISSCABAC_OK;
}

// getEncoderStats / getDecoderStats (SimpleCABACMex.cpp:356-466) for one context.  The state
// sequence of a context depends only on the bins coded with it, so the history is filtered to
// that context on the host and walked by the device trace kernel (stats.cu) as a one-context,
// one-stream job.
int simplecabac_get_stats(simplecabac* h, int decoder_set, unsigned ctx_idx, uint8_t* steps5, uint64_t cap_steps,
                          uint64_t* n_steps, uint32_t* trans) {
  if (!h || ctx_idx >= kMaxCtx) { set_error("simplecabac_get_stats: bad handle or context index"); return ISSCABAC_ERR_INVALID; }
  if (!h->trace) { set_error("tracing is off for this handle (simplecabac_set_trace)"); return ISSCABAC_ERR_STATE; }
  const std::vector<uint16_t>& log = decoder_set ? h->dec_log : h->enc_log;
  std::vector<uint8_t> ops;   // u8 op format, context 0
  for (uint16_t o : log)
    if ((unsigned)(o >> 1) == ctx_idx) ops.push_back((uint8_t)(o & 1u));
  const uint64_t n = ops.size();
  if (n_steps) *n_steps = n;
  if (trans) memset(trans, 0, 128 * 128 * 4);
  if (n == 0) return ISSCABAC_OK;
  uint8_t* d = nullptr;   // [ops n | pad | off 16 | ctx 8 | hist 1024 | trans 65536 | steps 2n]
  const size_t o_off = (n + 7) & ~(size_t)7, o_ctx = o_off + 16, o_hist = o_ctx + 8, o_trans = o_hist + 1024,
               o_steps = o_trans + 65536, total = o_steps + 2 * n;
  CK(cudaMalloc((void**)&d, total));
  const uint64_t off[2] = {0, n};
  const uint8_t ctx0 = h->init_bytes[ctx_idx];
  cudaError_t e = cudaMemcpyAsync(d, ops.data(), n, cudaMemcpyHostToDevice, h->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + o_off, off, 16, cudaMemcpyHostToDevice, h->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + o_ctx, &ctx0, 1, cudaMemcpyHostToDevice, h->st);
  int rc = e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "cudaMemcpyAsync(trace)");
  if (!rc)
    rc = cabac_ctx_trace_ops(1, reinterpret_cast<uint64_t*>(d + o_off), d, 1, d + o_ctx, 1, 0, 1,
                             reinterpret_cast<uint64_t*>(d + o_hist), reinterpret_cast<uint32_t*>(d + o_trans), nullptr,
                             d + o_steps, nullptr, h->st);
  std::vector<uint8_t> st2(2 * n);
  if (!rc) {
    e = cudaMemcpyAsync(st2.data(), d + o_steps, 2 * n, cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess && trans) e = cudaMemcpyAsync(trans, d + o_trans, 65536, cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e != cudaSuccess) rc = cuda_fail(e, "trace read-back");
  }
  cudaFree(d);
  if (rc) return rc;
  auto ts = [](uint32_t b) { return (b & 1u) ? (b >> 1) + 64u : 63u - (b >> 1); };   // ContextModel.cpp:128-129
  for (uint64_t i = 0; i < n && i < cap_steps && steps5; ++i) {
    const uint32_t p = st2[2 * i], a = st2[2 * i + 1];
    steps5[5 * i + 0] = decoder_set ? 0 : ops[i];   // the reference decoder logs the bin before decoding it
    steps5[5 * i + 1] = (uint8_t)ts(p);
    steps5[5 * i + 2] = (uint8_t)(p & 1u);
    steps5[5 * i + 3] = (uint8_t)ts(a);
    steps5[5 * i + 4] = (uint8_t)(a & 1u);
  }
  return ISSCABAC_OK;
}

int simplecabac_get_ctx_state(simplecabac* h, int decoder_set, unsigned ctx_idx, unsigned* state, unsigned* mps) {
  if (!h || ctx_idx >= kMaxCtx) return ISSCABAC_ERR_INVALID;
  if (h->encoding && !h->pending.empty()) {
    int rc = enc_flush(h, false);
    if (rc) return rc;
  }
  uint8_t b = 0;
  CK(cudaMemcpyAsync(&b, (decoder_set ? h->d_state->dec_ctx : h->d_state->enc_ctx) + ctx_idx, 1, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (state) *state = b >> 1;
  if (mps) *mps = b & 1u;
  return ISSCABAC_OK;
}

}  // extern "C"
