// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Batch driver around the UNMODIFIED reference engine.  This file contains no
// reference code: it includes the reference headers from where they lie
// (/root/reference/CABAC, given with -I at build time, see oracle/Makefile) and
// is linked against the reference's own translation units
//   CABAC_ArithmeticEncoder.cpp CABAC_ArithmeticDecoder.cpp
//   CABAC_BitstreamFile.cpp     ContextModel.cpp
// compiled in place.  The result (oracle/_ref/libref_cabac.so) is
//   * the byte-exactness oracle the C restatement (cabac_oracle.c) is pinned to,
//   * the generator of tests/golden/*.json (oracle/gen_golden.py),
//   * the timed CPU baseline of bench.py (`cpu_baseline.kind == "reference"`).
//
// The reference engine only talks to std::fstream (CABAC_BitstreamFile.cpp:50-65),
// so each worker thread owns one scratch file under `tmpdir` (use a tmpfs such
// as /dev/shm), exactly like the reference's encoder->file->decoder hand-off
// (SimpleCABACMex.cpp:195,288).
//
// Op format (shared with the product, see include/isscabac.h):
//   op = (code << 1) | bin ; u8 ops: code 0..124 = context index, 125 = terminate
//   bin, 126 = bypass (EP) bin; u16 ops: code 0..998 context, 0x7FFD TRM, 0x7FFE EP.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <unistd.h>

#include "CABAC_ArithmeticEncoder.h"
#include "CABAC_ArithmeticDecoder.h"
#include "CABAC_BitstreamFile.h"
#include "ContextModel.h"

namespace {

struct OpView {
  const uint8_t* p8;
  const uint16_t* p16;
  uint32_t ep, trm;
  OpView(const void* ops, int width)
      : p8(width == 1 ? (const uint8_t*)ops : nullptr),
        p16(width == 2 ? (const uint16_t*)ops : nullptr),
        ep(width == 1 ? 126u : 0x7FFEu),
        trm(width == 1 ? 125u : 0x7FFDu) {}
  inline uint32_t at(uint64_t i) const { return p8 ? p8[i] : p16[i]; }
};

// The reference keeps encodeBinTrm protected (CABAC_ArithmeticEncoder.h:71-72);
// a derived class re-exports it without touching the reference source.
struct EncoderWithTrm : public CABAC_ArithmeticEncoder {
  void trm(unsigned int b) { encodeBinTrm(b); }
};
// Same idea for the decoder's protected bit counter, needed to evaluate the
// stop-bit check of Decoder::finish() (CABAC_ArithmeticDecoder.cpp:81) without assert().
struct DecoderProbe : public CABAC_ArithmeticDecoder {
  int bitsNeeded() const { return m_bitsNeeded; }
};

std::string scratch_name(const char* tmpdir, int tid) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s/refcabac_%d_%d.bin", tmpdir, (int)getpid(), tid);
  return buf;
}

void init_ctx(std::vector<ContextModel>& ctx, const uint8_t* init, uint32_t n_ctx) {
  for (uint32_t c = 0; c < n_ctx; ++c) ctx[c].init(init[c] & 1u, init[c] >> 1);
}

}  // namespace

extern "C" {

// Encode n_streams independent streams: start(); ops...; finish().
// ctx_init: n_ctx bytes ((state<<1)|mps) shared by all streams, or
// n_streams*n_ctx when per_stream_init != 0.  Output: stream s is written to
// out + s*out_stride (at most out_stride bytes are copied), its true length to
// out_len[s].  Returns 0, or -1 on I/O failure.
int ref_encode_ops(uint32_t n_streams, const uint64_t* op_off, const void* ops, int op_width,
                   const uint8_t* ctx_init, uint32_t n_ctx, int per_stream_init,
                   uint8_t* out, uint64_t out_stride, uint32_t* out_len,
                   int n_threads, const char* tmpdir) {
  if (n_threads < 1) n_threads = 1;
  OpView ov(ops, op_width);
  // all ContextModel objects are constructed on this thread (the class keeps a
  // non-atomic static instance counter, ContextModel.cpp:45,56-57)
  std::vector<std::vector<ContextModel>> ctxs(n_threads);
  for (auto& v : ctxs) v.resize(n_ctx ? n_ctx : 1);
  std::atomic<uint32_t> next(0);
  std::atomic<int> err(0);
  auto work = [&](int tid) {
    std::string fn = scratch_name(tmpdir, tid);
    std::vector<ContextModel>& ctx = ctxs[tid];
    for (;;) {
      uint32_t s = next.fetch_add(1);
      if (s >= n_streams) break;
      init_ctx(ctx, ctx_init + (per_stream_init ? (uint64_t)s * n_ctx : 0), n_ctx);
      CABAC_BitstreamFile bs;
      if (!bs.openOutputFile(fn.c_str())) { err = -1; break; }
      EncoderWithTrm enc;
      enc.setBitstream(&bs);
      enc.start();
      for (uint64_t i = op_off[s]; i < op_off[s + 1]; ++i) {
        uint32_t o = ov.at(i), code = o >> 1, bin = o & 1u;
        if (code == ov.ep) enc.encodeBinEP(bin);
        else if (code == ov.trm) enc.trm(bin);
        else enc.encodeBin(bin, &ctx[code]);
      }
      enc.finish();
      bs.closeFile();
      FILE* f = fopen(fn.c_str(), "rb");
      if (!f) { err = -1; break; }
      fseek(f, 0, SEEK_END);
      long len = ftell(f);
      fseek(f, 0, SEEK_SET);
      out_len[s] = (uint32_t)len;
      uint64_t ncopy = (uint64_t)len < out_stride ? (uint64_t)len : out_stride;
      if (ncopy && fread(out + (uint64_t)s * out_stride, 1, ncopy, f) != ncopy) err = -1;
      fclose(f);
    }
    unlink(fn.c_str());
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  return err.load();
}

// Decode: stream s = bytes[byte_off[s] .. byte_off[s+1]); the op array gives the
// kind of every bin (bit 0 of each op is ignored).  out_bins[i] = decoded bin of
// op i.  finish_ok[s] (optional) = 1 when the terminate bin decodes to 1 and the
// stop-bit pattern matches (the two asserts of CABAC_ArithmeticDecoder.cpp:75-81,
// evaluated here without aborting; Decoder::finish() itself is not called so a
// corrupt test stream cannot abort the process).
int ref_decode_ops(uint32_t n_streams, const uint64_t* byte_off, const uint8_t* bytes,
                   const uint64_t* op_off, const void* ops, int op_width,
                   const uint8_t* ctx_init, uint32_t n_ctx, int per_stream_init,
                   uint8_t* out_bins, uint8_t* finish_ok, int n_threads, const char* tmpdir) {
  if (n_threads < 1) n_threads = 1;
  OpView ov(ops, op_width);
  std::vector<std::vector<ContextModel>> ctxs(n_threads);
  for (auto& v : ctxs) v.resize(n_ctx ? n_ctx : 1);
  std::atomic<uint32_t> next(0);
  std::atomic<int> err(0);
  auto work = [&](int tid) {
    std::string fn = scratch_name(tmpdir, 1000 + tid);
    std::vector<ContextModel>& ctx = ctxs[tid];
    for (;;) {
      uint32_t s = next.fetch_add(1);
      if (s >= n_streams) break;
      FILE* f = fopen(fn.c_str(), "wb");
      if (!f) { err = -1; break; }
      uint64_t len = byte_off[s + 1] - byte_off[s];
      if (len) fwrite(bytes + byte_off[s], 1, len, f);
      fclose(f);
      init_ctx(ctx, ctx_init + (per_stream_init ? (uint64_t)s * n_ctx : 0), n_ctx);
      CABAC_BitstreamFile bs;
      if (!bs.openInputFile(fn.c_str())) { err = -1; break; }
      DecoderProbe dec;
      dec.setBitstream(&bs);
      dec.start();
      for (uint64_t i = op_off[s]; i < op_off[s + 1]; ++i) {
        uint32_t code = ov.at(i) >> 1;
        unsigned int bin = 0;
        if (code == ov.ep) dec.decodeBinEP(bin);
        else if (code == ov.trm) dec.decodeBinTrm(bin);
        else dec.decodeBin(bin, &ctx[code]);
        out_bins[i] = (uint8_t)bin;
      }
      if (finish_ok) {
        unsigned int t = 0;
        dec.decodeBinTrm(t);
        bool stop = ((bs.getLastByteRead() << (8 + dec.bitsNeeded())) & 0xff) == 0x80;
        finish_ok[s] = (uint8_t)((t == 1) && stop);
      }
      bs.closeFile();
    }
    unlink(fn.c_str());
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  return err.load();
}

// Single-stream "engine script" runner for known-answer tests that use the
// multi-bin bypass call: script entries are (kind, a, b):
//   kind 0: encodeBin(bin=a, ctx=b)   kind 1: encodeBinEP(a)
//   kind 2: encodeBinsEP(values=a, n=b)  kind 3: encodeBinTrm(a)
// Returns the stream length (bytes copied to out, up to cap) or -1.
int ref_encode_script(const uint32_t* script, uint32_t n_entries,
                      const uint8_t* ctx_init, uint32_t n_ctx,
                      uint8_t* out, uint32_t cap, uint8_t* ctx_final, const char* tmpdir) {
  std::string fn = scratch_name(tmpdir, 9999);
  std::vector<ContextModel> ctx(n_ctx ? n_ctx : 1);
  init_ctx(ctx, ctx_init, n_ctx);
  CABAC_BitstreamFile bs;
  if (!bs.openOutputFile(fn.c_str())) return -1;
  EncoderWithTrm enc;
  enc.setBitstream(&bs);
  enc.start();
  for (uint32_t i = 0; i < n_entries; ++i) {
    uint32_t k = script[3 * i], a = script[3 * i + 1], b = script[3 * i + 2];
    if (k == 0) enc.encodeBin(a, &ctx[b]);
    else if (k == 1) enc.encodeBinEP(a);
    else if (k == 2) enc.encodeBinsEP(a, (int)b);
    else enc.trm(a);
  }
  enc.finish();
  bs.closeFile();
  if (ctx_final)
    for (uint32_t c = 0; c < n_ctx; ++c)
      ctx_final[c] = (uint8_t)((ctx[c].getState() << 1) | ctx[c].getMps());
  FILE* f = fopen(fn.c_str(), "rb");
  if (!f) return -1;
  int n = (int)fread(out, 1, cap, f);
  fclose(f);
  unlink(fn.c_str());
  return n;
}

// Mirror for decoding: kind 0 decodeBin(ctx=b), 1 decodeBinEP, 2 decodeBinsEP(n=b),
// 3 decodeBinTrm; results[i] = decoded value.  Returns 0 / -1.
int ref_decode_script(const uint32_t* script, uint32_t n_entries,
                      const uint8_t* ctx_init, uint32_t n_ctx,
                      const uint8_t* bytes, uint32_t len, uint32_t* results, const char* tmpdir) {
  std::string fn = scratch_name(tmpdir, 9998);
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) return -1;
  if (len) fwrite(bytes, 1, len, f);
  fclose(f);
  std::vector<ContextModel> ctx(n_ctx ? n_ctx : 1);
  init_ctx(ctx, ctx_init, n_ctx);
  CABAC_BitstreamFile bs;
  if (!bs.openInputFile(fn.c_str())) return -1;
  CABAC_ArithmeticDecoder dec;
  dec.setBitstream(&bs);
  dec.start();
  for (uint32_t i = 0; i < n_entries; ++i) {
    uint32_t k = script[3 * i], b = script[3 * i + 2];
    unsigned int v = 0;
    if (k == 0) dec.decodeBin(v, &ctx[b]);
    else if (k == 1) dec.decodeBinEP(v);
    else if (k == 2) dec.decodeBinsEP(v, (int)b);
    else dec.decodeBinTrm(v);
    results[i] = v;
  }
  bs.closeFile();
  unlink(fn.c_str());
  return 0;
}

}  // extern "C"
