// SimpleCABAC.hpp -- C++ facade over the C ABI (isscabac.h) with the method names of the
// reference engine classes, for callers of
//   CABAC_ArithmeticEncoder  (CABAC/CABAC_ArithmeticEncoder.h:52-72)
//   CABAC_ArithmeticDecoder  (CABAC/CABAC_ArithmeticDecoder.h:46-61)
//   class CABAC              (CABAC/SimpleCABACMex.cpp:69-80: bitstream + both context sets)
// One object = one stream at a time, coder state on the GPU.  Contexts are addressed by
// index (the reference passes ContextModel*; CABAC_ContextModels::getContextModel(idx),
// CABAC_ContextModelsInit.h:62, is the index form used by the MEX layer).
// Header-only; link against isscabac_b200/libisscabac.so.  Errors throw std::runtime_error.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "isscabac.h"

class SimpleCABAC {
 public:
  // filename == nullptr: in-memory sink/source (bytes() / setBytes())
  explicit SimpleCABAC(const char* filename = nullptr) { ck(simplecabac_create(&h_, filename)); }
  ~SimpleCABAC() { simplecabac_destroy(h_); }
  SimpleCABAC(const SimpleCABAC&) = delete;
  SimpleCABAC& operator=(const SimpleCABAC&) = delete;

  // CABAC_ContextModels::initContextModelsByP0Prob / ByMpsState (both context sets, like the MEX init)
  void initByProb(const double* p0, unsigned n) { ck(simplecabac_init_by_prob(h_, p0, n)); }
  void initByState(const double* triples, unsigned n) { ck(simplecabac_init_by_state(h_, triples, n)); }

  // ---- encoder: start / encodeBin / encodeBinEP / encodeBinsEP / encodeBinTrm / finish
  void start() { ck(simplecabac_encode_start(h_)); }
  void encodeBin(unsigned bin, unsigned ctxIdx) { ck(simplecabac_encode_bin(h_, bin, ctxIdx)); }
  void encodeBinEP(unsigned bin) { ck(simplecabac_encode_bin_ep(h_, bin)); }
  void encodeBinsEP(unsigned bins, int numBins) { ck(simplecabac_encode_bins_ep(h_, bins, numBins)); }
  void encodeBinTrm(unsigned bin) { ck(simplecabac_encode_bin_trm(h_, bin)); }
  void finish() { ck(simplecabac_encode_finish(h_)); }
  unsigned getNumBits() { uint64_t b = 0; ck(simplecabac_get_num_bits(h_, &b)); return (unsigned)b; }
  unsigned getBinsCoded() { uint64_t b = 0; ck(simplecabac_get_bins_coded(h_, &b)); return (unsigned)b; }
  std::vector<uint8_t> bytes() const {
    const uint8_t* p = nullptr; uint64_t n = 0;
    ck(simplecabac_get_bytes(h_, &p, &n));
    return std::vector<uint8_t>(p, p + n);
  }

  // ---- decoder: start / decodeBin / decodeBinEP / decodeBinsEP / decodeBinTrm / finish
  void setBytes(const std::vector<uint8_t>& b) { ck(simplecabac_set_bytes(h_, b.data(), b.size())); }
  void decodeStart() { ck(simplecabac_decode_start(h_)); }
  void decodeBin(unsigned& bin, unsigned ctxIdx) { ck(simplecabac_decode_bin(h_, ctxIdx, &bin)); }
  void decodeBinEP(unsigned& bin) { ck(simplecabac_decode_bin_ep(h_, &bin)); }
  void decodeBinsEP(unsigned& bins, int numBins) { ck(simplecabac_decode_bins_ep(h_, numBins, &bins)); }
  void decodeBinTrm(unsigned& bin) { ck(simplecabac_decode_bin_trm(h_, &bin)); }
  void decodeFinish() { ck(simplecabac_decode_finish(h_)); }

  // ContextModel::getState / getMps of context idx in the encoder (0) or decoder (1) set
  void getContext(int decoderSet, unsigned ctxIdx, unsigned& state, unsigned& mps) {
    ck(simplecabac_get_ctx_state(h_, decoderSet, ctxIdx, &state, &mps));
  }
  simplecabac* handle() { return h_; }

 private:
  static void ck(int rc) {
    if (rc != ISSCABAC_OK)
      throw std::runtime_error(std::string("SimpleCABAC: ") + isscabac_strerror(rc) + ": " + isscabac_last_error());
  }
  simplecabac* h_ = nullptr;
};
