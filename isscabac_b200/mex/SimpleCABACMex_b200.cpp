// SimpleCABACMex_b200.cpp -- the MATLAB-side binding: a drop-in `mexFunction` for
// CABAC/SimpleCABACMex.cpp (reference :100-472).  Build where MATLAB exists:
//
//   mex -R2017b CXXFLAGS='$CXXFLAGS -std=c++11' -I<repo>/include SimpleCABACMex_b200.cpp \
//       -L<repo>/isscabac_b200 -lisscabac -output SimpleCABACMex
//
// CABAC/cabacWrapper.m (:35-76) then works unchanged: every command string it sends
// (initByProb, initByState, encodeStart, encodeBin, getNumBits, encodeFinish, decodeStart,
// decodeBin, decodeFinish) is repacked into isscabac_mxarg and handed to simplecabac_dispatch,
// which reproduces the reference's arity checks and error texts.  Two extra commands expose
// the batch path (one call per matrix instead of one call per bin):
//
//   [payload, byteOff] = SimpleCABACMex('encodeSymbols', cfg, symbols, symOff, ctxInit)
//   symbols            = SimpleCABACMex('decodeSymbols', cfg, payload, byteOff, symOff, ctxInit)
//
// with cfg = [profile method Nq Nlbp types rows] (include/isscabac.h: isscabac_symcfg), symbols
// as doubles holding integers (the reference's convention), symOff/byteOff zero-based offset
// tables and ctxInit the state bytes ((state<<1)+mps) of every context.
//
// Only the mx* calls that also exist in oracle/mexstub/mex.h are used, so the CPU test-suite can
// compile this file without MATLAB.
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "mex.h"

#include "isscabac.h"

namespace {

const double* dbl(const mxArray* a) { return mxGetPr(a); }
size_t count(const mxArray* a) { return mxGetNumberOfElements(a); }

isscabac_symcfg read_cfg(const mxArray* a) {
  if (!mxIsDouble(a) || count(a) < 6) mexErrMsgTxt("Error: cfg must be [profile method Nq Nlbp types rows]\n");
  const double* d = dbl(a);
  isscabac_symcfg c;
  c.profile = (int32_t)d[0]; c.method = (int32_t)d[1]; c.Nq = (uint32_t)d[2];
  c.Nlbp = (int32_t)d[3]; c.types = (uint32_t)d[4]; c.rows = (uint32_t)d[5];
  return c;
}
template <class T>
std::vector<T> to_vec(const mxArray* a) {
  if (!mxIsDouble(a)) mexErrMsgTxt("Error: numeric arguments must be doubles\n");
  const double* d = dbl(a);
  std::vector<T> v(count(a));
  for (size_t i = 0; i < v.size(); ++i) v[i] = (T)d[i];
  return v;
}
mxArray* from_vec(const uint8_t* p, size_t n) {
  mxArray* a = mxCreateDoubleMatrix(1, n, mxREAL);
  double* d = mxGetPr(a);
  for (size_t i = 0; i < n; ++i) d[i] = p[i];
  return a;
}
void fail(int rc) {
  static std::string msg;
  msg = std::string("Error: ") + isscabac_strerror(rc) + ": " + isscabac_last_error() + "\n";
  mexErrMsgTxt(msg.c_str());
}

void encode_symbols(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 5) mexErrMsgTxt("Error: encodeSymbols needs cfg, symbols, symOff, ctxInit\n");
  const isscabac_symcfg cfg = read_cfg(prhs[1]);
  const std::vector<uint32_t> sym = to_vec<uint32_t>(prhs[2]);
  const std::vector<uint64_t> off = to_vec<uint64_t>(prhs[3]);
  const std::vector<uint8_t> ctx = to_vec<uint8_t>(prhs[4]);
  if (off.size() < 2) mexErrMsgTxt("Error: symOff needs at least two entries\n");
  const uint32_t n = (uint32_t)off.size() - 1;
  std::vector<uint8_t> payload(sym.size() * 9 + 16 * (size_t)n + 64);   // EG-k of a 32-bit value: <= 67 bins
  std::vector<uint64_t> boff(n + 1);
  int rc = cabac_encode_symbols_host(&cfg, n, off.data(), sym.data(), 4, ctx.data(), (uint32_t)ctx.size(), 0,
                                     payload.data(), payload.size(), boff.data(), nullptr);
  if (rc) fail(rc);
  if (nlhs > 0) plhs[0] = from_vec(payload.data(), (size_t)boff[n]);
  if (nlhs > 1) {
    plhs[1] = mxCreateDoubleMatrix(1, n + 1, mxREAL);
    for (uint32_t i = 0; i <= n; ++i) mxGetPr(plhs[1])[i] = (double)boff[i];
  }
}

void decode_symbols(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 6) mexErrMsgTxt("Error: decodeSymbols needs cfg, payload, byteOff, symOff, ctxInit\n");
  const isscabac_symcfg cfg = read_cfg(prhs[1]);
  std::vector<uint8_t> payload = to_vec<uint8_t>(prhs[2]);
  const std::vector<uint64_t> boff = to_vec<uint64_t>(prhs[3]);
  const std::vector<uint64_t> off = to_vec<uint64_t>(prhs[4]);
  const std::vector<uint8_t> ctx = to_vec<uint8_t>(prhs[5]);
  if (off.size() < 2 || boff.size() != off.size()) mexErrMsgTxt("Error: offset tables must have n+1 entries\n");
  const uint32_t n = (uint32_t)off.size() - 1;
  std::vector<uint32_t> sym((size_t)off[n] + 1);
  std::vector<uint8_t> ok(n);
  payload.resize(payload.size() + 16);
  int rc = cabac_decode_symbols_host(&cfg, n, boff.data(), payload.data(), off.data(), ctx.data(), (uint32_t)ctx.size(),
                                     0, sym.data(), 4, ok.data());
  if (rc) fail(rc);
  for (uint32_t i = 0; i < n; ++i)
    if (!ok[i]) mexErrMsgTxt("Error: bitstream not terminated properly\n");   // Decoder::finish() asserts, Decoder.cpp:75-81
  if (nlhs > 0) {
    plhs[0] = mxCreateDoubleMatrix(1, (size_t)off[n], mxREAL);
    for (uint64_t i = 0; i < off[n]; ++i) mxGetPr(plhs[0])[i] = (double)sym[i];
  }
}

// [trace, stats] = SimpleCABACMex('getEncoderStats' | 'getDecoderStats', handle, ctxIdx): the typed
// outputs of SimpleCABACMex.cpp:374,393 (uint8 5 x M, uint32 128 x 128), filled in the same memory order
void stats(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  std::vector<isscabac_mxarg> args((size_t)nrhs);
  std::vector<std::string> strs((size_t)nrhs);
  memset(args.data(), 0, args.size() * sizeof(isscabac_mxarg));
  for (int i = 0; i < nrhs; ++i) {
    if (mxIsClass(prhs[i], "char")) {
      char* c = mxArrayToString(prhs[i]);
      strs[i] = c ? c : "";
      args[i].is_char = 1; args[i].s = strs[i].c_str(); args[i].m = 1; args[i].n = (int32_t)strs[i].size();
    } else {
      args[i].d = mxGetPr(prhs[i]); args[i].m = (int32_t)mxGetM(prhs[i]); args[i].n = (int32_t)mxGetN(prhs[i]);
    }
  }
  static char err[512];
  uint64_t n = 0;
  if (simplecabac_dispatch_stats(nlhs, nrhs, args.data(), nullptr, 0, &n, nullptr, err, (int)sizeof err)) mexErrMsgTxt(err);
  plhs[0] = mxCreateNumericMatrix(5, (size_t)n, mxUINT8_CLASS, mxREAL);
  plhs[1] = mxCreateNumericMatrix(128, 128, mxUINT32_CLASS, mxREAL);
  if (simplecabac_dispatch_stats(nlhs, nrhs, args.data(), (uint8_t*)mxGetData(plhs[0]), n, &n,
                                 (uint32_t*)mxGetData(plhs[1]), err, (int)sizeof err))
    mexErrMsgTxt(err);
}

}  // namespace

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs >= 1 && mxIsClass(prhs[0], "char")) {
    char* c = mxArrayToString(prhs[0]);
    const std::string cmd(c ? c : "");
    if (cmd == "encodeSymbols") { encode_symbols(nlhs, plhs, nrhs, prhs); return; }
    if (cmd == "decodeSymbols") { decode_symbols(nlhs, plhs, nrhs, prhs); return; }
    if (cmd == "getEncoderStats" || cmd == "getDecoderStats") { stats(nlhs, plhs, nrhs, prhs); return; }
  }
  // the reference's own command set: repack and dispatch
  std::vector<isscabac_mxarg> args((size_t)(nrhs > 0 ? nrhs : 1));
  std::vector<std::string> strs((size_t)(nrhs > 0 ? nrhs : 1));
  memset(args.data(), 0, args.size() * sizeof(isscabac_mxarg));
  for (int i = 0; i < nrhs; ++i) {
    if (mxIsClass(prhs[i], "char")) {
      char* c = mxArrayToString(prhs[i]);
      strs[i] = c ? c : "";
      args[i].is_char = 1;
      args[i].s = strs[i].c_str();
      args[i].m = 1;
      args[i].n = (int32_t)strs[i].size();
    } else {
      args[i].is_char = 0;
      args[i].d = mxGetPr(prhs[i]);
      args[i].m = (int32_t)mxGetM(prhs[i]);
      args[i].n = (int32_t)mxGetN(prhs[i]);
    }
  }
  double out[4] = {0, 0, 0, 0};
  int out_n = 0;
  static char err[512];
  if (simplecabac_dispatch(nlhs, out, 4, &out_n, nrhs, args.data(), err, (int)sizeof err)) mexErrMsgTxt(err);
  if (out_n > 0) {   // like the reference, the value is returned even when nlhs == 0 (ans)
    plhs[0] = mxCreateDoubleMatrix(1, 1, mxREAL);
    mxGetPr(plhs[0])[0] = out[0];
  }
}
