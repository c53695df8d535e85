// dispatch.cpp -- the MEX command protocol of the reference (CABAC/SimpleCABACMex.cpp:100-472)
// as a plain C function: one call = one mexFunction invocation, same command strings, same
// arity checks in the same order, same error texts.  A MATLAB build only needs a 20-line
// mexFunction that repacks mxArray* into isscabac_mxarg.
//
// Deviations (all where the reference is undefined):
//  * handles are validated against a registry instead of being dereferenced blindly
//    (SimpleCABACMex.cpp:84-97 casts any double to a pointer);
//  * `initBy*` with exactly two arguments reports "invalid context initialization" instead of
//    reading prhs[2] out of bounds (SimpleCABACMex.cpp:123-143 lets nrhs == 2 through);
//  * decodeFinish on a corrupt stream raises an error instead of assert()-aborting.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/isscabac.h"

namespace {

std::mutex g_mu;
std::set<simplecabac*> g_handles;

struct MexError {
  std::string msg;
};
[[noreturn]] void mex_err(const char* m) { throw MexError{m}; }

const char* kNeedHandle = "Error: You need to provide the pointer to the initialized CABAC engine \n";

bool is_char(const isscabac_mxarg& a) { return a.is_char != 0; }
size_t numel(const isscabac_mxarg& a) { return is_char(a) ? (a.s ? strlen(a.s) : 0) : (size_t)a.m * (size_t)a.n; }

// getPointer, SimpleCABACMex.cpp:84-97
simplecabac* get_handle(const isscabac_mxarg* args) {
  const isscabac_mxarg& a = args[1];
  uintptr_t v = (!is_char(a) && a.d && numel(a) >= 1) ? (uintptr_t)a.d[0] : 0;
  simplecabac* h = reinterpret_cast<simplecabac*>(v);
  std::lock_guard<std::mutex> lk(g_mu);
  if (h == nullptr || !g_handles.count(h)) mex_err("Error: No initialized CABAC instance provided \n");
  return h;
}

void check_rc(int rc) {
  if (rc == ISSCABAC_OK) return;
  if (rc == ISSCABAC_ERR_IO) mex_err("Error: bitstreamfile access error\n");
  std::string m = std::string("Error: ") + isscabac_strerror(rc) + ": " + isscabac_last_error() + "\n";
  throw MexError{m};
}

void run(int nlhs, double* out, int out_cap, int* out_n, int nrhs, const isscabac_mxarg* args) {
  if (nrhs < 1 || !is_char(args[0]) || !args[0].s || strlen(args[0].s) >= 64)
    mex_err("Error: input 0 must be a valid keyword - init, encodeStart, encodeBin, encodeFinish, decodeStart, decodeBin, decodeFinish\n");
  const std::string cmd(args[0].s);
  auto put = [&](double v) {
    if (out && out_cap > 0) { out[0] = v; if (out_n) *out_n = 1; }
  };

  if (cmd == "initByState" || cmd == "initByProb") {
    if (nrhs < 2) mex_err("Error: please provide the filename string and the context initializations\n");
    if (nrhs > 3) mex_err("Error: too many parameters, provide the filename and the context initializations");
    if (!is_char(args[1])) mex_err("Error: invalid filename \n");
    if (nrhs < 3 || is_char(args[2])) mex_err("Error: invalid context initialization\n");
    simplecabac* h = nullptr;
    check_rc(simplecabac_create(&h, args[1].s ? args[1].s : ""));
    int rc;
    if (cmd == "initByState") rc = simplecabac_init_by_state(h, args[2].d, (uint32_t)(numel(args[2]) / 3));
    else rc = simplecabac_init_by_prob(h, args[2].d, (uint32_t)numel(args[2]));
    if (rc) { simplecabac_destroy(h); check_rc(rc); }
    {
      std::lock_guard<std::mutex> lk(g_mu);
      g_handles.insert(h);
    }
    put((double)(uintptr_t)h);  // the handle travels as a MATLAB double, like SimpleCABACMex.cpp:152-154
  } else if (cmd == "encodeStart") {
    if (nrhs < 2) mex_err(kNeedHandle);
    check_rc(simplecabac_encode_start(get_handle(args)));
  } else if (cmd == "encodeBin") {
    if (nrhs < 2) mex_err(kNeedHandle);
    simplecabac* h = get_handle(args);
    if (nrhs != 4) mex_err("Error: invalid input, provide the bin and context index to be encoded with\n");
    unsigned bin = (unsigned)(args[2].d ? args[2].d[0] : 2.0);
    int ctx = (int)(args[3].d ? args[3].d[0] : 0.0);
    if (bin != 0 && bin != 1) mex_err("Error: invalid input 3, bin to be encoded should either be 1 or 0\n");
    check_rc(simplecabac_encode_bin(h, bin, (unsigned)ctx));
  } else if (cmd == "getNumBits") {
    if (nrhs < 2) mex_err(kNeedHandle);
    simplecabac* h = get_handle(args);
    if (nrhs != 2) mex_err("Error: invalid input\n");
    uint64_t bits = 0;
    check_rc(simplecabac_get_num_bits(h, &bits));
    put((double)(uint32_t)bits);  // the reference counter is an unsigned int
  } else if (cmd == "encodeFinish") {
    if (nrhs < 2) mex_err(kNeedHandle);
    check_rc(simplecabac_encode_finish(get_handle(args)));
  } else if (cmd == "decodeStart") {
    if (nrhs < 2) mex_err(kNeedHandle);
    check_rc(simplecabac_decode_start(get_handle(args)));
  } else if (cmd == "decodeBin") {
    if (nrhs < 2) mex_err(kNeedHandle);
    if (nlhs != 1) mex_err("Error: invalid command, provide a variable to store the decoded bin \n");
    simplecabac* h = get_handle(args);
    int ctx = (int)((nrhs > 2 && args[2].d) ? args[2].d[0] : 0.0);
    unsigned bin = 0;
    check_rc(simplecabac_decode_bin(h, (unsigned)ctx, &bin));
    put((double)bin);
  } else if (cmd == "decodeFinish") {
    if (nrhs < 2) mex_err(kNeedHandle);
    check_rc(simplecabac_decode_finish(get_handle(args)));
  } else if (cmd == "setTrace") {
    // not in the reference: its Windows builds always trace (CommonDef.h:39-40), its Linux builds never do
    if (nrhs < 2) mex_err(kNeedHandle);
    simplecabac* h = get_handle(args);
    check_rc(simplecabac_set_trace(h, (nrhs > 2 && args[2].d) ? (int)args[2].d[0] : 1));
  } else if (cmd == "destroy") {
    // not in the reference (its instances are leaked by design, SimpleCABACMex.cpp:149-150)
    if (nrhs < 2) mex_err(kNeedHandle);
    simplecabac* h = get_handle(args);
    {
      std::lock_guard<std::mutex> lk(g_mu);
      g_handles.erase(h);
    }
    simplecabac_destroy(h);
  } else {
    mex_err("Error: Invalid Command\n");
  }
}

}  // namespace

// getEncoderStats / getDecoderStats, SimpleCABACMex.cpp:356-466 (same checks, same order, same texts)
extern "C" int simplecabac_dispatch_stats(int nlhs, int nrhs, const isscabac_mxarg* args, uint8_t* steps5,
                                          uint64_t cap_steps, uint64_t* n_steps, uint32_t* trans, char* err, int errcap) {
  if (n_steps) *n_steps = 0;
  if (err && errcap > 0) err[0] = 0;
  try {
    if (nrhs < 1 || !is_char(args[0]) || !args[0].s) mex_err("Error: Invalid Command\n");
    const std::string cmd(args[0].s);
    if (cmd != "getEncoderStats" && cmd != "getDecoderStats") mex_err("Error: Invalid Command\n");
    if (nrhs < 2) mex_err(kNeedHandle);
    if (nlhs != 2) mex_err("Error: invalid command, provide two variables to store the trace \n");
    simplecabac* h = get_handle(args);
    const int ctx = (int)((nrhs > 2 && args[2].d) ? args[2].d[0] : 0.0);
    check_rc(simplecabac_get_stats(h, cmd == "getDecoderStats", (unsigned)ctx, steps5, cap_steps, n_steps, trans));
  } catch (const MexError& e) {
    if (err && errcap > 0) {
      strncpy(err, e.msg.c_str(), (size_t)errcap - 1);
      err[errcap - 1] = 0;
    }
    return 1;
  } catch (...) {
    if (err && errcap > 0) snprintf(err, (size_t)errcap, "Error: internal failure\n");
    return 1;
  }
  return 0;
}

extern "C" int simplecabac_dispatch(int nlhs, double* out, int out_cap, int* out_n,
                                    int nrhs, const isscabac_mxarg* args, char* err, int errcap) {
  if (out_n) *out_n = 0;
  if (err && errcap > 0) err[0] = 0;
  try {
    run(nlhs, out, out_cap, out_n, nrhs, args);
  } catch (const MexError& e) {
    if (err && errcap > 0) {
      strncpy(err, e.msg.c_str(), (size_t)errcap - 1);
      err[errcap - 1] = 0;
    }
    return 1;
  } catch (...) {
    if (err && errcap > 0) snprintf(err, (size_t)errcap, "Error: internal failure\n");
    return 1;
  }
  return 0;
}
