// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Compiles the UNMODIFIED reference MEX dispatcher (SimpleCABACMex.cpp, a unity
// build that #includes the whole engine) against the stub mex.h in mexstub/ and
// exposes one C entry point so tests can drive all nine Linux commands
// (SimpleCABACMex.cpp:120-354) and compare them with the product dispatcher
// `simplecabac_dispatch` (include/isscabac.h).  No reference code is copied: the
// reference file is included from where it lies (-I/root/reference/CABAC).
#include "SimpleCABACMex.cpp"

extern "C" {

struct RefMexArg {
  int is_char;
  const char* s;
  const double* d;
  int m, n;
};

// Returns 0 on success, 1 when the reference raised mexErrMsgTxt (text in err).
int refmex_call(int nlhs, double* out, int out_cap, int* out_n,
                int nrhs, const RefMexArg* args, char* err, int errcap) {
  std::vector<mxArray> store(nrhs);
  std::vector<const mxArray*> prhs(nrhs + 4, nullptr);
  for (int i = 0; i < nrhs; ++i) {
    store[i].is_char = args[i].is_char != 0;
    if (args[i].is_char) store[i].s = args[i].s;
    else store[i].d.assign(args[i].d, args[i].d + (size_t)args[i].m * args[i].n);
    store[i].m = args[i].m; store[i].n = args[i].n;
    prhs[i] = &store[i];
  }
  mxArray* plhs[4] = {nullptr, nullptr, nullptr, nullptr};
  if (out_n) *out_n = 0;
  try {
    mexFunction(nlhs, plhs, nrhs, prhs.data());
  } catch (const MexStubError& e) {
    if (err && errcap > 0) { strncpy(err, e.what(), errcap - 1); err[errcap - 1] = 0; }
    return 1;
  }
  if (plhs[0] && out) {
    int n = (int)plhs[0]->d.size();
    if (n > out_cap) n = out_cap;
    for (int i = 0; i < n; ++i) out[i] = plhs[0]->d[i];
    if (out_n) *out_n = n;
    delete plhs[0];
  }
  return 0;
}

#if RWTH_TRACE_CABAC_STATES
// getEncoderStats / getDecoderStats (SimpleCABACMex.cpp:356-466; compiled in only when the
// reference is built the way its Windows project builds it, -D_WIN32 -> CommonDef.h:39-40).
// steps: 5 bytes per step [bin state_p mps_p state_a mps_a]; trans: 128*128 u32 in the
// reference's memory order (index state_p*128 + state_a).  Returns the number of steps, -1 on
// a mexErrMsgTxt (text in err).
long refmex_stats(int decoder, double handle, int ctx_idx, unsigned char* steps, long cap_steps, unsigned* trans,
                  char* err, int errcap) {
  mxArray cmd, h, c;
  cmd.is_char = true; cmd.s = decoder ? "getDecoderStats" : "getEncoderStats";
  h.d.assign(1, handle); h.m = h.n = 1;
  c.d.assign(1, (double)ctx_idx); c.m = c.n = 1;
  const mxArray* prhs[3] = {&cmd, &h, &c};
  mxArray* plhs[2] = {nullptr, nullptr};
  try {
    mexFunction(2, plhs, 3, prhs);
  } catch (const MexStubError& e) {
    if (err && errcap > 0) { strncpy(err, e.what(), errcap - 1); err[errcap - 1] = 0; }
    return -1;
  }
  long n = plhs[0] ? (long)plhs[0]->n : 0;
  if (plhs[0] && steps) memcpy(steps, plhs[0]->raw.data(), (size_t)(n < cap_steps ? n : cap_steps) * 5);
  if (plhs[1] && trans) memcpy(trans, plhs[1]->raw.data(), 128 * 128 * 4);
  delete plhs[0]; delete plhs[1];
  return n;
}
#endif

// p(0) -> (state<<1)|mps through the reference's own initContextModelsByP0Prob
// (CABAC_ContextModelsInit.cpp:82-148); n < 1000.
int refmex_prob_to_state(const double* p0, int n, unsigned char* out) {
  mxArray a;
  a.d.assign(p0, p0 + n);
  a.m = 1; a.n = (size_t)n;
  CABAC_ContextModels* cm = new CABAC_ContextModels;
  cm->initContextModelsByP0Prob(n, &a);
  for (int i = 0; i < n; ++i)
    out[i] = (unsigned char)((cm->getContextModel(i)->getState() << 1) | cm->getContextModel(i)->getMps());
  delete cm;
  return 0;
}

// [ctxIdx mps state] triples through initContextModelsByMpsState (:51-80)
int refmex_state_triples(const double* t, int n, unsigned char* out) {
  mxArray a;
  a.d.assign(t, t + 3 * n);
  a.m = 3; a.n = (size_t)n;
  CABAC_ContextModels* cm = new CABAC_ContextModels;
  cm->initContextModelsByMpsState(n, &a);
  for (int i = 0; i < n; ++i)
    out[i] = (unsigned char)((cm->getContextModel(i)->getState() << 1) | cm->getContextModel(i)->getMps());
  delete cm;
  return 0;
}

}  // extern "C"
