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
