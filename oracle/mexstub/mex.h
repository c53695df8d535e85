// TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for MATLAB's mex.h / matrix.h: just enough of the mx*/mex*
// API for the UNMODIFIED reference SimpleCABACMex.cpp (and the
// CABAC_ContextModelsInit.cpp it includes) to compile and run without MATLAB.
// mexErrMsgTxt throws, mirroring MATLAB's "abort the MEX call" semantics.
#ifndef ISSCABAC_ORACLE_MEXSTUB_H
#define ISSCABAC_ORACLE_MEXSTUB_H
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

struct mxArray {
  bool is_char = false;
  std::string s;
  std::vector<double> d;
  size_t m = 0, n = 0;
};
enum mxComplexity { mxREAL = 0, mxCOMPLEX = 1 };

struct MexStubError : public std::runtime_error {
  explicit MexStubError(const char* msg) : std::runtime_error(msg) {}
};

inline bool mxIsClass(const mxArray* a, const char* cls) {
  if (!a) return false;
  if (!strcmp(cls, "char")) return a->is_char;
  if (!strcmp(cls, "double")) return !a->is_char;
  return false;
}
inline bool mxIsDouble(const mxArray* a) { return a && !a->is_char; }
inline int mxGetString(const mxArray* a, char* buf, size_t buflen) {
  if (!a || !a->is_char || a->s.size() + 1 > buflen) return 1;
  memcpy(buf, a->s.c_str(), a->s.size() + 1);
  return 0;
}
inline char* mxArrayToString(const mxArray* a) {
  char* p = new char[a->s.size() + 1];  // leaked like the reference leaks it
  memcpy(p, a->s.c_str(), a->s.size() + 1);
  return p;
}
inline double* mxGetPr(const mxArray* a) { return const_cast<double*>(a->d.data()); }
inline mxArray* mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity) {
  mxArray* a = new mxArray;
  a->m = m; a->n = n; a->d.assign(m * n, 0.0);
  return a;
}
inline size_t mxGetNumberOfElements(const mxArray* a) { return a->is_char ? a->s.size() : a->d.size(); }
inline size_t mxGetM(const mxArray* a) { return a->m; }
inline size_t mxGetN(const mxArray* a) { return a->n; }
inline size_t mxGetElementSize(const mxArray* a) { return a->is_char ? 2 : sizeof(double); }
inline void mexMakeMemoryPersistent(void*) {}
inline void mexErrMsgTxt(const char* msg) { throw MexStubError(msg); }
inline int mexPrintf(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); int r = vfprintf(stderr, fmt, ap); va_end(ap); return r;
}
#endif
