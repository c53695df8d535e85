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
  // typed numeric matrices (mxCreateNumericMatrix): raw little-endian element storage
  int class_id = 0;                 // 0 = double / char, else mxUINT8_CLASS / mxUINT32_CLASS
  std::vector<unsigned char> raw;
};
enum mxComplexity { mxREAL = 0, mxCOMPLEX = 1 };
enum mxClassID { mxDOUBLE_CLASS = 6, mxUINT8_CLASS = 9, mxUINT32_CLASS = 13 };   // MATLAB's own numbering

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
inline mxArray* mxCreateNumericMatrix(size_t m, size_t n, mxClassID cls, mxComplexity) {
  mxArray* a = new mxArray;
  a->m = m; a->n = n; a->class_id = (int)cls;
  if (cls == mxDOUBLE_CLASS) { a->class_id = 0; a->d.assign(m * n, 0.0); }
  else a->raw.assign(m * n * (cls == mxUINT32_CLASS ? 4 : 1), 0);
  return a;
}
inline void* mxGetData(const mxArray* a) {
  return a->class_id ? (void*)const_cast<unsigned char*>(a->raw.data()) : (void*)const_cast<double*>(a->d.data());
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
