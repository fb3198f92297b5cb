// mex.h -- a stand-in for MATLAB's MEX C API, just large enough to compile and drive
// parallelfdtd_b200/host/mex_FDTD.cpp from tests/cpp/mex_tests.cpp (MATLAB is not part of this image).  Column-major
// numeric matrices only.  mexErrMsgTxt throws MexError (MATLAB unwinds out of the gateway; so does this).
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

typedef enum { mxUNKNOWN_CLASS = 0, mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7, mxUINT8_CLASS = 9, mxINT32_CLASS = 12, mxUINT32_CLASS = 13 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef size_t mwSize;

struct mxArray_tag {
  mxClassID cls;
  size_t m, n;
  void* data;
};
typedef struct mxArray_tag mxArray;

struct MexError : std::runtime_error {
  explicit MexError(const std::string& s) : std::runtime_error(s) {}
};

inline size_t mxStubElementSize(mxClassID c) { return c == mxDOUBLE_CLASS ? 8 : (c == mxUINT8_CLASS ? 1 : 4); }
inline mxArray* mxCreateNumericMatrix(size_t m, size_t n, mxClassID cls, mxComplexity) {
  mxArray* a = (mxArray*)std::malloc(sizeof(mxArray));
  a->cls = cls; a->m = m; a->n = n;
  const size_t bytes = m * n * mxStubElementSize(cls);
  a->data = std::calloc(bytes ? bytes : 1, 1);
  return a;
}
inline void mxDestroyArray(mxArray* a) { if (a) { std::free(a->data); std::free(a); } }
inline size_t mxGetM(const mxArray* a) { return a->m; }
inline size_t mxGetN(const mxArray* a) { return a->n; }
inline size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
inline void* mxGetData(const mxArray* a) { return a->data; }
inline mxClassID mxGetClassID(const mxArray* a) { return a->cls; }
inline bool mxIsSingle(const mxArray* a) { return a->cls == mxSINGLE_CLASS; }
inline bool mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
inline bool mxIsUint32(const mxArray* a) { return a->cls == mxUINT32_CLASS; }

inline std::string& mexStubLog() { static std::string s; return s; }
inline int mexPrintf(const char* fmt, ...) {
  char buf[2048]; va_list ap; va_start(ap, fmt); int n = std::vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  mexStubLog() += buf;
  return n;
}
inline void mexErrMsgTxt(const char* msg) { throw MexError(msg ? msg : ""); }
inline int mexEvalString(const char*) { return 0; }

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
