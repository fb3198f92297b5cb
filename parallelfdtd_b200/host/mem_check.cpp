// mem_check.cpp -- MATLAB helper `mem = mem_check()` used by matlab/runFDTD.m:19 before a run: a column vector with the
// free memory of every CUDA device in bytes (reference matlab/mem_check.cpp:5-23), through the C ABI's device queries.
#include "mex.h"

#include "../../include/pfdtd.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  (void)nlhs; (void)nrhs; (void)prhs;
  int n = 0;
  if (pfdtd_device_count(&n) != PFDTD_OK) n = 0;             // no device: an empty vector, runFDTD.m then refuses the run
  plhs[0] = mxCreateNumericMatrix((size_t)n, 1, mxDOUBLE_CLASS, mxREAL);
  double* out = (double*)mxGetData(plhs[0]);
  for (int i = 0; i < n; i++) {
    int total_mb = 0, free_mb = 0;
    if (pfdtd_device_mem_mb(i, &total_mb, &free_mb) != PFDTD_OK) mexErrMsgTxt(pfdtd_last_error());
    mexPrintf("Memory on device %d, %d MB\n", i, free_mb);
    out[i] = (double)free_mb * 1048576.0;
  }
}
