// device_reset.cpp -- MATLAB helper `device_reset` called by matlab/runFDTD.m:42 before every run.  The reference
// resets every CUDA device (matlab/device_reset.cpp:5-17) to get rid of whatever an earlier, aborted run left behind.
// This build has nothing to clean up that way -- the gateway's App lives for one call and its destructor releases
// every allocation, also on the error path (mex_FDTD.cpp) -- and cudaDeviceReset would also destroy MATLAB's own
// gpuArray state, so the helper reports the devices it sees and returns the device blocks the library keeps for the
// next mesh (pfdtd_release_cached_memory) to the driver.
#include "mex.h"

#include "../../include/pfdtd.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  (void)nlhs; (void)plhs; (void)nrhs; (void)prhs;
  int n = 0;
  if (pfdtd_device_count(&n) != PFDTD_OK) n = 0;
  if (n > 0) pfdtd_release_cached_memory(-1);
  mexPrintf("Number of Cuda Devices: %d (nothing to reset: allocations are released at the end of every run)\n", n);
}
