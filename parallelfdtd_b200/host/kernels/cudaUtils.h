// cudaUtils.h -- error convention of the reference's CUDA helpers (reference src/kernels/cudaUtils.h:47-51:
// log + `throw(-1)`), applied to the status codes of the C ABI.  The host layer itself contains no CUDA code.
#pragma once
#include "../../../include/pfdtd.h"
#include <stdexcept>
#include "../logger.h"

#if defined(__CUDACC__) || __has_include(<vector_types.h>)
#include <vector_types.h>
#include <vector_functions.h>
#else
struct uint3 { unsigned int x, y, z; };
static inline uint3 make_uint3(unsigned int x, unsigned int y, unsigned int z) { uint3 r = {x, y, z}; return r; }
#endif

// status of a libpfdtd_b200 call -> the reference's behaviour on a failed CUDA call
inline void pfdtd_safe(int rc, const char* where) {
  if (rc == PFDTD_OK) return;
  c_log_msg(LOG_ERROR, "%s: %s", where, pfdtd_last_error());
  if (rc == PFDTD_ERR_RANGE) throw std::out_of_range(pfdtd_last_error());
  throw(-1);
}
