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

// ---- allocation / copy helpers with the reference's names and argument order (reference cudaUtils.h:59-171), over the
// C ABI's device-memory entry points.  mem_size counts ELEMENTS.  Failures follow pfdtd_safe (log + throw(-1)).
#include <cstdlib>

template <typename T> T* toDevice(unsigned int mem_size, unsigned int device) {        // zero-initialised
  void* d = 0;
  pfdtd_safe(pfdtd_device_alloc((int)device, (size_t)mem_size * sizeof(T), &d), "toDevice: alloc");
  const unsigned char zero = 0;
  pfdtd_safe(pfdtd_device_fill((int)device, d, (size_t)mem_size * sizeof(T), 1, &zero), "toDevice: clear");
  return (T*)d;
}
template <typename T> T* toDevice(unsigned int mem_size, const T* h_data, unsigned int device) {
  void* d = 0;
  pfdtd_safe(pfdtd_device_alloc((int)device, (size_t)mem_size * sizeof(T), &d), "toDevice: alloc");
  pfdtd_safe(pfdtd_device_upload((int)device, d, h_data, (size_t)mem_size * sizeof(T)), "toDevice: copy");
  return (T*)d;
}
template <typename T> T* valueToDevice(unsigned int mem_size, T val, unsigned int device) {
  static_assert(sizeof(T) == 1 || sizeof(T) == 2 || sizeof(T) == 4 || sizeof(T) == 8, "valueToDevice: 1, 2, 4 or 8 byte elements");
  void* d = 0;
  pfdtd_safe(pfdtd_device_alloc((int)device, (size_t)mem_size * sizeof(T), &d), "valueToDevice: alloc");
  pfdtd_safe(pfdtd_device_fill((int)device, d, mem_size, sizeof(T), &val), "valueToDevice: fill");
  return (T*)d;
}
// calloc'ed host copy, the caller frees it (reference cudaUtils.h:101-110)
template <typename T> T* fromDevice(unsigned int mem_size, const T* d_data, unsigned int device) {
  T* h = (T*)std::calloc(mem_size ? mem_size : 1, sizeof(T));
  pfdtd_safe(pfdtd_device_download((int)device, h, d_data, (size_t)mem_size * sizeof(T)), "fromDevice");
  return h;
}
template <typename T> void resetData(unsigned int mem_size, T* d_data, unsigned int device) {
  const unsigned char zero = 0;
  pfdtd_safe(pfdtd_device_fill((int)device, d_data, (size_t)mem_size * sizeof(T), 1, &zero), "resetData");
}
template <typename T> T getSample(unsigned int element_idx, T* P) {
  T v = (T)0;
  pfdtd_safe(pfdtd_device_download(-1, &v, P + element_idx, sizeof(T)), "getSample");
  return v;
}
template <typename T> void copyHostToDevice(unsigned int mem_size, T* d_dest, T* h_data, unsigned int device) {
  pfdtd_safe(pfdtd_device_upload((int)device, d_dest, h_data, (size_t)mem_size * sizeof(T)), "copyHostToDevice");
}
template <typename T> void copyDeviceToHost(unsigned int mem_size, T* h_dest, T* d_src, unsigned int device) {
  pfdtd_safe(pfdtd_device_download((int)device, h_dest, d_src, (size_t)mem_size * sizeof(T)), "copyDeviceToHost");
}
template <typename T> void destroyMem(T* d_data) { pfdtd_safe(pfdtd_device_free(-1, d_data), "destroyMem"); }
template <typename T> void destroyMem(T* d_data, unsigned int device) { pfdtd_safe(pfdtd_device_free((int)device, d_data), "destroyMem"); }
inline int getCurrentDevice() { int d = 0; pfdtd_safe(pfdtd_current_device(&d), "getCurrentDevice"); return d; }
