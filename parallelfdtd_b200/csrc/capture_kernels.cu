// capture_kernels.cu -- slice gather for the capture path that runs right after a step
// (reference src/kernels/visualizationUtils.cu:127-254 captureSliceFast + captureSliceKernel, and
// CudaMesh::getSlice / getPositionSlice for the xy orientation).  One kernel gathers the pressure
// and the position byte of one axis-aligned slice of a partition into two dense device buffers,
// so only the slice travels to the host (the reference allocates, launches and copies per
// partition in the same way; the whole-field D2H is never needed).
//
//   orientation 0 (xy at z = slice): out[y][x]          -- contiguous plane, no kernel needed
//   orientation 1 (xz at y = slice): out[z][x]          -- coalesced row reads
//   orientation 2 (yz at x = slice): out[z][y]          -- one 32 B sector per element (strided in x)
#include "pfdtd_internal.h"

namespace pfdtd {

template <typename T>
__global__ void capture_slice_kernel(const T* __restrict__ P, const uint8_t* __restrict__ pos, T* __restrict__ out_p,
                                     uint8_t* __restrict__ out_pos, uint32_t X, uint32_t Y, uint32_t z_lo, uint32_t nz,
                                     uint32_t slice, int orientation) {
  const uint32_t w = orientation == 1 ? X : Y;                      // row length of the output
  const uint64_t n = (uint64_t)w * nz;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t u = (uint32_t)(i % w);
    const uint32_t z = z_lo + (uint32_t)(i / w);
    const uint64_t e = orientation == 1 ? ((uint64_t)z * Y + slice) * X + u : ((uint64_t)z * Y + u) * X + slice;
    out_p[i] = P[e];
    if (out_pos) out_pos[i] = pos[e];
  }
}

int launch_capture_slice(int dtype, const void* P, const uint8_t* pos, void* out_p, uint8_t* out_pos, uint32_t X, uint32_t Y,
                         uint32_t z_lo, uint32_t nz, uint32_t slice, int orientation, cudaStream_t stream) {
  const uint64_t n = (uint64_t)(orientation == 1 ? X : Y) * nz;
  if (n == 0) return PFDTD_OK;
  int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (dtype == PFDTD_F32)
    capture_slice_kernel<float><<<blocks, 256, 0, stream>>>((const float*)P, pos, (float*)out_p, out_pos, X, Y, z_lo, nz, slice, orientation);
  else
    capture_slice_kernel<double><<<blocks, 256, 0, stream>>>((const double*)P, pos, (double*)out_p, out_pos, X, Y, z_lo, nz, slice, orientation);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

}  // namespace pfdtd
