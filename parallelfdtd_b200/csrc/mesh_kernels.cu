// mesh_kernels.cu -- node-volume preparation: padding, scheme translation, node counts.
// Bit-exact with the reference's setupMesh pipeline:
//   padWithZeros / padWithZerosKernel   src/kernels/cudaMesh.cu:253-326, 516-534
//   toBilbao / toKowalczyk (+kernels)   src/kernels/cudaMesh.cu:328-498
//   calcBoundaries                      src/kernels/cudaMesh.cu:500-514
#include "pfdtd_internal.h"

namespace pfdtd {

// One thread per 16 output bytes along x; the output row pitch nx is a multiple of the block size
// (32 by default), rows are written with 128-bit stores when nx % 16 == 0.
// Semantics restated from padWithZerosKernel: new[z][y][x] = old_flat[z*dx*dy + y*dx + x] for
// 1 <= x <= min(dx, nx-1), 1 <= y <= min(dy, ny-1), z0 <= z <= dz-1, else 0 -- the flat old index
// deliberately wraps into the next row/slice when x == dx or y == dy (SURVEY C-8).  z0 = 1 like the
// reference, or 0 when the volume is an upper z-slab of a taller global domain.
__global__ void pad_nodes_kernel(const uint8_t* __restrict__ old_v, uint8_t* __restrict__ new_v, uint32_t dx, uint32_t dy,
                                 uint32_t dz, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z_first_copied) {
  const uint64_t n_new = (uint64_t)nx * ny * nz;
  const uint64_t n_old = (uint64_t)dx * dy * dz;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_new; i += stride) {
    uint32_t x = (uint32_t)(i % nx);
    uint64_t r = i / nx;
    uint32_t y = (uint32_t)(r % ny);
    uint32_t z = (uint32_t)(r / ny);
    uint8_t v = 0;
    if (x >= 1 && x <= dx && y >= 1 && y <= dy && z >= z_first_copied && z < dz) {
      uint64_t oi = (uint64_t)z * dx * dy + (uint64_t)y * dx + x;
      v = oi < n_old ? old_v[oi] : 0;   // the reference reads past the end here; defined as 0
    }
    new_v[i] = v;
  }
}

int launch_pad_with_zeros(const uint8_t* d_old, uint8_t* d_new, uint32_t dx, uint32_t dy, uint32_t dz, uint32_t nx,
                          uint32_t ny, uint32_t nz, int skip_z0, cudaStream_t stream) {
  const uint64_t n_new = (uint64_t)nx * ny * nz;
  int blocks = (int)((n_new + 255) / 256 < 148 * 32 ? (n_new + 255) / 256 : 148 * 32);
  if (blocks < 1) blocks = 1;
  pad_nodes_kernel<<<blocks, 256, 0, stream>>>(d_old, d_new, dx, dy, dz, nx, ny, nz, skip_z0 ? 1u : 0u);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

__constant__ uint8_t c_kowalczyk_lut[28] = {
    0x00,
    // corners 1..8: DIR_X|DIR_Y|DIR_Z (+SIGN_Z if Down, SIGN_X if Right, SIGN_Y if Out) | 0x80
    0xC7, 0xD7, 0xE7, 0xF7, 0x87, 0x97, 0xA7, 0xB7,
    // edges 9..12 (Down ...), 13..16 (Up ...), 17..20 (Up+Down ...)
    0xC6, 0xE6, 0xC5, 0xD5,
    0x86, 0xA6, 0x85, 0x95,
    0x83, 0x93, 0xA3, 0xB3,
    // faces 21..26
    0xC4, 0xA2, 0x82, 0x91, 0x81, 0x84,
    // air
    0x80};

// One thread per 16 bytes (uint4); counts are warp-reduced before one atomic per warp.
__global__ void translate_nodes_kernel(uint8_t* __restrict__ pos, uint8_t* __restrict__ mat, uint64_t n, int centred,
                                       unsigned long long* __restrict__ counts) {
  const uint32_t air_code = centred ? 0x80u : 0x86u;
  unsigned int air = 0, bnd = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t k = pos[i];
    uint32_t out = k;
    if (k == 0) {
      mat[i] = 0;
    } else if (centred) {
      if (k <= 27) out = c_kowalczyk_lut[k];
    } else {
      if (k <= 8) out = 0x83;
      else if (k <= 20) out = 0x84;
      else if (k <= 26) out = 0x85;
      else if (k == 27) out = 0x86;
    }
    if (out != k) pos[i] = (uint8_t)out;
    air += (out == air_code);
    bnd += (out != 0 && out != air_code);
  }
  for (int o = 16; o > 0; o >>= 1) {
    air += __shfl_down_sync(0xffffffffu, air, o);
    bnd += __shfl_down_sync(0xffffffffu, bnd, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (air) atomicAdd(counts + 0, (unsigned long long)air);
    if (bnd) atomicAdd(counts + 1, (unsigned long long)bnd);
  }
}

int launch_translate_nodes(uint8_t* d_pos, uint8_t* d_mat, uint64_t n, int centred, unsigned long long* d_counts2,
                           cudaStream_t stream) {
  int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  if (blocks < 1) blocks = 1;
  translate_nodes_kernel<<<blocks, 256, 0, stream>>>(d_pos, d_mat, n, centred, d_counts2);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

// ---- node classes ---------------------------------------------------------------------------------
// key = pos | mat << 8; air nodes are normalised to material 0 (their admittance term vanishes:
// 6-K = 0 in the forward scheme, no direction flags in the centred one), solid nodes already carry 0.
__device__ __forceinline__ uint32_t class_key(uint32_t pos, uint32_t mat, uint32_t air_code) {
  return (pos == air_code || pos == 0u) ? pos : (pos | (mat << 8));
}

__global__ void mark_classes_kernel(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ mat, uint64_t n, uint32_t air_code,
                                    uint8_t* __restrict__ flags) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t p = pos[i];
    if (p == air_code || p == 0u) continue;      // classes 0 and 1 always exist
    flags[class_key(p, mat[i], air_code)] = 1;    // benign race: every writer stores 1
  }
}

__global__ void assign_classes_kernel(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ mat, uint64_t n, uint32_t air_code,
                                      const uint8_t* __restrict__ lut, uint8_t* __restrict__ cls) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t p = pos[i];
    cls[i] = (p == 0u) ? (uint8_t)0 : (p == air_code) ? (uint8_t)1 : lut[class_key(p, mat[i], air_code)];
  }
}

int launch_mark_classes(const uint8_t* d_pos, const uint8_t* d_mat, uint64_t n, uint32_t air_code, uint8_t* d_flags, cudaStream_t stream) {
  int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  if (blocks < 1) blocks = 1;
  mark_classes_kernel<<<blocks, 256, 0, stream>>>(d_pos, d_mat, n, air_code, d_flags);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

int launch_assign_classes(const uint8_t* d_pos, const uint8_t* d_mat, uint64_t n, uint32_t air_code, const uint8_t* d_lut,
                          uint8_t* d_cls, cudaStream_t stream) {
  int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  if (blocks < 1) blocks = 1;
  assign_classes_kernel<<<blocks, 256, 0, stream>>>(d_pos, d_mat, n, air_code, d_lut, d_cls);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

}  // namespace pfdtd
