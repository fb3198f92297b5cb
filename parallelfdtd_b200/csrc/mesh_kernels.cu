// mesh_kernels.cu -- node-volume preparation: padding, scheme translation, node counts.
// Bit-exact with the reference's setupMesh pipeline:
//   padWithZeros / padWithZerosKernel   src/kernels/cudaMesh.cu:253-326, 516-534
//   toBilbao / toKowalczyk (+kernels)   src/kernels/cudaMesh.cu:328-498
//   calcBoundaries                      src/kernels/cudaMesh.cu:500-514
#include "pfdtd_internal.h"
#include <algorithm>
#include <cstdlib>

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

// ---- fused padWithZeros + toBilbao/toKowalczyk + calcBoundaries: 16 output bytes per thread -----------------
// Same results as pad_nodes_kernel (twice) followed by translate_nodes_kernel, in one pass: each thread produces 16
// consecutive x positions of one output row for BOTH volumes.  The source bytes of a row start at an arbitrary
// byte offset of the old (unpadded) volume, so they are fetched as five aligned 32-bit words per volume and
// funnel-shifted into place; bytes outside the copied range (x = 0, x > dx, rows y = 0 / y > dy, planes below
// z_first_copied or >= dz, reads past the end of the old volume) are masked to zero.  The `bid` -> node byte
// translation goes through a 256-entry table in shared memory (values above 27 pass through like the reference's
// kernels leave them).  Traffic: 2 B read + 2 B written per voxel.
__global__ void __launch_bounds__(256) prepare_nodes_kernel(const uint8_t* __restrict__ old_bid, const uint8_t* __restrict__ old_mat,
                                                            uint8_t* __restrict__ pos, uint8_t* __restrict__ mat, uint32_t dx, uint32_t dy,
                                                            uint32_t dz, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z_first_copied,
                                                            int centred, unsigned long long* __restrict__ counts) {
  __shared__ uint8_t lut[256];
  {
    const uint32_t k = threadIdx.x;
    uint32_t out = k;
    if (centred) { if (k <= 27) out = c_kowalczyk_lut[k]; }
    else if (k == 0) out = 0;
    else if (k <= 8) out = 0x83;
    else if (k <= 20) out = 0x84;
    else if (k <= 26) out = 0x85;
    else if (k == 27) out = 0x86;
    lut[k] = (uint8_t)out;
  }
  __syncthreads();
  const uint32_t air_code = centred ? 0x80u : 0x86u;
  const uint64_t n_old = (uint64_t)dx * dy * dz;
  const uint64_t n_old4 = (n_old + 3) & ~(uint64_t)3;          // cudaMalloc'ed: the last partial word is readable
  const uint32_t cpr = nx / 16;                                 // chunks per row
  const uint64_t n_chunks = (uint64_t)cpr * ny * nz;
  const uint32_t* __restrict__ wb = reinterpret_cast<const uint32_t*>(old_bid);
  const uint32_t* __restrict__ wm = reinterpret_cast<const uint32_t*>(old_mat);
  unsigned int air = 0, bnd = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
    const uint32_t x0 = (uint32_t)(c % cpr) * 16;
    const uint64_t r = c / cpr;
    const uint32_t y = (uint32_t)(r % ny), z = (uint32_t)(r / ny);
    uint32_t ob[4] = {0, 0, 0, 0}, om[4] = {0, 0, 0, 0};
    if (y >= 1 && y <= dy && z >= z_first_copied && z < dz && x0 <= dx) {
      const uint64_t oi0 = (uint64_t)z * dx * dy + (uint64_t)y * dx + x0;
      const uint64_t w0 = oi0 >> 2;
      const uint32_t sh = (uint32_t)(oi0 & 3) * 8;
      uint32_t b[5], m[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const bool ok = (w0 + j) * 4 < n_old4;
        b[j] = ok ? __ldg(wb + w0 + j) : 0u;
        m[j] = ok ? __ldg(wm + w0 + j) : 0u;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        uint32_t vb = __funnelshift_r(b[j], b[j + 1], sh), vm = __funnelshift_r(m[j], m[j + 1], sh);
        uint32_t keep = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint32_t x = x0 + 4 * j + q;
          if (x >= 1 && x <= dx && oi0 + 4 * j + q < n_old) keep |= 0xffu << (8 * q);
        }
        vb &= keep; vm &= keep;
        uint32_t tb = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint32_t k = (vb >> (8 * q)) & 0xffu;
          const uint32_t o = lut[k];
          tb |= o << (8 * q);
          if (k == 0) vm &= ~(0xffu << (8 * q));               // solid nodes carry material 0 (cudaMesh.cu:332-336,367-371)
          air += (o == air_code);
          bnd += (o != 0 && o != air_code);
        }
        ob[j] = tb; om[j] = vm;
      }
    }
    reinterpret_cast<uint4*>(pos)[c] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
    reinterpret_cast<uint4*>(mat)[c] = make_uint4(om[0], om[1], om[2], om[3]);
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

static int stream_blocks(uint64_t items) {
  const uint64_t want = (items + 255) / 256;
  return (int)std::max<uint64_t>(1, std::min<uint64_t>(want, 148ull * 16));
}

int launch_prepare_nodes(const uint8_t* d_old_bid, const uint8_t* d_old_mat, uint8_t* d_pos, uint8_t* d_mat, uint32_t dx, uint32_t dy,
                         uint32_t dz, uint32_t nx, uint32_t ny, uint32_t nz, int skip_z0, int centred, unsigned long long* d_counts2,
                         cudaStream_t stream) {
  const uint64_t n_new = (uint64_t)nx * ny * nz;
  if (nx % 16 == 0) {
    prepare_nodes_kernel<<<stream_blocks(n_new / 16), 256, 0, stream>>>(d_old_bid, d_old_mat, d_pos, d_mat, dx, dy, dz, nx, ny, nz,
                                                                      skip_z0 ? 1u : 0u, centred, d_counts2);
    PF_CUDA(cudaGetLastError());
    return PFDTD_OK;
  }
  PF_TRY(launch_pad_with_zeros(d_old_bid, d_pos, dx, dy, dz, nx, ny, nz, skip_z0, stream));
  PF_TRY(launch_pad_with_zeros(d_old_mat, d_mat, dx, dy, dz, nx, ny, nz, skip_z0, stream));
  return launch_translate_nodes(d_pos, d_mat, n_new, centred, d_counts2, stream);
}

// ---- node classes ---------------------------------------------------------------------------------
// key = pos | mat << 8 | K12 << 16 | K8 << 20 (mat == nullptr: without the material -- the position classes of wide meshes).  K12 / K8 = number of non-solid voxels among the 12 edge- and
// 8 corner-neighbours; only the interpolated (27-point) schemes need them, the 7-point schemes use 0.
// Nodes whose admittance term vanishes (forward byte with K = 6 / centred byte without direction flags)
// are normalised to material 0; solid nodes already carry material 0.
__device__ __forceinline__ uint32_t node_key(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ mat, uint64_t i, uint32_t p,
                                             uint32_t air_code, int interp, uint32_t X, uint32_t Y, uint32_t Z) {
  uint32_t key = (p == air_code || mat == nullptr) ? p : (p | ((uint32_t)mat[i] << 8));   // mat == nullptr: position classes only
  if (interp) {
    const int x = (int)(i % X), y = (int)((i / X) % Y), z = (int)(i / ((uint64_t)X * Y));
    uint32_t k12 = 0, k8 = 0;
#pragma unroll
    for (int dz = -1; dz <= 1; dz++)
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          const int nz_ = (dx != 0) + (dy != 0) + (dz != 0);
          if (nz_ < 2) continue;
          const int xx = x + dx, yy = y + dy, zz = z + dz;
          uint32_t in = 0;
          if (xx >= 0 && yy >= 0 && zz >= 0 && xx < (int)X && yy < (int)Y && zz < (int)Z)
            in = pos[((uint64_t)zz * Y + yy) * X + xx] >> 7;
          if (nz_ == 2) k12 += in; else k8 += in;
        }
    key |= (k12 << 16) | (k8 << 20);
  }
  return key;
}

__device__ __forceinline__ uint32_t key_hash(uint32_t k) { k ^= k >> 15; k *= 0x2c1b3c6dU; k ^= k >> 12; k *= 0x297a2d39U; k ^= k >> 15; return k; }

// open-addressing set of keys (EMPTY = 0xffffffff); `count` = number of distinct keys inserted
__global__ void mark_classes_kernel(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ mat, uint64_t n, uint32_t air_key,
                                    uint32_t air_code, int interp, uint32_t X, uint32_t Y, uint32_t Z, uint32_t* __restrict__ table,
                                    uint32_t cap, uint32_t* __restrict__ count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t p = pos[i];
    if (p == 0u) continue;
    const uint32_t key = node_key(pos, mat, i, p, air_code, interp, X, Y, Z);
    if (key == air_key) continue;                 // classes 0 (solid) and 1 (air) always exist
    uint32_t slot = key_hash(key) % cap;
    for (uint32_t probe = 0; probe < cap; probe++) {
      const uint32_t cur = table[slot];
      if (cur == key) break;
      if (cur == 0xffffffffu) {
        const uint32_t prev = atomicCAS(&table[slot], 0xffffffffu, key);
        if (prev == 0xffffffffu) { atomicAdd(count, 1u); break; }
        if (prev == key) break;
      }
      slot = (slot + 1) % cap;
    }
  }
}

__global__ void assign_classes_kernel(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ mat, uint64_t n, uint32_t air_key,
                                      uint32_t air_code, int interp, uint32_t X, uint32_t Y, uint32_t Z,
                                      const uint32_t* __restrict__ table, const uint8_t* __restrict__ ids, uint32_t cap,
                                      uint8_t* __restrict__ cls) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t p = pos[i];
    uint8_t c = 0;
    if (p != 0u) {
      const uint32_t key = node_key(pos, mat, i, p, air_code, interp, X, Y, Z);
      if (key == air_key) c = 1;
      else {
        uint32_t slot = key_hash(key) % cap;
        for (uint32_t probe = 0; probe < cap; probe++) {
          if (table[slot] == key) { c = ids[slot]; break; }
          slot = (slot + 1) % cap;
        }
      }
    }
    cls[i] = c;
  }
}

// ---- the same two passes, 16 voxels per thread (X % 16 == 0) ---------------------------------------------------
// A thread reads its 16 position and material bytes with one 128-bit load each.  Chunks that are all solid or
// all open air -- nearly all of a room -- are settled from those two words.  For the interpolated schemes the key of
// an inside voxel also needs its edge / corner neighbour counts: when the nine 18-byte rows around the chunk are all
// inside, every voxel of the chunk has the full counts (one test instead of 20 byte loads per voxel); otherwise the
// per-voxel path of node_key is taken.  Runs of equal keys (a row lying along a wall) touch the hash table once.
__device__ __forceinline__ bool chunk_neighbourhood_all_inside(const uint8_t* __restrict__ pos, uint64_t i0, uint32_t X, uint32_t Y, uint32_t Z) {
  const int x0 = (int)(i0 % X), y = (int)((i0 / X) % Y), z = (int)(i0 / ((uint64_t)X * Y));
  if (x0 < 1 || x0 + 16 >= (int)X || y < 1 || y + 1 >= (int)Y || z < 1 || z + 1 >= (int)Z) return false;
  uint32_t all = 0x80808080u;
#pragma unroll
  for (int dz = -1; dz <= 1; dz++)
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      const uint64_t r = ((uint64_t)(z + dz) * Y + (y + dy)) * X + x0;
      const uint4 w = *reinterpret_cast<const uint4*>(pos + r);
      all &= w.x & w.y & w.z & w.w;
      all &= (uint32_t)pos[r - 1] * 0x01010101u;
      all &= (uint32_t)pos[r + 16] * 0x01010101u;
    }
  return (all & 0x80808080u) == 0x80808080u;
}

__device__ __forceinline__ uint32_t byte_of(const uint4& w, int i) {
  const uint32_t v = i < 4 ? w.x : (i < 8 ? w.y : (i < 12 ? w.z : w.w));
  return (v >> (8 * (i & 3))) & 0xffu;
}

__device__ __forceinline__ void class_set_insert(uint32_t* __restrict__ table, uint32_t cap, uint32_t* __restrict__ count, uint32_t key) {
  uint32_t slot = key_hash(key) % cap;
  for (uint32_t probe = 0; probe < cap; probe++) {
    const uint32_t cur = table[slot];
    if (cur == key) return;
    if (cur == 0xffffffffu) {
      const uint32_t prev = atomicCAS(&table[slot], 0xffffffffu, key);
      if (prev == 0xffffffffu) { atomicAdd(count, 1u); return; }
      if (prev == key) return;
    }
    slot = (slot + 1) % cap;
  }
}

template <bool ASSIGN>
__global__ void __launch_bounds__(256) classes16_kernel(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ mat, uint64_t n,
                                                        uint32_t air_key, uint32_t air_code, int interp, uint32_t X, uint32_t Y, uint32_t Z,
                                                        uint32_t* __restrict__ table, const uint8_t* __restrict__ ids, uint32_t cap,
                                                        uint32_t* __restrict__ count, uint8_t* __restrict__ cls) {
  const uint64_t n_chunks = n / 16;
  const uint32_t air4 = air_code * 0x01010101u;
  const uint32_t full = (12u << 16) | (8u << 20);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
    const uint4 pw = reinterpret_cast<const uint4*>(pos)[c];
    uint32_t out[4] = {0, 0, 0, 0};
    const bool all_solid = (pw.x | pw.y | pw.z | pw.w) == 0u;
    const bool all_air = pw.x == air4 && pw.y == air4 && pw.z == air4 && pw.w == air4;
    bool settled = all_solid;
    bool nb_full = false;
    if (!all_solid && interp) nb_full = chunk_neighbourhood_all_inside(pos, c * 16, X, Y, Z);
    if (all_air && (!interp || nb_full)) {
      settled = true;
      out[0] = out[1] = out[2] = out[3] = 0x01010101u;           // class 1
    }
    if (!settled) {
      const uint4 mw = mat != nullptr ? reinterpret_cast<const uint4*>(mat)[c] : make_uint4(0u, 0u, 0u, 0u);
      uint32_t last_key = 0xffffffffu, last_id = 0;
#pragma unroll 1
      for (int i = 0; i < 16; i++) {
        const uint32_t p = byte_of(pw, i);
        if (p == 0u) continue;
        uint32_t key;
        if (!interp) key = (p == air_code) ? p : (p | (byte_of(mw, i) << 8));
        else if (nb_full) key = ((p == air_code) ? p : (p | (byte_of(mw, i) << 8))) | full;
        else key = node_key(pos, mat, c * 16 + i, p, air_code, interp, X, Y, Z);
        uint32_t id = 1;
        if (key != air_key) {
          if (key == last_key) id = last_id;
          else if (ASSIGN) {
            uint32_t slot = key_hash(key) % cap;
            id = 0;
            for (uint32_t probe = 0; probe < cap; probe++) {
              if (table[slot] == key) { id = ids[slot]; break; }
              slot = (slot + 1) % cap;
            }
          } else {
            class_set_insert(table, cap, count, key);
          }
          last_key = key; last_id = id;
        }
        if (ASSIGN) out[i >> 2] |= id << (8 * (i & 3));
      }
    }
    if (ASSIGN) reinterpret_cast<uint4*>(cls)[c] = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// one warp per 128-voxel row segment (4 class bytes per lane)
__global__ void count_dif_segments_kernel(const uint8_t* __restrict__ cls, int X, int Y, int nz, uint32_t dif_lo,
                                          uint32_t* __restrict__ counts, int keep) {
  const int segs = (X + 127) / 128;
  const int64_t n_seg = (int64_t)nz * Y * segs;
  const int lane = threadIdx.x & 31;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_seg; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int sx = (int)(w % segs);
    const int64_t row = w / segs;             // z*Y + y
    const int z = (int)(row / Y);
    uint32_t m = 0;                             // bit q: voxel q of this lane is a filter voxel
    if (z >= 1 && z < nz - 1) {
      const int x = sx * 128 + 4 * lane;
      for (int q = 0; q < 4; q++)
        if (x + q < X && cls[row * X + x + q] >= dif_lo) m |= 1u << q;
    }
    uint32_t c = __popc(m);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    const uint32_t lanes = __ballot_sync(0xffffffffu, m != 0u);
    const int first = lanes ? __ffs((int)lanes) - 1 : 0;
    const uint32_t mf = __shfl_sync(0xffffffffu, m, first);
    if (lanes) {   // x offset of the first filter voxel; bit 15: they form one contiguous run
      const int last = 31 - __clz((int)lanes);
      const uint32_t ml = __shfl_sync(0xffffffffu, m, last);
      const int xf = 4 * first + __ffs((int)mf) - 1, xl = 4 * last + (31 - __clz((int)ml));
      c |= (uint32_t)xf << 8;
      if ((uint32_t)(xl - xf + 1) == (c & 0xffu)) c |= 1u << 15;
    }
    // timing probes only (PFDTD_DEBUG_DIF_KEEP): 0 drops every filter voxel, 1 keeps single-voxel segments, 2 the others
    if (keep >= 0 && (keep == 0 || (keep == 1 && (c & 0xffu) != 1u) || (keep == 2 && (c & 0xffu) == 1u))) c = 0;
    if (lane == 0) counts[w] = c;
  }
}

int launch_count_dif_segments(const uint8_t* d_cls, int X, int Y, int nz, uint32_t dif_lo, uint32_t* d_counts, cudaStream_t stream) {
  const int64_t n_seg = (int64_t)nz * Y * ((X + 127) / 128);
  int blocks = (int)std::min<int64_t>((n_seg * 32 + 255) / 256, 148 * 16);
  if (blocks < 1) blocks = 1;
  const char* dbg = getenv("PFDTD_DEBUG_DIF_KEEP");
  count_dif_segments_kernel<<<blocks, 256, 0, stream>>>(d_cls, X, Y, nz, dif_lo, d_counts, dbg ? atoi(dbg) : -1);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

// ---- row-segment entries of the filter boundaries (update_math.cuh): filter voxels are numbered z-fastest within a
// (row, tile column) column, so the index of a segment's first voxel is an exclusive prefix over (column, z) ----
// one thread per column: number of filter voxels of the column over all planes
__global__ void dif_column_totals_kernel(const uint32_t* __restrict__ counts, int n_cols, int nz, uint32_t* __restrict__ totals) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  uint32_t t = 0;
  for (int z = 0; z < nz; z++) t += counts[(size_t)z * n_cols + c] & 0xffu;
  totals[c] = t;
}
// single block: exclusive scan of the column totals in place, grand total to *nb
__global__ void dif_scan_columns_kernel(uint32_t* __restrict__ totals, int n_cols, unsigned long long* __restrict__ nb) {
  __shared__ unsigned long long s_part[1024];
  const int t = threadIdx.x, per = (n_cols + blockDim.x - 1) / blockDim.x;
  const int lo = min(t * per, n_cols), hi = min(lo + per, n_cols);
  unsigned long long sum = 0;
  for (int i = lo; i < hi; i++) sum += totals[i];
  s_part[t] = sum;
  __syncthreads();
  if (t == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < (int)blockDim.x; i++) { const unsigned long long v = s_part[i]; s_part[i] = run; run += v; }
    *nb = run;
  }
  __syncthreads();
  unsigned long long run = s_part[t];
  for (int i = lo; i < hi; i++) { const uint32_t v = totals[i]; totals[i] = (uint32_t)run; run += v; }
}
// one thread per column walks the planes: entry = {index of the first filter voxel, DIF_HAS | DIF_SINGLE | DIF_RUN | count << 8 | x offset}
__global__ void dif_build_entries_kernel(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ col_base, int n_cols, int nz,
                                         uint2* __restrict__ entries) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  uint32_t run = col_base[c];
  for (int z = 0; z < nz; z++) {
    const size_t i = (size_t)z * n_cols + c;
    const uint32_t w = counts[i], n = w & 0xffu, xoff = (w >> 8) & 0x7fu;
    uint32_t fl = 0u;
    if (n == 1) fl = 0xC0000000u | xoff;
    else if (n > 1) fl = ((w >> 15) & 1u) ? (0xA0000000u | (n << 8) | xoff) : 0x80000000u;
    entries[i] = make_uint2(run, fl);
    run += n;
  }
}

int launch_build_dif_entries(const uint32_t* d_counts, int n_cols, int nz, uint32_t* d_col_scratch, unsigned long long* d_nb,
                             uint32_t* d_entries, cudaStream_t stream) {
  const int th = 128, bl = (n_cols + th - 1) / th;
  dif_column_totals_kernel<<<bl, th, 0, stream>>>(d_counts, n_cols, nz, d_col_scratch);
  dif_scan_columns_kernel<<<1, 1024, 0, stream>>>(d_col_scratch, n_cols, d_nb);
  dif_build_entries_kernel<<<bl, th, 0, stream>>>(d_counts, d_col_scratch, n_cols, nz, reinterpret_cast<uint2*>(d_entries));
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

static int class_blocks(uint64_t n) {
  int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  return blocks < 1 ? 1 : blocks;
}

int launch_mark_classes(const uint8_t* d_pos, const uint8_t* d_mat, uint64_t n, uint32_t air_key, uint32_t air_code, int interp,
                        uint32_t X, uint32_t Y, uint32_t Z, uint32_t* d_table, uint32_t cap, uint32_t* d_count, cudaStream_t stream) {
  if (X % 16 == 0)
    classes16_kernel<false><<<stream_blocks(n / 16), 256, 0, stream>>>(d_pos, d_mat, n, air_key, air_code, interp, X, Y, Z, d_table, nullptr, cap,
                                                                       d_count, nullptr);
  else
    mark_classes_kernel<<<class_blocks(n), 256, 0, stream>>>(d_pos, d_mat, n, air_key, air_code, interp, X, Y, Z, d_table, cap, d_count);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

int launch_assign_classes(const uint8_t* d_pos, const uint8_t* d_mat, uint64_t n, uint32_t air_key, uint32_t air_code, int interp,
                          uint32_t X, uint32_t Y, uint32_t Z, const uint32_t* d_table, const uint8_t* d_ids, uint32_t cap,
                          uint8_t* d_cls, cudaStream_t stream) {
  if (X % 16 == 0)
    classes16_kernel<true><<<stream_blocks(n / 16), 256, 0, stream>>>(d_pos, d_mat, n, air_key, air_code, interp, X, Y, Z,
                                                                      const_cast<uint32_t*>(d_table), d_ids, cap, nullptr, d_cls);
  else
    assign_classes_kernel<<<class_blocks(n), 256, 0, stream>>>(d_pos, d_mat, n, air_key, air_code, interp, X, Y, Z, d_table, d_ids, cap, d_cls);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

}  // namespace pfdtd
