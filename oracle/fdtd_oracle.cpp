// =============================================================================
// oracle/fdtd_oracle.cpp -- CPU restatement of ParallelFDTD's time-stepping path.
//
// TEST INFRASTRUCTURE ONLY. Nothing under parallelfdtd_b200/ may include, link or
// call this file; it is the checker for tests/, __graft_entry__.smoke() and the
// `cpu_baseline` / `--impl reference --ref-kind cpu` legs of bench.py.
//
// Parity status: PINNED. The arithmetic below reproduces, operation by operation
// (including the FMA contraction nvcc 12.9 applies to the reference source for
// sm_100 -- read off `cuobjdump -sass` of oracle/_ref/kernels3d.o), what the
// reference's own CUDA kernels compute; tests/golden/*.npz hold responses and
// node bytes produced by the reference itself (oracle/_ref/ref_fdtd run on a
// B200) and tests/test_oracle_golden.py checks this file against them bit for bit.
// PARITY UNPINNED for two parts, which the reference does not contain: the interpolated
// 27-point schemes (run_interp) and the digital-impedance-filter boundaries (run_dif).  They restate the literature
// and are anchored on the pinned path (SRL weights / order-0 filters reproduce it bit for bit), see DESIGN.md section 2.
//
// Every function cites the reference file:line (relative to the reference root)
// it restates. 64-bit indexing throughout (the reference is 32-bit, SURVEY C-11).
// Build: see oracle/Makefile (-ffp-contract=off: contraction is written explicitly).
// =============================================================================
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// fma helpers (explicit contraction, see header)
inline float  FMA(float a, float b, float c)    { return fmaf(a, b, c); }
inline double FMA(double a, double b, double c) { return fma(a, b, c); }

struct Slab {
  int64_t first;   // first global z slice held
  int64_t size;    // number of slices held (incl. halos)
};

// src/kernels/cudaMesh.h:280-307  CudaMesh::getPartitionIndexing
std::vector<Slab> partition_indexing(int64_t dim, int num_parts) {
  std::vector<Slab> ret;
  int64_t part_size = dim / num_parts;
  for (int i = 0; i < num_parts; i++) {
    int64_t s_inc = (i == 0) ? 0 : 1;
    int64_t e_inc = (i == num_parts - 1) ? 0 : 1;
    int64_t cur = part_size + s_inc + e_inc;
    if (i != 0 && i == num_parts - 1) cur += dim - (int64_t)(i + 1) * part_size;
    ret.push_back({(int64_t)i * part_size - s_inc, cur});
  }
  return ret;
}

// src/kernels/kernels3d.cu:485-529  fdtd3dStdMaterials<T>  (SRL_FORWARD)
// One voxel. `matidx_mode`: 1 = as written (`mat*20*+octave`, kernels3d.cu:513),
// 0 = intended (`mat*20+octave`).
template <typename T>
inline T update_forward(uint8_t pos, uint8_t mat, const T* P, int64_t cur, int64_t dimx, int64_t dimxy,
                        T p_old, const T* params, const T* materials, int matidx_mode) {
  T position = (T)(pos & 0x7F);
  T sw = (T)(pos >> 7);
  unsigned mat_idx;
  if (matidx_mode == 1) mat_idx = (unsigned)((T)((unsigned)mat * 20u) * params[3]);
  else                  mat_idx = (unsigned)((unsigned)mat * 20u) + (unsigned)params[3];
  T t = (materials[mat_idx] * ((T)6 - position)) * params[0];
  T one_p_beta = FMA(t, (T)0.5, (T)1);     // 1 + beta,  beta = 0.5*t (exact)
  T one_m_beta = FMA(t, (T)-0.5, (T)1);    // 1 - beta
  T S = P[cur + dimxy] + P[cur - dimxy];
  S = S + P[cur + dimx];
  S = S + P[cur - dimx];
  S = S + P[cur + 1];
  S = S + P[cur - 1];
  T p = P[cur];
  T a = FMA(position, -params[1], (T)2);   // 2 - K*lambda^2 (contracted)
  T inner = FMA(S, params[1], p * a);
  inner = FMA(p_old, -one_m_beta, inner);
  T rcp = (T)1 / one_p_beta;
  return (sw * rcp) * inner;
}

// src/kernels/kernels3d.cu:608-665  fdtd3dStdKowalczykMaterials<T>  (SRL, centred)
template <typename T>
inline T update_centred(uint8_t pos, uint8_t mat, const T* P, int64_t cur, int64_t dimx, int64_t dimxy,
                        T p_old, const T* params, const T* materials) {
  T sw = (T)(pos >> 7);
  T dir_x = (T)(pos & 0x01);
  T dir_y = (T)((pos & 0x02) >> 1);
  T dir_z = (T)((pos & 0x04) >> 2);
  unsigned mat_idx = (unsigned)((T)((unsigned)mat * 20u) + params[3]);
  T dsum = (dir_x + dir_y) + dir_z;        // small integers: exact in any order
  T cl = materials[mat_idx] * params[0];
  T one_p_beta = FMA(cl, dsum, (T)1);
  T beta_m_one = FMA(cl, dsum, (T)-1);
  T p_z[2], p_y[2], p_x[2];
  p_z[1] = P[cur - dimxy];  // SIGN_Z is "down": inverted (kernels3d.cu:643-644)
  p_z[0] = P[cur + dimxy];
  p_y[0] = P[cur - dimx];
  p_y[1] = P[cur + dimx];
  p_x[0] = P[cur - 1];
  p_x[1] = P[cur + 1];
  T p = P[cur];
  int sign_x = (pos & 0x10) >> 4, sign_y = (pos & 0x20) >> 5, sign_z = (pos & 0x40) >> 6;
  // dir_* are 0/1 so every product is exact; grouping follows the source
  T S_b = (p_x[sign_x] * dir_x + p_y[sign_y] * dir_y) + p_z[sign_z] * dir_z;
  T S = p_x[0] + p_x[1];
  S = S + p_y[0];
  S = S + p_y[1];
  S = S + p_z[0];
  S = S + p_z[1];
  S = S + S_b;
  T c = FMA(params[1], (T)-6, (T)2);       // 2 - 6*lambda^2 (contracted)
  T q = p * c;
  T inner = FMA(S, params[1], -q);
  inner = FMA(p_old, beta_m_one, inner);
  T rcp = (T)1 / one_p_beta;
  return (sw * inner) * rcp;
}

// Interpolated 27-point compact schemes (IISO / IWB).  NOT in the mounted reference:
// "parity unpinned" -- this restates the builder-defined equation documented in
// parallelfdtd_b200/csrc/update_math.cuh ("interpolated"), operation by operation:
//   p_new = sw/(1+beta) * ( d1*A6 + d2*A12 + d3*A8 + c0*p - (1-beta)*p_old ),
//   c0 = d4 + d1*(6-K6) + d2*(12-K12) + d3*(8-K8),  beta = 0.5*Y*(6-K6)*lam,
// which for d = (lam2, 0, 0, 2-6*lam2) is the reference's SRL_FORWARD update (kernels3d.cu:508-528).
// Values outside the xy extent read as 0; K12/K8 count the non-solid edge/corner neighbours.
template <typename T>
inline T update_interp(const uint8_t* pos, const uint8_t* mat, const T* P, int64_t x, int64_t y, int64_t z, int64_t X, int64_t Y,
                       T p_old, const T* params, const T* materials, const T* d) {
  const int64_t XY = X * Y;
  auto at = [&](int64_t xx, int64_t yy, int64_t zz) -> T {
    if (xx < 0 || yy < 0 || xx >= X || yy >= Y) return (T)0;
    return P[zz * XY + yy * X + xx];
  };
  auto inside = [&](int64_t xx, int64_t yy, int64_t zz) -> unsigned {
    if (xx < 0 || yy < 0 || xx >= X || yy >= Y) return 0u;
    return (unsigned)(pos[zz * XY + yy * X + xx] >> 7);
  };
  T c[3], a4[3], g4[3];
  for (int k = 0; k < 3; k++) {
    const int64_t zz = z - 1 + k;
    c[k] = at(x, y, zz);
    a4[k] = ((at(x - 1, y, zz) + at(x + 1, y, zz)) + at(x, y - 1, zz)) + at(x, y + 1, zz);
    g4[k] = ((at(x - 1, y - 1, zz) + at(x + 1, y - 1, zz)) + at(x - 1, y + 1, zz)) + at(x + 1, y + 1, zz);
  }
  const uint8_t pb = pos[z * XY + y * X + x];
  unsigned k12 = 0, k8 = 0;
  for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
    const int nz_ = (dx != 0) + (dy != 0) + (dz != 0);
    if (nz_ == 2) k12 += inside(x + dx, y + dy, z + dz);
    if (nz_ == 3) k8 += inside(x + dx, y + dy, z + dz);
  }
  const T K = (T)(pb & 0x7F);
  const T sw = (T)(pb >> 7);
  const unsigned m = pb == 0 ? 0u : (unsigned)mat[z * XY + y * X + x];
  const unsigned idx = m * 20u + (unsigned)params[3];
  const T t = (materials[idx] * ((T)6 - K)) * params[0];
  T c0 = FMA(d[0], (T)6 - K, d[3]);
  c0 = FMA(d[1], (T)(12u - k12), c0);
  c0 = FMA(d[2], (T)(8u - k8), c0);
  const T c1 = -FMA(t, (T)-0.5, (T)1);
  const T c2 = sw * ((T)1 / FMA(t, (T)0.5, (T)1));
  const T A6 = (a4[1] + c[0]) + c[2];
  const T A12 = (g4[1] + a4[0]) + a4[2];
  T inner = FMA(A6, d[0], c[1] * c0);
  inner = FMA(A12, d[1], inner);
  if (d[2] != (T)0) inner = FMA(g4[0] + g4[2], d[2], inner);
  inner = FMA(p_old, c1, inner);
  return c2 * inner;
}

// Digital impedance filter boundary (frequency-dependent admittance).  NOT in the reference (its dif_ members
// are a stub, cudaMesh.h:88,122-123): "parity unpinned"; restates csrc/update_math.cuh "digital impedance filters".
// Material row = [b0..bN, a1..aN]; val0 = the frequency-independent update evaluated with Y = b0; st = the voxel's
// N transposed-direct-form-II states (stride `stride`).  Returns p_new and advances the states.
template <typename T>
inline T dif_voxel(uint8_t pos, uint8_t mat, int scheme, T val0, T p_old, const T* params, const T* materials, int order, T* st,
                   int64_t stride) {
  const T* row = materials + (size_t)(pos == 0 ? 0 : mat) * 20;
  const T b0 = row[0];
  const T sw = (T)(pos >> 7);
  T kap, rc;
  if (scheme == 2) {
    const T dsum = (T)((pos & 1) + ((pos >> 1) & 1) + ((pos >> 2) & 1));
    kap = params[0] * dsum;
    rc = sw * ((T)1 / FMA(b0 * params[0], dsum, (T)1));
  } else {
    const T K = (T)(pos & 0x7F);
    kap = ((T)0.5 * ((T)6 - K)) * params[0];
    const T t = (b0 * ((T)6 - K)) * params[0];
    rc = sw * ((T)1 / FMA(t, (T)0.5, (T)1));
  }
  const T c3 = kap * rc;
  T s[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < order; i++) s[i] = st[i * stride];
  const T p_new = FMA(-c3, s[0], val0);
  const T u = p_new + (-p_old);
  const T y = FMA(b0, u, s[0]);
  for (int i = 0; i < order; i++) st[i * stride] = FMA(row[1 + i], u, FMA(-row[1 + order + i], y, s[i + 1]));
  return p_new;
}
inline bool is_lossy(uint8_t pos, int scheme) { return scheme == 2 ? (pos & 7) != 0 : (pos != 0 && (pos & 0x7F) < 6); }

// One launch of the update kernel over a slab: local slices 1..nz-2 of `Q` (the past field) are
// overwritten with the next field (kernels3d.cu:109-153: grid.z = slab slices - 2, pointers offset
// by one slice).  pos/mat point at the slab's first slice.
template <typename T>
void update_slab(const uint8_t* pos, const uint8_t* mat, int64_t X, int64_t Y, int64_t nz, int scheme, const T* params,
                 const T* materials, int matidx_mode, const T* P, T* Q, int dif_order = 0, T* dif_state = nullptr) {
  const int64_t XY = X * Y;
  const int64_t nvox = nz * XY;
  if (dif_order > 0) {
    // filters on: the scalar admittance of material m is its b0 = materials[m*20] (octave ignored)
    T prm[8];
    for (int i = 0; i < 8; i++) prm[i] = (scheme >= 3 || i < 4) ? params[i] : (T)0;
    prm[3] = (T)0;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t z = 1; z < nz - 1; z++)
      for (int64_t y = 0; y < Y; y++)
        for (int64_t x = 0; x < X; x++) {
          const int64_t e = z * XY + y * X + x;
          const T p_old = Q[e];
          T v;
          if (scheme >= 3) v = update_interp<T>(pos, mat, P, x, y, z, X, Y, p_old, prm, materials, params + 4);
          else if (scheme == 2) v = update_centred<T>(pos[e], mat[e], P, e, X, XY, p_old, prm, materials);
          else v = update_forward<T>(pos[e], mat[e], P, e, X, XY, p_old, prm, materials, 0);
          if (is_lossy(pos[e], scheme)) v = dif_voxel<T>(pos[e], mat[e], scheme, v, p_old, prm, materials, dif_order, dif_state + e, nvox);
          Q[e] = v;
        }
    return;
  }
  if (scheme >= 3) {   // interpolated: params[4..7] = d1..d4
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t z = 1; z < nz - 1; z++)
      for (int64_t y = 0; y < Y; y++)
        for (int64_t x = 0; x < X; x++) {
          const int64_t e = z * XY + y * X + x;
          Q[e] = update_interp<T>(pos, mat, P, x, y, z, X, Y, Q[e], params, materials, params + 4);
        }
    return;
  }
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t z = 1; z < nz - 1; z++) {
    for (int64_t y = 0; y < Y; y++) {
      const int64_t row = z * XY + y * X;
      const uint8_t* prow = pos + row;
      const uint8_t* mrow = mat + row;
      if (scheme == 2) {
        for (int64_t x = 0; x < X; x++)
          Q[row + x] = update_centred<T>(prow[x], mrow[x], P, row + x, X, XY, Q[row + x], params, materials);
      } else {
        for (int64_t x = 0; x < X; x++)
          Q[row + x] = update_forward<T>(prow[x], mrow[x], P, row + x, X, XY, Q[row + x], params, materials, matidx_mode);
      }
    }
  }
}

// src/kernels/kernels3d.cu:31-203 / 205-374 : launchFDTD3d / launchFDTD3dDouble.
// Emulates N partitions with halo slices exactly as CudaMesh holds them
// (cudaMesh.h:648-751), setSample/addSample semantics (cudaMesh.h:321-369),
// flipPressurePointers (:766-779), switchHalos (:432-463), receiver = first slab
// containing z (:251-266).
template <typename T>
double run_sim(const uint8_t* pos, const uint8_t* mat, int64_t X, int64_t Y, int64_t Z, int scheme,
               const T* params, const T* materials, int matidx_mode, int soft_mode, int n_parts,
               int n_src, const int32_t* src_xyz, const int32_t* src_type, const T* src_samples,
               int n_rec, const int32_t* rec_xyz, int64_t steps, T* out, int timed_from_step, int dif_order = 0) {
  const int64_t XY = X * Y;
  std::vector<Slab> slabs = partition_indexing(Z, n_parts);
  std::vector<std::vector<T>> Pc(n_parts), Pp(n_parts), Ds(n_parts);
  if (dif_order > 0)
    for (int k = 0; k < n_parts; k++) Ds[k].assign((size_t)dif_order * slabs[k].size * XY, (T)0);
  for (int k = 0; k < n_parts; k++) {
    // +1+dimX slack like the reference (cudaMesh.h:735)
    Pc[k].assign(slabs[k].size * XY + 1 + X, (T)0);
    Pp[k].assign(slabs[k].size * XY + 1 + X, (T)0);
  }
  std::vector<T*> cur(n_parts), past(n_parts);
  for (int k = 0; k < n_parts; k++) { cur[k] = Pc[k].data(); past[k] = Pp[k].data(); }

  auto t0 = std::chrono::steady_clock::now();
  for (int64_t step = 0; step < steps; step++) {
    if (step == timed_from_step) t0 = std::chrono::steady_clock::now();
    // (1) sources (kernels3d.cu:93-104)
    for (int s = 0; s < n_src; s++) {
      T v = src_samples[(int64_t)s * steps + step];
      int64_t x = src_xyz[3 * s], y = src_xyz[3 * s + 1], z = src_xyz[3 * s + 2];
      for (int k = 0; k < n_parts; k++) {
        if (z > slabs[k].first + slabs[k].size - 1) continue;
        if (z < slabs[k].first) break;
        T* dst = cur[k] + (z - slabs[k].first) * XY + y * X + x;
        if (src_type[s] == 0 || soft_mode == 0) *dst = v;  // HARD, or addSample as written (cudaMesh.h:362-366)
        else *dst += v;                                     // true accumulate (soft_mode 1)
      }
    }
    // (2) update local slices 1..size-2 of every slab (kernels3d.cu:109-153)
    for (int k = 0; k < n_parts; k++)
      update_slab<T>(pos + slabs[k].first * XY, mat + slabs[k].first * XY, X, Y, slabs[k].size, scheme, params, materials,
                     matidx_mode, cur[k], past[k], dif_order, dif_order > 0 ? Ds[k].data() : nullptr);
    // (3) flip (kernels3d.cu:160)
    for (int k = 0; k < n_parts; k++) std::swap(cur[k], past[k]);
    // (4) halos (kernels3d.cu:161, cudaMesh.h:432-463)
    for (int k = 0; k + 1 < n_parts; k++) {
      int64_t nzk = slabs[k].size;
      memcpy(cur[k + 1], cur[k] + (nzk - 2) * XY, XY * sizeof(T));          // slab_k[last-1] -> slab_{k+1}[0]
      memcpy(cur[k] + (nzk - 1) * XY, cur[k + 1] + XY, XY * sizeof(T));      // slab_{k+1}[1]  -> slab_k[last]
    }
    // (5) receivers (kernels3d.cu:164-173)
    for (int r = 0; r < n_rec; r++) {
      int64_t x = rec_xyz[3 * r], y = rec_xyz[3 * r + 1], z = rec_xyz[3 * r + 2];
      for (int k = 0; k < n_parts; k++) {
        if (z > slabs[k].first + slabs[k].size - 1) continue;
        if (z < slabs[k].first) break;
        out[(int64_t)r * steps + step] = cur[k][(z - slabs[k].first) * XY + y * X + x];
        break;
      }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace

extern "C" {

int pfo_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// cudaMesh.h:280-307. out_first/out_size: [num_parts]
void pfo_partition_indexing(int64_t dim, int num_parts, int64_t* out_first, int64_t* out_size) {
  auto s = partition_indexing(dim, num_parts);
  for (int i = 0; i < num_parts; i++) { out_first[i] = s[i].first; out_size[i] = s[i].size; }
}

// cudaMesh.cu:253-304 padded dims.
void pfo_padded_dims(uint32_t dx, uint32_t dy, uint32_t dz, uint32_t bx, uint32_t by, uint32_t bz, uint32_t* out3) {
  uint32_t px = 0, py = 0, pz = 0;
  if (dx % bx) px = bx - dx % bx;
  if (dy % by) py = by - dy % by;
  if (dz % bz) pz = bz - dz % bz;
  out3[0] = dx + px; out3[1] = dy + py; out3[2] = dz + pz;
}

// cudaMesh.cu:253-304 padWithZeros + :516-534 padWithZerosKernel.
// Copies only x,y,z >= 1; x may equal dim_x (and y dim_y) when that axis was padded, in which
// case the flat old index wraps into the next row / slice (SURVEY C-8). The reference reads past
// the end of the old volume for (y==dim_y, z==dim_z-1); that undefined read is defined as 0 here.
void pfo_pad_with_zeros(const uint8_t* old_v, uint32_t dx, uint32_t dy, uint32_t dz, uint32_t bx, uint32_t by,
                        uint32_t bz, uint8_t* new_v) {
  uint32_t nd[3];
  pfo_padded_dims(dx, dy, dz, bx, by, bz, nd);
  const uint64_t NX = nd[0], NY = nd[1], NZ = nd[2];
  memset(new_v, 0, NX * NY * NZ);
  const uint64_t old_n = (uint64_t)dx * dy * dz;
  for (uint64_t z = 1; z < dz; z++)            // grid.z = dim_z  => z <= dim_z-1
    for (uint64_t y = 1; y < NY && y <= dy; y++)
      for (uint64_t x = 1; x < NX && x <= dx; x++) {
        uint64_t oi = z * dx * dy + y * dx + x;
        new_v[z * NX * NY + y * NX + x] = oi < old_n ? old_v[oi] : 0;
      }
}

// cudaMesh.cu:328-361 toBilbao (in place on padded volumes) + :500-514 calcBoundaries with
// air value 0x86 (:228-230). counts[0]=air, counts[1]=boundary.
void pfo_to_bilbao(uint8_t* pos, uint8_t* mat, uint64_t n, uint64_t* counts) {
  uint64_t air = 0, bnd = 0;
  for (uint64_t i = 0; i < n; i++) {
    unsigned k = pos[i];
    if (k == 0) { pos[i] = 0; mat[i] = 0; }
    else if (k <= 8) pos[i] = 0x80 | 3;
    else if (k <= 20) pos[i] = 0x80 | 4;
    else if (k <= 26) pos[i] = 0x80 | 5;
    else if (k == 27) pos[i] = 0x80 | 6;
    if (pos[i] == 0x86) air++;
    if (pos[i] != 0 && pos[i] != 0x86) bnd++;
  }
  counts[0] = air; counts[1] = bnd;
}

// cudaMesh.cu:363-480 toKowalczyk + calcBoundaries with air value 0x80 (:180-182).
void pfo_to_kowalczyk(uint8_t* pos, uint8_t* mat, uint64_t n, uint64_t* counts) {
  enum { DX = 0x01, DY = 0x02, DZ = 0x04, SX = 0x10, SY = 0x20, SZ = 0x40, C = 0x80 };
  static const uint8_t lut[28] = {
      0,
      SZ | DX | DY | DZ | C, SZ | SX | DX | DY | DZ | C, SZ | SY | DX | DY | DZ | C, SZ | SY | SX | DX | DY | DZ | C,
      DX | DY | DZ | C, SX | DX | DY | DZ | C, SY | DX | DY | DZ | C, SY | SX | DX | DY | DZ | C,
      DY | DZ | SZ | C, DY | DZ | SZ | SY | C, DX | DZ | SZ | C, DX | DZ | SZ | SX | C,
      DY | DZ | C, DY | DZ | SY | C, DX | DZ | C, DX | DZ | SX | C,
      DY | DX | C, DX | DY | SX | C, DX | DY | SY | C, DY | DX | SY | SX | C,
      DZ | SZ | C, DY | SY | C, DY | C, DX | SX | C, DX | C, DZ | C,
      C};
  uint64_t air = 0, bnd = 0;
  for (uint64_t i = 0; i < n; i++) {
    unsigned k = pos[i];
    if (k == 0) { pos[i] = 0; mat[i] = 0; }
    else if (k <= 27) pos[i] = lut[k];
    if (pos[i] == 0x80) air++;
    if (pos[i] != 0 && pos[i] != 0x80) bnd++;
  }
  counts[0] = air; counts[1] = bnd;
}

// Full run; pos/mat are the padded, scheme-translated GLOBAL volumes [Z][Y][X].
// scheme: 0 SRL_FORWARD, 1 SHARED (mapped to the forward equations, SURVEY C-3), 2 SRL (centred),
// 3 interpolated 27-point family (params then has 8 entries: [lam, lam2, 1/3, octave, d1, d2, d3, d4]).
// src_samples [n_src][steps], out [n_rec][steps]. Returns wall seconds spent in steps >= timed_from_step.
double pfo_run_f32(const uint8_t* pos, const uint8_t* mat, int64_t X, int64_t Y, int64_t Z, int scheme,
                   const float* params, const float* materials, int matidx_mode, int soft_mode, int n_parts,
                   int n_src, const int32_t* src_xyz, const int32_t* src_type, const float* src_samples, int n_rec,
                   const int32_t* rec_xyz, int64_t steps, float* out, int timed_from_step) {
  return run_sim<float>(pos, mat, X, Y, Z, scheme, params, materials, matidx_mode, soft_mode, n_parts, n_src, src_xyz,
                        src_type, src_samples, n_rec, rec_xyz, steps, out, timed_from_step);
}
// same with digital-impedance-filter boundaries of order dif_order (material rows = [b0..bN, a1..aN])
double pfo_run_dif_f32(const uint8_t* pos, const uint8_t* mat, int64_t X, int64_t Y, int64_t Z, int scheme,
                       const float* params, const float* materials, int dif_order, int n_parts, int n_src, const int32_t* src_xyz,
                       const int32_t* src_type, const float* src_samples, int n_rec, const int32_t* rec_xyz, int64_t steps, float* out) {
  return run_sim<float>(pos, mat, X, Y, Z, scheme, params, materials, 0, 0, n_parts, n_src, src_xyz, src_type, src_samples, n_rec,
                        rec_xyz, steps, out, 0, dif_order);
}
double pfo_run_dif_f64(const uint8_t* pos, const uint8_t* mat, int64_t X, int64_t Y, int64_t Z, int scheme,
                       const double* params, const double* materials, int dif_order, int n_parts, int n_src, const int32_t* src_xyz,
                       const int32_t* src_type, const double* src_samples, int n_rec, const int32_t* rec_xyz, int64_t steps, double* out) {
  return run_sim<double>(pos, mat, X, Y, Z, scheme, params, materials, 0, 0, n_parts, n_src, src_xyz, src_type, src_samples, n_rec,
                         rec_xyz, steps, out, 0, dif_order);
}
double pfo_run_f64(const uint8_t* pos, const uint8_t* mat, int64_t X, int64_t Y, int64_t Z, int scheme,
                   const double* params, const double* materials, int matidx_mode, int soft_mode, int n_parts,
                   int n_src, const int32_t* src_xyz, const int32_t* src_type, const double* src_samples, int n_rec,
                   const int32_t* rec_xyz, int64_t steps, double* out, int timed_from_step) {
  return run_sim<double>(pos, mat, X, Y, Z, scheme, params, materials, matidx_mode, soft_mode, n_parts, n_src, src_xyz,
                         src_type, src_samples, n_rec, rec_xyz, steps, out, timed_from_step);
}

// One update launch on a single slab whose end planes are halos (multi-process emulation in
// tests/test_slabs_gloo.py).  cur/past: [nz][Y][X] with X+1 elements of slack NOT required: the x=0 /
// x=X-1 taps that wrap rows stay inside the slab because planes 0 and nz-1 are never updated.
void pfo_step_slab_f32(const uint8_t* pos, const uint8_t* mat, int64_t X, int64_t Y, int64_t nz, int scheme, const float* params,
                       const float* materials, int matidx_mode, const float* cur, float* past) {
  update_slab<float>(pos, mat, X, Y, nz, scheme, params, materials, matidx_mode, cur, past);
}
void pfo_step_slab_f64(const uint8_t* pos, const uint8_t* mat, int64_t X, int64_t Y, int64_t nz, int scheme, const double* params,
                       const double* materials, int matidx_mode, const double* cur, double* past) {
  update_slab<double>(pos, mat, X, Y, nz, scheme, params, materials, matidx_mode, cur, past);
}

// ---- host-side pieces of the path -------------------------------------------
// SimulationParameters.cpp:379-396 getParameterPtr[Double]
void pfo_params_f32(double lambda, unsigned octave, float* out4) {
  out4[0] = (float)lambda; out4[1] = (float)(lambda * lambda); out4[2] = 1.f / 3.f; out4[3] = (float)octave;
}
void pfo_params_f64(double lambda, unsigned octave, double* out4) {
  out4[0] = lambda; out4[1] = lambda * lambda; out4[2] = (double)1 / (double)3; out4[3] = (double)octave;
}
// SimulationParameters.cpp:362-364 getDx
float pfo_dx(float c, unsigned fs, double lambda) { return (float)((double)c / ((double)fs * lambda)); }

// SrcRec.cpp:26-34 Position::getElementIdx + SimulationParameters.cpp:200-224 (+1 padding)
void pfo_element_idx(float px, float py, float pz, unsigned fs, float c, float lambda, int add_padding, int32_t* out3) {
  float dx = c / ((float)fs * lambda);
  out3[0] = (int)floorf(px / dx + 0.5f) + (add_padding ? 1 : 0);
  out3[1] = (int)floorf(py / dx + 0.5f) + (add_padding ? 1 : 0);
  out3[2] = (int)floorf(pz / dx + 0.5f) + (add_padding ? 1 : 0);
}

// SimulationParameters.cpp:262-298 getRegularSourceSample (float).  input_type: 0 IMPULSE 1 GAUSSIAN 2 SINE 3 DATA
float pfo_regular_sample_f32(int input_type, unsigned step, unsigned fs, const float* data, unsigned n_data) {
  float sample = 0.f;
  switch (input_type) {
    case 0: sample += (step == 1) ? 1.f : 0.f; break;
    case 1: { float t0 = 40, width = 4; float e = ((float)(step - t0) / width); sample += expf(-0.5f * (e * e)); break; }
    case 3: sample += (step < n_data) ? data[step] : 0.f; break;
    case 2: { float freq = 120; float t = (float)step / (float)fs; sample += sinf(2.f * (float)3.14159265358979323846 * freq * t); break; }
  }
  return sample;
}
// SimulationParameters.cpp:312-348 (double)
double pfo_regular_sample_f64(int input_type, unsigned step, unsigned fs, const double* data, unsigned n_data) {
  double sample = 0.0;
  switch (input_type) {
    case 0: sample += (step == 1) ? 1.0 : 0.0; break;
    case 1: { double t0 = 40, width = 4; double e = ((float)(step - t0) / width); sample += exp(-0.5f * (e * e)); break; }
    case 3: sample += (step < n_data) ? data[step] : 0.0; break;
    case 2: { double freq = 120; double t = (double)step / (double)fs; sample += sin(2.f * (double)3.14159265358979323846 * freq * t); break; }
  }
  return sample;
}
// SimulationParameters.cpp:300-310 getTransparentSourceSample: regular(step) - sum_{i<step} ir[step-i]*regular(i)
float pfo_transparent_sample_f32(int input_type, unsigned step, unsigned fs, const float* data, unsigned n_data,
                                 const float* grid_ir, unsigned n_ir) {
  float s = 0.f;
  for (unsigned i = 0; i < step; i++) {
    float ir = (step - i) < n_ir ? grid_ir[step - i] : 0.f;
    s += ir * pfo_regular_sample_f32(input_type, i, fs, data, n_data);
  }
  return pfo_regular_sample_f32(input_type, step, fs, data, n_data) - s;
}
double pfo_transparent_sample_f64(int input_type, unsigned step, unsigned fs, const double* data, unsigned n_data,
                                  const float* grid_ir, unsigned n_ir) {
  double s = 0.f;
  for (unsigned i = 0; i < step; i++) {
    float ir = (step - i) < n_ir ? grid_ir[step - i] : 0.f;
    s += ir * pfo_regular_sample_f64(input_type, i, fs, data, n_data);
  }
  return pfo_regular_sample_f64(input_type, step, fs, data, n_data) - s;
}

}  // extern "C"
