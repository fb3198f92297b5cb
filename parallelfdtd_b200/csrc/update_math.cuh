// update_math.cuh -- per-voxel update arithmetic shared by every update kernel.
//
// One device function per scheme, used by the plain kernel, the TMA z-march kernel, the edge
// launches and the interior launches alike, so a voxel's result never depends on which kernel or
// how many partitions computed it (reference invariant: tests/CudaMeshTest.cpp:508-518).
//
// The operation order and the FMA contraction reproduce what nvcc emits for the reference
// kernels (src/kernels/kernels3d.cu:485-529 and :608-665) when built for sm_100 -- every op is an
// explicit round-to-nearest intrinsic, so the compiler cannot re-associate or re-contract.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pfdtd {

template <typename T> struct Ar;
template <> struct Ar<float> {
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float rcp(float a) { return __frcp_rn(a); }   // == 1.f/a (IEEE)
};
template <> struct Ar<double> {
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double rcp(double a) { return __drcp_rn(a); }  // == 1.0/a (IEEE)
};


// Per-launch constants. params = [lambda, lambda^2, 1/3, octave] (SimulationParameters.cpp:379-396).
template <typename T>
struct UpdConst {
  T lam;          // params[0]
  T lam2;         // params[1]
  T octave;       // params[3]
  T a_air;        // forward:  fma(6, -lam2, 2)     centred: fma(lam2, -6, 2)
  T d[4];         // interpolated schemes: d1 axial, d2 edge, d3 corner, d4 centre weights
  const T* materials;      // [n_unique][20] admittances on this device
  uint32_t n_coefs;        // entries in `materials` (index is clamped; the reference would read out of bounds)
  int matidx_as_written;   // forward kernel only, kernels3d.cu:513
  int dif_order;           // > 0: material rows hold filter coefficients [b0..bN, a1..aN]; the scalar admittance is b0
};

template <typename T>
__device__ __forceinline__ T load_coef(const UpdConst<T>& c, uint32_t idx) {
  idx = idx < c.n_coefs ? idx : c.n_coefs - 1;
  return __ldg(c.materials + idx);
}

// ---- SRL_FORWARD (kernels3d.cu:508-528) --------------------------------------------------------
// S is summed as ((((z+ + z-) + y+) + y-) + x+) + x-.
template <typename T>
__device__ __forceinline__ T sum6_forward(T zp, T zm, T yp, T ym, T xp, T xm) {
  T S = Ar<T>::add(zp, zm);
  S = Ar<T>::add(S, yp);
  S = Ar<T>::add(S, ym);
  S = Ar<T>::add(S, xp);
  S = Ar<T>::add(S, xm);
  return S;
}

template <typename T>
__device__ __forceinline__ T voxel_forward(uint32_t pos, const uint8_t* mat_ptr, T p, T S, T p_old, const UpdConst<T>& c) {
  if (pos == 0x86u) {  // air: K=6, beta=0 -> identical bits to the general expression
    T inner = Ar<T>::fma(S, c.lam2, Ar<T>::mul(p, c.a_air));
    return Ar<T>::fma(p_old, (T)-1, inner);
  }
  T K = (T)(pos & 0x7Fu);
  T sw = (T)(pos >> 7);
  uint32_t m = (pos != 0u) ? (uint32_t)(*mat_ptr) : 0u;   // solid nodes always carry material 0 (cudaMesh.cu:332-336)
  uint32_t idx = c.matidx_as_written ? (uint32_t)Ar<T>::mul((T)(m * 20u), c.octave)
                                     : m * 20u + (uint32_t)c.octave;
  T coef = load_coef(c, idx);
  T t = Ar<T>::mul(Ar<T>::mul(coef, Ar<T>::add((T)6, -K)), c.lam);
  T one_p_beta = Ar<T>::fma(t, (T)0.5, (T)1);
  T one_m_beta = Ar<T>::fma(t, (T)-0.5, (T)1);
  T a = Ar<T>::fma(K, -c.lam2, (T)2);
  T inner = Ar<T>::fma(S, c.lam2, Ar<T>::mul(p, a));
  inner = Ar<T>::fma(p_old, -one_m_beta, inner);
  return Ar<T>::mul(Ar<T>::mul(sw, Ar<T>::rcp(one_p_beta)), inner);
}

// ---- SRL centred (kernels3d.cu:632-663) ----------------------------------------------------------
// sum = (((((x- + x+) + y-) + y+) + z+) + z-) + S_boundary
template <typename T>
__device__ __forceinline__ T voxel_centred(uint32_t pos, const uint8_t* mat_ptr, T p, T zp, T zm, T yp, T ym, T xp, T xm,
                                           T p_old, const UpdConst<T>& c) {
  T S = Ar<T>::add(xm, xp);
  S = Ar<T>::add(S, ym);
  S = Ar<T>::add(S, yp);
  S = Ar<T>::add(S, zp);
  S = Ar<T>::add(S, zm);
  T q = Ar<T>::mul(p, c.a_air);
  if (pos == 0x80u) {  // air: no direction flags, beta = 0
    S = Ar<T>::add(S, (T)0);
    T inner = Ar<T>::fma(S, c.lam2, -q);
    return Ar<T>::fma(p_old, (T)-1, inner);
  }
  T sw = (T)(pos >> 7);
  T dir_x = (T)(pos & 1u), dir_y = (T)((pos >> 1) & 1u), dir_z = (T)((pos >> 2) & 1u);
  T dsum = (T)((pos & 1u) + ((pos >> 1) & 1u) + ((pos >> 2) & 1u));
  uint32_t m = (pos != 0u) ? (uint32_t)(*mat_ptr) : 0u;
  uint32_t idx = (uint32_t)Ar<T>::add((T)(m * 20u), c.octave);
  T cl = Ar<T>::mul(load_coef(c, idx), c.lam);
  T one_p_beta = Ar<T>::fma(cl, dsum, (T)1);
  T beta_m_one = Ar<T>::fma(cl, dsum, (T)-1);
  T sx = (pos & 0x10u) ? xp : xm;   // p_x[sign_x], SIGN_X selects x+1
  T sy = (pos & 0x20u) ? yp : ym;   // p_y[sign_y]
  T sz = (pos & 0x40u) ? zm : zp;   // p_z[sign_z], SIGN_Z selects z-1 (kernels3d.cu:643-644)
  // dir_* are 0/1: products exact, so only the grouping of the two adds matters
  T Sb = Ar<T>::add(Ar<T>::add(Ar<T>::mul(sx, dir_x), Ar<T>::mul(sy, dir_y)), Ar<T>::mul(sz, dir_z));
  S = Ar<T>::add(S, Sb);
  T inner = Ar<T>::fma(S, c.lam2, -q);
  inner = Ar<T>::fma(p_old, beta_m_one, inner);
  return Ar<T>::mul(Ar<T>::mul(sw, inner), Ar<T>::rcp(one_p_beta));
}

// ---- node classes -----------------------------------------------------------------------------------
// The product kernel does not read the reference's two node volumes in the step loop.  Every distinct
// (position byte, material byte) pair that occurs in the mesh is a *node class* (<= 256 of them: one
// byte per voxel, class 0 = solid, class 1 = air); everything the update needs from the pair is
// evaluated once per class -- with exactly the operations of voxel_forward / voxel_centred above, so
// the per-voxel result keeps its bits -- and kept in a small table:
//   forward:  c0 = 2 - K*lam2,  c1 = -(1 - beta),  c2 = sw / (1 + beta)
//   centred:  c0 = beta - 1,    c1 = 1 / (1 + beta), c2 = sw,   flags = position byte (DIR_* / SIGN_*)
template <typename T>
struct alignas(16) ClassEntry {
  T c0, c1, c2;
  uint32_t flags;
};

enum : uint32_t { CLS_SOLID = 0, CLS_AIR = 1 };

// Wide meshes: more distinct (position, material) pairs than the 256 a class byte can name (or more lossy ones than the
// shared filter table holds).  The class byte then names the POSITION class alone; lossless classes (solid, air-like)
// keep their entry in the shared table, a lossy voxel takes table[(class - lo) * n_mat + material byte] from global
// memory, the material byte from the slab's material volume -- two dependent loads, in boundary voxels only.
template <typename T>
struct WideArgs {
  const uint8_t* mat;            // null: narrow mode
  const ClassEntry<T>* table;    // [n_lossy][n_mat]
  uint32_t lo, n_mat;
};
template <typename T, bool WIDE>
__device__ __forceinline__ ClassEntry<T> class_entry(const ClassEntry<T>* __restrict__ s_table, const WideArgs<T>& w, uint32_t cid, int64_t vox) {
  if (WIDE) { if (cid >= w.lo) return w.table[(cid - w.lo) * w.n_mat + w.mat[vox]]; }
  return s_table[cid];
}

// ---- digital impedance filters (frequency-dependent boundaries; not in the reference) ----
// Material m has the admittance Y_m(z) = (b0 + b1 z^-1 + .. + bN z^-N) / (1 + a1 z^-1 + .. + aN z^-N) acting on
// u^n = p^(n+1) - p^(n-1).  With the filter in transposed direct form II (states s_1..s_N per boundary voxel):
//     p_new = val0 - c3 * s_1          val0 = the frequency-independent update with Y = b0 (class entry c0..c2)
//     u = p_new - p_old ;  y = b0*u + s_1 ;  s_i <- b_i*u - a_i*y + s_(i+1)   (s_(N+1) = 0)
// c3 = kappa * sw/(1+beta0), kappa = 0.5*(6-K)*lam (forward / interpolated) or lam*(dir_x+dir_y+dir_z) (centred).
// Order 0 leaves p_new = val0: exactly the reference's locally-reacting boundary.
#define PFDTD_DIF_MAX_ORDER 4
template <typename T>
struct DifEntry {
  T c3, b0;
  T b[PFDTD_DIF_MAX_ORDER], a[PFDTD_DIF_MAX_ORDER];
};
template <typename T>
struct DifArgs {
  T* state;                      // [nb][P] filter states of this slab's boundary voxels, P = order rounded up to 1, 2 or 4
  const uint2* rowbase;          // [nz][Y][segs] row-segment entries (below)
  const DifEntry<T>* table;      // [n_dif] per lossy class
  uint32_t nb;
  int order;
  uint32_t dif_lo;               // class ids >= dif_lo are lossy boundary classes
  int n_dif;
  int segs;                      // row segments per row = ceil(X / 128)
  // wide meshes (more node classes than a byte / the shared tables hold): the class byte is the POSITION class only and
  // the entry of a lossy voxel is table[(class - dif_lo) * n_mat + material byte] in global memory
  const uint8_t* wide_mat;       // material byte volume of the slab, or null
  const DifEntry<T>* wide_table; // [n_dif][n_mat]
  uint32_t n_mat;
};

// entry of a lossy voxel of class `cid` whose material byte is `m` (only read in wide mode)
template <typename T, bool WIDE>
__device__ __forceinline__ const DifEntry<T>& dif_entry(const DifArgs<T>& d, const DifEntry<T>* __restrict__ s_dif, uint32_t cid, uint32_t m) {
  if (WIDE) return d.wide_table[(cid - d.dif_lo) * d.n_mat + m];
  return s_dif[cid - d.dif_lo];
}
template <typename T, bool WIDE>
__device__ __forceinline__ uint32_t dif_mat(const DifArgs<T>& d, int64_t vox) { return WIDE ? (uint32_t)d.wide_mat[vox] : 0u; }

// Row-segment entry (one per plane, row and 128-voxel tile column): x = index of the segment's first filter voxel
// (its voxels are numbered consecutively in x order), y = flags:
//   DIF_HAS     the segment has filter voxels at all
//   DIF_SINGLE  exactly one -- a wall crossing the row, by far the most common case; bits 0..6 = its x offset
//   DIF_RUN     two or more forming one contiguous run -- a row lying in a wall; bits 0..6 = x offset of the first,
//               bits 8..15 = their number
//   neither     anything else (the row crosses several walls inside one segment)
constexpr uint32_t DIF_HAS = 0x80000000u;
constexpr uint32_t DIF_SINGLE = 0x40000000u;
constexpr uint32_t DIF_RUN = 0x20000000u;
constexpr uint32_t DIF_ANY = 0x10000000u;      // in-register only: the 32-plane block this entry belongs to has filter voxels

__host__ __device__ constexpr int dif_pad(int order) { return order <= 1 ? 1 : (order == 2 ? 2 : 4); }

// the states of one voxel are P consecutive elements, moved with one (fp64 order 3-4: two) vector access
template <typename T, int N>
struct alignas(sizeof(T) * N > 16 ? 16 : sizeof(T) * N) DifVec { T v[N]; };
template <typename T, int ORD>
__device__ __forceinline__ void dif_ld(const T* p, T (&s)[ORD]) {
  const DifVec<T, dif_pad(ORD)> t = *reinterpret_cast<const DifVec<T, dif_pad(ORD)>*>(p);
#pragma unroll
  for (int i = 0; i < ORD; i++) s[i] = t.v[i];
}
template <typename T, int ORD>
__device__ __forceinline__ void dif_st(T* p, const T (&s)[ORD]) {
  DifVec<T, dif_pad(ORD)> t;
#pragma unroll
  for (int i = 0; i < dif_pad(ORD); i++) t.v[i] = i < ORD ? s[i] : (T)0;
  *reinterpret_cast<DifVec<T, dif_pad(ORD)>*>(p) = t;
}

template <typename T>
__device__ __forceinline__ T sel4(const T (&v)[4], int q) { return q == 0 ? v[0] : (q == 1 ? v[1] : (q == 2 ? v[2] : v[3])); }
template <typename T>
__device__ __forceinline__ void put4(T (&v)[4], int q, T x) {
  if (q == 0) v[0] = x;
  else if (q == 1) v[1] = x;
  else if (q == 2) v[2] = x;
  else v[3] = x;
}

template <typename T> struct Quad { T v[4]; };

template <typename T, int ORD>
__device__ __forceinline__ void dif_filter(const DifEntry<T>& e, const T (&s)[ORD], T val0, T p_old, T& p_new, T (&ns)[ORD]) {
  p_new = Ar<T>::fma(-e.c3, s[0], val0);
  const T u = Ar<T>::add(p_new, -p_old);
  const T y = Ar<T>::fma(e.b0, u, s[0]);
#pragma unroll
  for (int i = 0; i < ORD; i++) ns[i] = Ar<T>::fma(e.b[i], u, Ar<T>::fma(-e.a[i], y, i + 1 < ORD ? s[i + 1 < ORD ? i + 1 : 0] : (T)0));
}

// the two less common kinds of row segment (see DifRow): contiguous runs and the general case
template <typename T, int ORD, bool WIDE>
__device__ __noinline__ Quad<T> dif_apply_multi(Quad<T> res_q, const Quad<T> old_q, uint32_t pw, bool active, int lane, uint32_t fl,
                                                uint32_t base, const DifArgs<T>& d, const DifEntry<T>* __restrict__ s_dif, int64_t vox0) {
  constexpr int P = dif_pad(ORD);
  T (&res)[4] = res_q.v;
  const T (&old)[4] = old_q.v;
    if (fl & DIF_RUN) {
      const uint32_t cnt = (fl >> 8) & 0xffu;
      const int r0 = 4 * lane - (int)(fl & 127u);               // rank of this lane's voxel 0 (negative left of the run)
      T* sp = d.state + ((int64_t)base + r0) * P;                // only dereferenced for ranks inside the run
      T s[4][ORD];
#pragma unroll
      for (int q = 0; q < 4; q++)
        if ((uint32_t)(r0 + q) < cnt) dif_ld<T, ORD>(sp + q * P, s[q]);
#pragma unroll
      for (int q = 0; q < 4; q++)
        if ((uint32_t)(r0 + q) < cnt) {
          const DifEntry<T>& e = dif_entry<T, WIDE>(d, s_dif, (pw >> (8 * q)) & 0xffu, dif_mat<T, WIDE>(d, vox0 + q));
          T ns[ORD], p_new;
          dif_filter<T, ORD>(e, s[q], res[q], old[q], p_new, ns);
          dif_st<T, ORD>(sp + q * P, ns);
          res[q] = p_new;
        }
      return res_q;
    }
    // General case.  The lossy boundary classes are the highest class ids, so "is a filter voxel" is an unsigned
    // byte compare, done for the four bytes at once.  Ranks follow x order.
    uint32_t ge = 0;   // bit 8q+7: voxel q of this lane is a filter voxel
    if (active && pw != CLS_AIR * 0x01010101u) {
      const uint32_t k = d.dif_lo * 0x01010101u;
      const uint32_t t = (pw | 0x80808080u) - (k & 0x7f7f7f7fu);
      ge = ((pw & ~k) | (~(pw ^ k) & t)) & 0x80808080u;
    }
    const uint32_t cnt = __popc(ge);
    const uint32_t below = (1u << lane) - 1u;
    uint32_t rank = __popc(__ballot_sync(0xffffffffu, cnt & 1u) & below) + 2u * __popc(__ballot_sync(0xffffffffu, cnt & 2u) & below) +
                    4u * __popc(__ballot_sync(0xffffffffu, cnt & 4u) & below);
    for (uint32_t it = 0; it < cnt; it++, rank++) {
      const int bit = __ffs((int)ge) - 1;   // 7, 15, 23 or 31
      ge &= ge - 1u;
      const int q = bit >> 3;
      T* sp = d.state + ((size_t)base + rank) * P;
      T s[ORD], ns[ORD], p_new;
      dif_ld<T, ORD>(sp, s);
      const DifEntry<T>& e = dif_entry<T, WIDE>(d, s_dif, (pw >> (bit - 7)) & 0xffu, dif_mat<T, WIDE>(d, vox0 + q));
      dif_filter<T, ORD>(e, s, sel4<T>(res, q), sel4<T>(old, q), p_new, ns);
      dif_st<T, ORD>(sp, ns);
      put4<T>(res, q, p_new);
    }
  return res_q;
}

// Filter-boundary bookkeeping of one consumer warp (one tile row, 128 voxels, marching in z).  ORD = the filter
// order the kernel is compiled for.  The filter states are the only dependent global loads of the march, and an
// L2 round trip is longer than a plane's worth of work:
//   * entries: lane l keeps the entry of plane 32*b + l of the current block of 32 planes (one load per block);
//     DIF_ANY says whether the block has a filter voxel in this row segment at all -- segments in open air skip
//     everything with one test per plane.
//   * single-voxel planes (a wall crossing the rows: half of all row segments of a shoebox): lane l also loads the
//     states of plane 32*b + l's voxel right away, up to a whole block ahead (filter voxels are numbered z-fastest
//     within a row segment, so this is one coalesced access); the plane's turn hands them to the voxel's lane by
//     shuffle.  (Tried on B200 and dropped, profiles/r02_dif_ab.md: stashing (val0, p_old, class) per plane and running
//     the filter for 32 planes at once, from shared memory or from registers -- the one-lane evaluation it saves costs
//     less than the register pressure of the deferred pass.)
//   * runs (rows lying in a wall): every lane loads the states of its own (up to four) voxels when the plane's turn
//     comes -- one exposed round trip per plane, in the few warps that own such rows.
//   * anything else: ranked by ballots, one pass per voxel of the busiest lane.
// Per voxel (transposed direct form II, see above):
//     p_new = val0 - c3*s_1 ; u = p_new - p_old ; y = b0*u + s_1 ; s_i <- b_i*u - a_i*y + s_(i+1)
template <typename T, int ORD, bool WIDE>
struct DifRow {
  static constexpr int P = dif_pad(ORD);
  uint2 ent;            // lane l: entry of plane 32*b + l; DIF_ANY is set in every lane's flags when any plane of the block has voxels
  T st_blk[ORD];        // lane l: states of the single filter voxel of plane 32*b + l

  __device__ __forceinline__ void load_block(const DifArgs<T>& d, int z_first, int z_end, int gy, int Y, int lane) {
    const int z = z_first + lane;
    ent = (gy < Y && z < z_end) ? __ldg(d.rowbase + ((size_t)z * Y + gy) * d.segs + blockIdx.x) : make_uint2(0u, 0u);
    if (ent.y & DIF_SINGLE) dif_ld<T, ORD>(d.state + (size_t)ent.x * P, st_blk);
    if (ent.y & DIF_RUN) {   // states of a run are loaded when its plane's turn comes: pull them into L2 now (they were
                             // last touched a step ago, i.e. they are in DRAM, and a loaded DRAM round trip is several planes long)
      const char* p = reinterpret_cast<const char*>(d.state + (size_t)ent.x * P);
      const uint32_t bytes = ((ent.y >> 8) & 0xffu) * (uint32_t)(P * sizeof(T));
      for (uint32_t off = 0; off < bytes + 128u; off += 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + min(off, bytes - 1u)));
    }
    if (__ballot_sync(0xffffffffu, (ent.y & DIF_HAS) != 0u)) ent.y |= DIF_ANY;
  }
  __device__ __forceinline__ void start(const DifArgs<T>& d, int z_lo, int z_hi, int gy, int Y, int lane) {
#pragma unroll
    for (int i = 0; i < ORD; i++) st_blk[i] = (T)0;
    load_block(d, z_lo, z_hi, gy, Y, lane);
  }

  // end of plane j (of n): the next block of entries is due
  __device__ __forceinline__ void next(const DifArgs<T>& d, int j, int n, int z_lo, int z_hi, int gy, int Y, int lane) {
    if (((j + 1) & 31) == 0 && j + 1 < n) load_block(d, z_lo + j + 1, z_hi, gy, Y, lane);
  }

  // Warp-convergent, plane j of the chunk.  res = the frequency-independent results of this lane's four voxels
  // (updated in place), pw = their class bytes.
  __device__ __forceinline__ void apply(int j, T (&res)[4], const T (&old)[4], uint32_t pw, bool active, int lane, const DifArgs<T>& d,
                                        const DifEntry<T>* __restrict__ s_dif, const T* vox_ptr, const T* field) {
    if (!(ent.y & DIF_ANY)) return;   // the same in every lane
    // voxel index of this lane's voxel 0 (wide meshes look the material byte up with it): only formed where it is used
    const int64_t vox0 = WIDE ? (int64_t)(vox_ptr - field) : 0;
    const uint32_t fl = __shfl_sync(0xffffffffu, ent.y, j & 31);
    if (!(fl & DIF_HAS)) return;
    if (fl & DIF_SINGLE) {
      T s[ORD];
#pragma unroll
      for (int i = 0; i < ORD; i++) s[i] = __shfl_sync(0xffffffffu, st_blk[i], j & 31);
      const uint32_t base = __shfl_sync(0xffffffffu, ent.x, j & 31);
      if ((uint32_t)lane == ((fl & 127u) >> 2)) {
        const int q = (int)(fl & 3u);
        const DifEntry<T>& e = dif_entry<T, WIDE>(d, s_dif, (pw >> (8 * q)) & 0xffu, dif_mat<T, WIDE>(d, vox0 + q));
        T ns[ORD], p_new;
        dif_filter<T, ORD>(e, s, sel4<T>(res, q), sel4<T>(old, q), p_new, ns);
        dif_st<T, ORD>(d.state + (size_t)base * P, ns);
        put4<T>(res, q, p_new);
      }
      return;
    }
    // rows lying in a wall and everything else: out of line, so that their register needs stay out of the march
    const uint32_t base = __shfl_sync(0xffffffffu, ent.x, j & 31);
    Quad<T> r{{res[0], res[1], res[2], res[3]}}, o{{old[0], old[1], old[2], old[3]}};
    r = dif_apply_multi<T, ORD, WIDE>(r, o, pw, active, lane, fl, base, d, s_dif, vox0);
#pragma unroll
    for (int q = 0; q < 4; q++) res[q] = r.v[q];
  }
};

template <typename T, int SCHEME>
__device__ __forceinline__ ClassEntry<T> make_class_entry(uint32_t pos, uint32_t m, const UpdConst<T>& c) {
  ClassEntry<T> e;
  e.flags = pos;
  T sw = (T)(pos >> 7);
  if (pos == 0u) m = 0u;
  if (SCHEME == SCH_CENTRED) {
    T dsum = (T)((pos & 1u) + ((pos >> 1) & 1u) + ((pos >> 2) & 1u));
    uint32_t idx = c.dif_order > 0 ? m * 20u : (uint32_t)Ar<T>::add((T)(m * 20u), c.octave);
    T cl = Ar<T>::mul(load_coef(c, idx), c.lam);
    e.c0 = Ar<T>::fma(cl, dsum, (T)-1);
    e.c1 = Ar<T>::rcp(Ar<T>::fma(cl, dsum, (T)1));
    e.c2 = sw;
  } else {
    T K = (T)(pos & 0x7Fu);
    uint32_t idx = c.dif_order > 0 ? m * 20u
                   : (c.matidx_as_written ? (uint32_t)Ar<T>::mul((T)(m * 20u), c.octave) : m * 20u + (uint32_t)c.octave);
    T t = Ar<T>::mul(Ar<T>::mul(load_coef(c, idx), Ar<T>::add((T)6, -K)), c.lam);
    e.c0 = Ar<T>::fma(K, -c.lam2, (T)2);
    e.c1 = -Ar<T>::fma(t, (T)-0.5, (T)1);
    e.c2 = Ar<T>::mul(sw, Ar<T>::rcp(Ar<T>::fma(t, (T)0.5, (T)1)));
  }
  return e;
}

// ---- interpolated (27-point) compact schemes: IISO, IWB (not in the reference) ---------
//   p_new = sw/(1+beta) * ( d1*A6 + d2*A12 + d3*A8 + c0*p - (1-beta)*p_old )
//   A6/A12/A8 = sums over the axial / edge / corner neighbours (solid voxels hold 0),
//   c0 = d4 + d1*(6-K6) + d2*(12-K12) + d3*(8-K8): a missing (solid) neighbour is replaced by the centre
//   value -- the rigid-wall mirror that turns (2 - 6 lam2) into (2 - K lam2) in the reference's SRL kernel --
//   beta = 0.5*Y*(6-K6)*lam: the reference's forward-difference admittance loss (kernels3d.cu:514-528).
//   With (d1,d2,d3,d4) = (lam2, 0, 0, 2-6 lam2) this IS the reference's SRL_FORWARD equation.
template <typename T>
__device__ __forceinline__ ClassEntry<T> make_class_entry_interp(uint32_t pos, uint32_t m, uint32_t k12, uint32_t k8, const UpdConst<T>& c) {
  ClassEntry<T> e;
  e.flags = pos;
  const T sw = (T)(pos >> 7);
  if (pos == 0u) m = 0u;
  const T K = (T)(pos & 0x7Fu);
  const uint32_t idx = c.dif_order > 0 ? m * 20u : m * 20u + (uint32_t)c.octave;
  const T t = Ar<T>::mul(Ar<T>::mul(load_coef(c, idx), Ar<T>::add((T)6, -K)), c.lam);
  T c0 = Ar<T>::fma(c.d[0], Ar<T>::add((T)6, -K), c.d[3]);
  c0 = Ar<T>::fma(c.d[1], (T)(12u - k12), c0);
  c0 = Ar<T>::fma(c.d[2], (T)(8u - k8), c0);
  e.c0 = c0;
  e.c1 = -Ar<T>::fma(t, (T)-0.5, (T)1);
  e.c2 = Ar<T>::mul(sw, Ar<T>::rcp(Ar<T>::fma(t, (T)0.5, (T)1)));
  return e;
}
// DIF entry of a lossy class: c3 = kappa * (sw / (1 + beta0)); material row = [b0..bN, a1..aN]
template <typename T, int SCHEME>
__device__ __forceinline__ DifEntry<T> make_dif_entry(uint32_t pos, uint32_t m, const UpdConst<T>& c, int order) {
  DifEntry<T> e;
  const T* row = c.materials + (size_t)m * 20u;
  const T b0 = __ldg(row);
  const T sw = (T)(pos >> 7);
  T kap, rc;
  if (SCHEME == SCH_CENTRED) {
    const T dsum = (T)((pos & 1u) + ((pos >> 1) & 1u) + ((pos >> 2) & 1u));
    kap = Ar<T>::mul(c.lam, dsum);
    rc = Ar<T>::mul(sw, Ar<T>::rcp(Ar<T>::fma(Ar<T>::mul(b0, c.lam), dsum, (T)1)));
  } else {
    const T K = (T)(pos & 0x7Fu);
    kap = Ar<T>::mul(Ar<T>::mul((T)0.5, Ar<T>::add((T)6, -K)), c.lam);
    const T t = Ar<T>::mul(Ar<T>::mul(b0, Ar<T>::add((T)6, -K)), c.lam);
    rc = Ar<T>::mul(sw, Ar<T>::rcp(Ar<T>::fma(t, (T)0.5, (T)1)));
  }
  e.c3 = Ar<T>::mul(kap, rc);
  e.b0 = b0;
  for (int i = 0; i < PFDTD_DIF_MAX_ORDER; i++) {
    e.b[i] = i < order ? __ldg(row + 1 + i) : (T)0;
    e.a[i] = i < order ? __ldg(row + 1 + order + i) : (T)0;
  }
  return e;
}

// in-plane partial sums of one voxel: a4 = ((x- + x+) + y-) + y+ ; g4 = ((x-y- + x+y-) + x-y+) + x+y+
template <typename T>
__device__ __forceinline__ T interp_a4(T xm, T xp, T ym, T yp) { return Ar<T>::add(Ar<T>::add(Ar<T>::add(xm, xp), ym), yp); }
template <typename T>
__device__ __forceinline__ T interp_g4(T mm, T pm, T mp, T pp) { return Ar<T>::add(Ar<T>::add(Ar<T>::add(mm, pm), mp), pp); }
// combine the partial sums of planes z-1, z, z+1:  A6 = (a4 + c-) + c+ ; A12 = (g4 + a4-) + a4+ ; A8 = g4- + g4+
template <typename T, bool HAS_D3>
__device__ __forceinline__ T voxel_interp(T c0, T c1, T c2, T p, T a4, T g4, T c_m, T a4_m, T g4_m, T c_p, T a4_p, T g4_p, T p_old,
                                          T d1, T d2, T d3) {
  const T A6 = Ar<T>::add(Ar<T>::add(a4, c_m), c_p);
  const T A12 = Ar<T>::add(Ar<T>::add(g4, a4_m), a4_p);
  T inner = Ar<T>::fma(A6, d1, Ar<T>::mul(p, c0));
  inner = Ar<T>::fma(A12, d2, inner);
  if (HAS_D3) inner = Ar<T>::fma(Ar<T>::add(g4_m, g4_p), d3, inner);
  inner = Ar<T>::fma(p_old, c1, inner);
  return Ar<T>::mul(c2, inner);
}

// forward, any class: identical bits to voxel_forward (air: c1 = -1, c2 = 1, multiplying by 1 is exact)
template <typename T>
__device__ __forceinline__ T voxel_forward_cls(const ClassEntry<T>& e, T p, T S, T p_old, T lam2) {
  T inner = Ar<T>::fma(S, lam2, Ar<T>::mul(p, e.c0));
  inner = Ar<T>::fma(p_old, e.c1, inner);
  return Ar<T>::mul(e.c2, inner);
}
template <typename T>
__device__ __forceinline__ T voxel_forward_air(T p, T S, T p_old, T lam2, T a_air) {
  T inner = Ar<T>::fma(S, lam2, Ar<T>::mul(p, a_air));
  return Ar<T>::fma(p_old, (T)-1, inner);
}

template <typename T>
__device__ __forceinline__ T sum6_centred(T zp, T zm, T yp, T ym, T xp, T xm) {
  T S = Ar<T>::add(xm, xp);
  S = Ar<T>::add(S, ym);
  S = Ar<T>::add(S, yp);
  S = Ar<T>::add(S, zp);
  S = Ar<T>::add(S, zm);
  return S;
}
template <typename T>
__device__ __forceinline__ T voxel_centred_air(T p, T S, T p_old, T lam2, T a_air) {
  S = Ar<T>::add(S, (T)0);
  T inner = Ar<T>::fma(S, lam2, -Ar<T>::mul(p, a_air));
  return Ar<T>::fma(p_old, (T)-1, inner);
}
template <typename T>
__device__ __forceinline__ T voxel_centred_cls(const ClassEntry<T>& e, T p, T S, T zp, T zm, T yp, T ym, T xp, T xm, T p_old,
                                               T lam2, T a_air) {
  const uint32_t pos = e.flags;
  T dir_x = (T)(pos & 1u), dir_y = (T)((pos >> 1) & 1u), dir_z = (T)((pos >> 2) & 1u);
  T sx = (pos & 0x10u) ? xp : xm;
  T sy = (pos & 0x20u) ? yp : ym;
  T sz = (pos & 0x40u) ? zm : zp;
  T Sb = Ar<T>::add(Ar<T>::add(Ar<T>::mul(sx, dir_x), Ar<T>::mul(sy, dir_y)), Ar<T>::mul(sz, dir_z));
  S = Ar<T>::add(S, Sb);
  T inner = Ar<T>::fma(S, lam2, -Ar<T>::mul(p, a_air));
  inner = Ar<T>::fma(p_old, e.c0, inner);
  return Ar<T>::mul(Ar<T>::mul(e.c2, inner), e.c1);
}

template <typename T, int SCHEME>
__device__ __forceinline__ T voxel_update(uint32_t pos, const uint8_t* mat_ptr, T p, T zp, T zm, T yp, T ym, T xp, T xm,
                                          T p_old, const UpdConst<T>& c) {
  if (SCHEME == SCH_CENTRED) return voxel_centred<T>(pos, mat_ptr, p, zp, zm, yp, ym, xp, xm, p_old, c);
  return voxel_forward<T>(pos, mat_ptr, p, sum6_forward<T>(zp, zm, yp, ym, xp, xm), p_old, c);
}

}  // namespace pfdtd
