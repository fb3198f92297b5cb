// interp_kernels.cu -- the 27-point "interpolated" compact explicit schemes (IISO, IWB) as a TMA z-march.
//
// Not in the mounted reference: the update equation is the compact explicit family of
// Kowalczyk & van Walstijn with the reference's forward-difference boundary carried over (update_math.cuh,
// "interpolated"); parity is against our own CPU oracle, anchored on the fact that (d1,d2,d3,d4) =
// (lam2,0,0,2-6lam2) is the reference's SRL_FORWARD equation.
//
// Same producer/consumer structure, shared-memory stages and tensor maps as fdtd_update_tma.  What
// changes is what a consumer keeps in registers: per voxel and per plane the three in-plane partial sums
//   c  = centre,  a4 = x-+x++y-+y+ (axial),  g4 = the four in-plane diagonals,
// for planes z-1, z, z+1.  A plane's tile is read from shared memory exactly once, when it arrives as z+1
// (three 128-bit row loads + six shuffles per lane), and the 27-point sums are assembled from registers:
//   A6 = a4(z) + c(z-1) + c(z+1),  A12 = g4(z) + a4(z-1) + a4(z+1),  A8 = g4(z-1) + g4(z+1).
// HBM traffic per voxel update is the same 13 B / 25 B as the 7-point kernel.
#include "pfdtd_internal.h"
#include "update_math.cuh"
#include "tma_common.cuh"
#include "update_host.cuh"

namespace pfdtd {

namespace {

template <typename T> struct Plane3 { T c[4], a4[4], g4[4]; };

// in-plane sums of the lane's four voxels of tile row `r` (0-based, without halo) from a halo tile
template <typename T, int TY, bool HAS_D3>
__device__ __forceinline__ void load_plane(const T* __restrict__ pt, int r, int lane, Plane3<T>& o) {
  using G = TileGeom<T, TY>;
  const int xl = 4 * lane;
  V4<T> vm, v0, vp;
  const T* rowm = pt + (r + 0) * G::PW + G::HX;
  const T* row0 = pt + (r + 1) * G::PW + G::HX;
  const T* rowp = pt + (r + 2) * G::PW + G::HX;
  lds4(rowm + xl, vm);
  lds4(row0 + xl, v0);
  lds4(rowp + xl, vp);
  T l0 = __shfl_up_sync(0xffffffffu, v0.v[3], 1), r0 = __shfl_down_sync(0xffffffffu, v0.v[0], 1);
  T lm = __shfl_up_sync(0xffffffffu, vm.v[3], 1), rm = __shfl_down_sync(0xffffffffu, vm.v[0], 1);
  T lp = __shfl_up_sync(0xffffffffu, vp.v[3], 1), rp = __shfl_down_sync(0xffffffffu, vp.v[0], 1);
  if (lane == 0) { l0 = row0[-1]; lm = rowm[-1]; lp = rowp[-1]; }
  if (lane == 31) { r0 = row0[TX]; rm = rowm[TX]; rp = rowp[TX]; }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const T xm = q == 0 ? l0 : v0.v[q - 1 < 0 ? 0 : q - 1], xp = q == 3 ? r0 : v0.v[q + 1 > 3 ? 3 : q + 1];
    const T mm = q == 0 ? lm : vm.v[q - 1 < 0 ? 0 : q - 1], pm = q == 3 ? rm : vm.v[q + 1 > 3 ? 3 : q + 1];
    const T mp = q == 0 ? lp : vp.v[q - 1 < 0 ? 0 : q - 1], pp = q == 3 ? rp : vp.v[q + 1 > 3 ? 3 : q + 1];
    o.c[q] = v0.v[q];
    o.a4[q] = interp_a4<T>(xm, xp, vm.v[q], vp.v[q]);
    o.g4[q] = interp_g4<T>(mm, pm, mp, pp);
  }
}

}  // namespace

// grid: (ceil(X/128), ceil(Y/TY), n_chunks); block: TY consumer warps (one tile row each) + 1 producer warp
template <typename T, bool HAS_D3, int TY, int NST, int DIF, bool WIDE, int TAIL>
__global__ void __launch_bounds__((TY + 1) * 32) __maxnreg__(sizeof(T) == 4 ? (TY == 7 ? (NST >= 6 ? 80 : 64) : (TY == 8 ? 72 : 56)) : (TY == 7 ? 128 : 96))
    fdtd_update_interp_tma(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_old,
                           const __grid_constant__ CUtensorMap tm_cls, const ClassEntry<T>* __restrict__ g_table, int n_classes,
                           T* __restrict__ Pn, T d1, T d2, T d3, T d4, int X, int Y, int z_begin, int z_end, int chunk, int hints,
                           const DifArgs<T> dif, const WideArgs<T> wide, T* __restrict__ peer, const __grid_constant__ FusedParams fused_p,
                    const FusedSrcRec<T>* __restrict__ fused) {
  using G = TileGeom<T, TY>;
  constexpr int NW = TY;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[NST];
  __shared__ __align__(8) uint64_t bar_empty[NST];
  __shared__ ClassEntry<T> s_table[256];
  __shared__ DifEntry<T> s_dif[DIF ? 64 : 1];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int cby, cbz;
  cta_tile(cby, cbz);
  const int x0 = blockIdx.x * TX;
  const int y0 = cby * TY;
  const int z_lo = z_begin + cbz * chunk;
  const int z_hi = min(z_lo + chunk, z_end);
  const int n = z_hi - z_lo;
  if (n <= 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; s++) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_classes; i += blockDim.x) s_table[i] = g_table[i];
  if (DIF) for (int i = threadIdx.x; i < dif.n_dif; i += blockDim.x) s_dif[i] = dif.table[i];
  __syncthreads();

  if (warp == NW) {
    if (lane == 0) {
      // load i: P(z_lo-1+i) with xy halo; from i >= 2 also P_old / class bytes of plane z_lo+i-2
      for (int i = 0; i < n + 2; i++) {
        const int slot = i % NST;
        const uint32_t round = (uint32_t)(i / NST);
        mbar_wait(&bar_empty[slot], (round & 1u) ^ 1u);
        unsigned char* st = smem_raw + (size_t)slot * G::STAGE_BYTES;
        const uint32_t bytes = (i >= 2) ? (uint32_t)(G::PT_BYTES + G::PO_BYTES + G::PS_BYTES) : (uint32_t)G::PT_BYTES;
        mbar_expect_tx(&bar_full[slot], bytes);
        tma_load_3d(st + G::PT_OFF, &tm_p, x0 - G::HX, y0 - 1, z_lo - 1 + i, &bar_full[slot]);
        if (i >= 2) {
          tma_load_3d(st + G::PO_OFF, &tm_old, x0, y0, z_lo + i - 2, &bar_full[slot]);
          tma_load_3d(st + G::PS_OFF, &tm_cls, x0, y0, z_lo + i - 2, &bar_full[slot]);
        }
      }
    }
    return;
  }

  const int r = warp;                        // tile row of this warp
  const int xl = 4 * lane;
  const int gx = x0 + xl;
  const int gy = y0 + r;
  const bool active = gx < X && gy < Y;
  const int64_t XY = (int64_t)X * Y;
  constexpr uint32_t AIR4 = CLS_AIR * 0x01010101u;

  // filter boundaries: entries and prefetched states of this warp's row (update_math.cuh DifRow), started before the
  // first wait on the pipeline so that the entry -> state round trips overlap the first planes' TMA loads
  constexpr int DMO = DIF ? DIF : 1;
  DifRow<T, DMO, WIDE> drow;
  if (DIF) drow.start(dif, z_lo, z_hi, gy, Y, lane);
  Plane3<T> pm, pc, pp;                      // planes z-1, z, z+1
  {
    mbar_wait(&bar_full[0], 0);
    load_plane<T, TY, HAS_D3>(reinterpret_cast<const T*>(smem_raw + G::PT_OFF), r, lane, pm);
    mbar_wait(&bar_full[1 % NST], (uint32_t)((1 / NST) & 1));
    load_plane<T, TY, HAS_D3>(reinterpret_cast<const T*>(smem_raw + (size_t)(1 % NST) * G::STAGE_BYTES + G::PT_OFF), r, lane, pc);
    // No stage is handed back before a global store that depends on what was loaded from it: an issued but not
    // yet performed shared-memory load is not ordered before the mbarrier arrive, so the producer's next TMA
    // write into the stage could overtake it (update_kernels.cu has the B200 observation).
  }

  // stage and barrier parity of plane z+1 are carried along instead of being derived from the plane index (a division per
  // plane in a kernel that is issue-bound: ncu, 72 % of the issue slots busy with filters on); the voxel address likewise
  int s2c = 2 % NST;
  uint32_t par2c = (uint32_t)((2 / NST) & 1);
  // (fp64 derives all three from the plane index instead: at its 128-register budget every value carried across the
  // filter pass spills, and the carried form is 9 % slower there)
  constexpr bool CARRY_PTR = sizeof(T) == 4;
  T* vox_carry = Pn + (int64_t)z_lo * XY + (int64_t)gy * X + gx;     // this lane's four voxels of plane z
  for (int j = 0; j < n; j++) {
    const int s2 = CARRY_PTR ? s2c : (j + 2) % NST;
    const uint32_t par2 = CARRY_PTR ? par2c : (uint32_t)(((j + 2) / NST) & 1);
    mbar_wait(&bar_full[s2], par2);
    const unsigned char* st2 = smem_raw + (size_t)s2 * G::STAGE_BYTES;
    load_plane<T, TY, HAS_D3>(reinterpret_cast<const T*>(st2 + G::PT_OFF), r, lane, pp);
    V4<T> old;
    lds4(reinterpret_cast<const T*>(st2 + G::PO_OFF) + r * TX + xl, old);
    const uint32_t pw = *reinterpret_cast<const uint32_t*>(st2 + G::PS_OFF + r * TX + xl);
    V4<T> res;
    if (active) {
      if (pw == AIR4) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          res.v[q] = voxel_interp<T, HAS_D3>(d4, (T)-1, (T)1, pc.c[q], pc.a4[q], pc.g4[q], pm.c[q], pm.a4[q], pm.g4[q], pp.c[q],
                                             pp.a4[q], pp.g4[q], old.v[q], d1, d2, d3);
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const ClassEntry<T> ce = class_entry<T, WIDE>(s_table, wide, (pw >> (8 * q)) & 0xffu,
                                                        WIDE ? (int64_t)(z_lo + j) * XY + (int64_t)gy * X + gx + q : 0);
          res.v[q] = voxel_interp<T, HAS_D3>(ce.c0, ce.c1, ce.c2, pc.c[q], pc.a4[q], pc.g4[q], pm.c[q], pm.a4[q], pm.g4[q], pp.c[q],
                                             pp.a4[q], pp.g4[q], old.v[q], d1, d2, d3);
        }
      }
    }
    // (the voxel address is formed after the filter pass in narrow mode: the pass is the register-pressure peak of the loop)
    if (DIF) drow.apply(j, res.v, old.v, pw, active, lane, dif, s_dif,
                        (CARRY_PTR || !WIDE) ? (const T*)vox_carry : (const T*)(Pn + (int64_t)(z_lo + j) * XY + (int64_t)gy * X + gx), Pn);
    if (active) {
      T* const vox_ptr = CARRY_PTR ? vox_carry : Pn + (int64_t)(z_lo + j) * XY + (int64_t)gy * X + gx;
      stg4(vox_ptr, res);
      if (TAIL == 1 && peer != nullptr) stg4(peer + (int64_t)gy * X + gx, res);   // edge launch: the neighbour slab's halo plane (update_kernels.cu)
    }
    // the store above consumed everything read from stage s2 (and, at j == 0, from the two prologue stages)
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&bar_empty[s2]);
      if (j == 0) { mbar_arrive(&bar_empty[0]); mbar_arrive(&bar_empty[1 % NST]); }
    }
    if (DIF) drow.next(dif, j, n, z_lo, z_hi, gy, Y, lane);
    pm = pc;
    pc = pp;
    if (CARRY_PTR) vox_carry += XY;
    if (CARRY_PTR && ++s2c == NST) { s2c = 0; par2c ^= 1u; }
  }
  if (TAIL == 2) fused_srcrec<T>(fused_p, fused, Pn, X, Y, TY, z_begin, z_end, chunk, hints, NW * 32);
  // edge launch of a slab (one plane): send the plane into the neighbour slab's halo plane (update_kernels.cu)
}

// one thread per voxel; same arithmetic (fallback for dimensions the TMA path does not take, cross-check)
template <typename T, bool HAS_D3>
__global__ void __launch_bounds__(128) fdtd_update_interp_plain(const uint8_t* __restrict__ cls, const ClassEntry<T>* __restrict__ table,
                                                                const T* __restrict__ P, T* __restrict__ Pn, T d1, T d2, T d3, int X,
                                                                int Y, int z_begin, const WideArgs<T> wide) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = z_begin + blockIdx.z;
  if (x >= X || y >= Y) return;
  const int64_t XY = (int64_t)X * Y;
  auto at = [&](int xx, int yy, int zz) -> T {   // outside the xy extent reads 0, like the TMA tile's out-of-bounds fill
    if (xx < 0 || yy < 0 || xx >= X || yy >= Y) return (T)0;
    return P[(int64_t)zz * XY + (int64_t)yy * X + xx];
  };
  T c[3], a4[3], g4[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int zz = z - 1 + k;
    c[k] = at(x, y, zz);
    a4[k] = interp_a4<T>(at(x - 1, y, zz), at(x + 1, y, zz), at(x, y - 1, zz), at(x, y + 1, zz));
    g4[k] = interp_g4<T>(at(x - 1, y - 1, zz), at(x + 1, y - 1, zz), at(x - 1, y + 1, zz), at(x + 1, y + 1, zz));
  }
  const int64_t cur = (int64_t)z * XY + (int64_t)y * X + x;
  const ClassEntry<T> ce = (wide.mat != nullptr ? class_entry<T, true>(table, wide, cls[cur], cur) : table[cls[cur]]);
  Pn[cur] = voxel_interp<T, HAS_D3>(ce.c0, ce.c1, ce.c2, c[1], a4[1], g4[1], c[0], a4[0], g4[0], c[2], a4[2], g4[2], Pn[cur], d1, d2, d3);
}

namespace {

template <typename T, bool HAS_D3, int TY, int NST, int DIF, bool WIDE = false, int TAIL = 0>
int launch_interp_t(const UpdateArgs& a, const TmaMaps& m, int chunk, int* occupancy_out) {
  auto kern = fdtd_update_interp_tma<T, HAS_D3, TY, NST, DIF, WIDE, TAIL>;
  const int smem = NST * TileGeom<T, TY>::STAGE_BYTES;
  static bool attr_set[64] = {false};
  int dev = 0;
  PF_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dev] = true;
  }
  const int threads = (TY + 1) * 32;
  if (occupancy_out) {
    PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy_out, kern, threads, smem));
    return PFDTD_OK;
  }
  const int nplanes = a.z_end - a.z_begin;
  dim3 grid((a.X + TX - 1) / TX, (a.Y + TY - 1) / TY, (nplanes + chunk - 1) / chunk);
  const UpdConst<T> c = make_const<T>(a);
  kern<<<grid, threads, smem, a.stream>>>(m.p_halo, m.p_old, m.cls, (const ClassEntry<T>*)a.class_table, a.n_classes, (T*)a.Pn, c.d[0],
                                          c.d[1], c.d[2], c.d[3], a.X, a.Y, a.z_begin, a.z_end, chunk, a.tma_hints, make_dif<T>(a), make_wide<T>(a),
                                          (a.z_end - a.z_begin == 1) ? (T*)a.peer_plane : nullptr,
                                          a.fused_params, (const FusedSrcRec<T>*)a.fused_srcrec);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

// filter boundaries, wide meshes and launches with a tail: one kernel per order on the 128x7 shape (fp32 six stages:
// 80 registers, three CTAs per SM; fp64 five)
template <typename T, bool HAS_D3, bool WIDE, int TAIL>
int dispatch_interp_order(const UpdateArgs& a, const TmaMaps& m, int chunk, int* occ) {
  constexpr int DTY = 7, DNST = sizeof(T) == 4 ? 6 : 5;
  switch (a.dif_order) {
    case 0: return launch_interp_t<T, HAS_D3, DTY, DNST, 0, WIDE, TAIL>(a, m, chunk, occ);
    case 1: return launch_interp_t<T, HAS_D3, DTY, DNST, 1, WIDE, TAIL>(a, m, chunk, occ);
    case 2: return launch_interp_t<T, HAS_D3, DTY, DNST, 2, WIDE, TAIL>(a, m, chunk, occ);
    case 3: return launch_interp_t<T, HAS_D3, DTY, DNST, 3, WIDE, TAIL>(a, m, chunk, occ);
    case 4: return launch_interp_t<T, HAS_D3, DTY, DNST, 4, WIDE, TAIL>(a, m, chunk, occ);
  }
  set_error("filter order %d is not supported", a.dif_order);
  return PFDTD_ERR_INVALID;
}

template <typename T, bool HAS_D3>
int dispatch_interp(const UpdateArgs& a, const TmaMaps& m, int tile, int chunk, int* occ) {
  // tile variants shared with the 7-point kernel: only the one-row-per-warp shapes apply here
  constexpr int DTY = 7, DNST = sizeof(T) == 4 ? 6 : 5;
  if (a.tail == 1) return a.wide ? dispatch_interp_order<T, HAS_D3, true, 1>(a, m, chunk, occ) : dispatch_interp_order<T, HAS_D3, false, 1>(a, m, chunk, occ);
  if (a.tail == 2) {
    PF_CHECK(a.dif_order == 0 && !a.wide, PFDTD_ERR_INVALID, "fused sources / receivers are built for the frequency-independent kernels");
    return launch_interp_t<T, HAS_D3, DTY, DNST, 0, false, 2>(a, m, chunk, occ);
  }
  if (a.wide) return dispatch_interp_order<T, HAS_D3, true, 0>(a, m, chunk, occ);
  if (a.dif_order > 0) return dispatch_interp_order<T, HAS_D3, false, 0>(a, m, chunk, occ);
  switch (tile) {
    case 0: case 3: return launch_interp_t<T, HAS_D3, 8, 4, 0>(a, m, chunk, occ);
    case 2: case 1: case 4: return launch_interp_t<T, HAS_D3, 16, 4, 0>(a, m, chunk, occ);
    case 6: case 7: return launch_interp_t<T, HAS_D3, 7, 5, 0>(a, m, chunk, occ);
    case 8: return launch_interp_t<T, HAS_D3, 8, 5, 0>(a, m, chunk, occ);
  }
  set_error("tile variant %d is not available for the interpolated schemes", tile);
  return PFDTD_ERR_INVALID;
}

}  // namespace

int launch_update_interp_tma(const UpdateArgs& a, const TmaMaps& maps, const TmaConfig& cfg, int* occupancy_out) {
  if (!occupancy_out && a.z_end <= a.z_begin) return PFDTD_OK;
  const bool d3 = a.dcoef[2] != 0.0;
  if (a.dtype == PFDTD_F32) return d3 ? dispatch_interp<float, true>(a, maps, cfg.tile, cfg.chunk, occupancy_out)
                                      : dispatch_interp<float, false>(a, maps, cfg.tile, cfg.chunk, occupancy_out);
  return d3 ? dispatch_interp<double, true>(a, maps, cfg.tile, cfg.chunk, occupancy_out)
            : dispatch_interp<double, false>(a, maps, cfg.tile, cfg.chunk, occupancy_out);
}

int launch_update_interp_plain(const UpdateArgs& a) {
  if (a.z_end <= a.z_begin) return PFDTD_OK;
  dim3 block(32, 4, 1);
  dim3 grid((a.X + 31) / 32, (a.Y + 3) / 4, a.z_end - a.z_begin);
  const bool d3 = a.dcoef[2] != 0.0;
  if (a.dtype == PFDTD_F32) {
    const UpdConst<float> c = make_const<float>(a);
    if (d3) fdtd_update_interp_plain<float, true><<<grid, block, 0, a.stream>>>(a.cls, (const ClassEntry<float>*)a.class_table, (const float*)a.P, (float*)a.Pn, c.d[0], c.d[1], c.d[2], a.X, a.Y, a.z_begin, make_wide<float>(a));
    else fdtd_update_interp_plain<float, false><<<grid, block, 0, a.stream>>>(a.cls, (const ClassEntry<float>*)a.class_table, (const float*)a.P, (float*)a.Pn, c.d[0], c.d[1], c.d[2], a.X, a.Y, a.z_begin, make_wide<float>(a));
  } else {
    const UpdConst<double> c = make_const<double>(a);
    if (d3) fdtd_update_interp_plain<double, true><<<grid, block, 0, a.stream>>>(a.cls, (const ClassEntry<double>*)a.class_table, (const double*)a.P, (double*)a.Pn, c.d[0], c.d[1], c.d[2], a.X, a.Y, a.z_begin, make_wide<double>(a));
    else fdtd_update_interp_plain<double, false><<<grid, block, 0, a.stream>>>(a.cls, (const ClassEntry<double>*)a.class_table, (const double*)a.P, (double*)a.Pn, c.d[0], c.d[1], c.d[2], a.X, a.Y, a.z_begin, make_wide<double>(a));
  }
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

}  // namespace pfdtd
