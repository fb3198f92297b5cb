// update_kernels.cu -- the leapfrog pressure update (the hot path).
//
// Replaces fdtd3dStdMaterials / fdtd3dSliced / fdtd3dStdKowalczykMaterials
// (reference src/kernels/kernels3d.cu:485-665).  Two implementations of the same arithmetic
// (update_math.cuh):
//
//  * fdtd_update_tma  -- the product kernel.  Each CTA owns a 128 x TY xy-tile and marches a chunk of
//    z-planes.  A producer warp streams, per plane, three TMA boxes into a ring of shared-memory
//    stages (mbarrier full/empty pipeline): the current field with a one-voxel xy halo, the field
//    being overwritten, and the packed node byte (the node *class*: one byte that stands for the
//    reference's position byte AND material byte, see update_math.cuh).  Consumer warps keep
//    z-1 / z / z+1 of their four x-adjacent voxels in registers, take y-neighbours from the shared
//    tile, x-neighbours by warp shuffle (halo column from the tile), and write the result with one
//    128-bit store per four voxels.  The boundary term is evaluated in the same pass from a per-class
//    table in shared memory (no dependent global loads, no separate boundary kernel).  HBM traffic per
//    voxel update: read P 4(8) B once, read P_old 4(8) B, write 4(8) B, node byte 1 B = 13 B fp32 /
//    25 B fp64.
//
//  * fdtd_update_plain -- one thread per voxel, neighbour reads through L1/L2; used for dimensions the
//    TMA path does not cover and as an on-device cross-check.
#include "pfdtd_internal.h"
#include "update_math.cuh"
#include "tma_common.cuh"
#include "update_host.cuh"

#include <algorithm>
#include <cstring>
#include <mutex>

namespace pfdtd {

// =====================================================================================================
// plain kernel
// =====================================================================================================
template <typename T, int SCHEME>
__global__ void __launch_bounds__(128) fdtd_update_plain(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ mat,
                                                         const T* __restrict__ P, T* __restrict__ Pn, UpdConst<T> c, int X,
                                                         int Y, int z_begin) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = z_begin + blockIdx.z;
  if (x >= X || y >= Y) return;
  const int64_t XY = (int64_t)X * Y;
  const int64_t cur = (int64_t)z * XY + (int64_t)y * X + x;
  const uint32_t ps = pos[cur];
  // neighbour indexing identical to the reference: linear offsets, rows/slices wrap (kernels3d.cu:516-523)
  T zp = P[cur + XY], zm = P[cur - XY], yp = P[cur + X], ym = P[cur - X], xp = P[cur + 1], xm = P[cur - 1];
  T p = P[cur], p_old = Pn[cur];
  Pn[cur] = voxel_update<T, SCHEME>(ps, mat + cur, p, zp, zm, yp, ym, xp, xm, p_old, c);
}

template <typename T, int SCHEME>
static int launch_plain_t(const UpdateArgs& a) {
  dim3 block(32, 4, 1);
  dim3 grid((a.X + 31) / 32, (a.Y + 3) / 4, a.z_end - a.z_begin);
  fdtd_update_plain<T, SCHEME><<<grid, block, 0, a.stream>>>(a.pos, a.mat, (const T*)a.P, (T*)a.Pn, make_const<T>(a), a.X,
                                                             a.Y, a.z_begin);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

int launch_update_plain(const UpdateArgs& a) {
  if (a.z_end <= a.z_begin) return PFDTD_OK;
  if (a.dtype == PFDTD_F32) return a.scheme == SCH_CENTRED ? launch_plain_t<float, SCH_CENTRED>(a) : launch_plain_t<float, SCH_FORWARD>(a);
  return a.scheme == SCH_CENTRED ? launch_plain_t<double, SCH_CENTRED>(a) : launch_plain_t<double, SCH_FORWARD>(a);
}

// =====================================================================================================
// TMA z-march kernel
// =====================================================================================================
// Register budget per variant.  Registers are allocated per SM sub-partition (4 x 16384), so k resident warps cost
// ceil(k/4) warps' worth on the fullest one.  The default 128x8 one-row-per-warp shape (9 warps, 42.5 / 79 KiB of
// stages): fp32 forward fits four CTAs per SM in 56 registers, the centred scheme and the filter kernels three in
// 72, fp64 two in 96; the other shapes just have to fit one CTA.
constexpr int tma_max_regs(size_t esize, int scheme, int ty, int rpw, int nst, int dif) {
  // 128x7: seven consumer warps + the producer = 8 warps, two per sub-partition and CTA: 64 registers with four
  // resident CTAs (fp32), 80 with three, 128 with two (fp64)
  if (ty == 7 && rpw == 1) return esize == 4 ? (nst >= 6 ? 80 : 64) : 128;
  if (ty == 8 && rpw == 1) return esize == 4 ? ((dif || scheme == SCH_CENTRED) ? 72 : 56) : 96;
  if (ty == 16 && rpw == 1 && esize == 4 && scheme != SCH_CENTRED) return 56;   // two CTAs of 17 warps
  const int warps = ty / rpw + 1;
  const int fit = (16384 / (32 * ((warps + 3) / 4))) & ~7;
  return fit > 255 ? 255 : fit;
}

// Unroll factor of the plane loop (see the loop).  The frequency-independent kernels unroll over all stages (they run at
// the copy bandwidth that way).  The filter kernels do NOT unroll: their six-fold body (3584 / 4344 SASS instructions
// forward / centred) misses the instruction cache -- ncu `no_instruction` 0.5 / 2.1 stalls per issue
// (profiles/r02_ncu_f32_centred_dif2_512_unrolled.json) -- and the rolled loop is 3 % / 14 % faster at 512^3 in fp32, 4 % / 7 % in
// fp64; for the frequency-independent kernels it changes nothing (profiles/r02_dif_ab.md).  0 = all stages.
#ifndef PFDTD_UNR_DIF_F32
#define PFDTD_UNR_DIF_F32 1
#endif
#ifndef PFDTD_UNR_DIF_F64
#define PFDTD_UNR_DIF_F64 1
#endif
#ifndef PFDTD_UNR_NODIF
#define PFDTD_UNR_NODIF 0
#endif
__host__ __device__ constexpr int tma_unroll(size_t esize, int scheme, int dif, int nst) {
  (void)scheme;
  const int want = dif > 0 ? (esize == 4 ? PFDTD_UNR_DIF_F32 : PFDTD_UNR_DIF_F64) : PFDTD_UNR_NODIF;
  return want > 0 ? want : nst;
}

// grid: (ceil(X/128), ceil(Y/TY), n_chunks); block: (TY/RPW + 1) warps (last warp = TMA producer)
// DIF = 0: frequency-independent boundaries; 1..4: digital impedance filters of that order
// (one-row-per-warp 128x8 tile: three resident CTAs per SM in fp32, two in fp64)
// TAIL: what a CTA does after its march -- 0 nothing (the bulk launches), 1 edge launch of a slab (store the plane into the
// neighbour's halo plane and publish the step), 2 record receivers / inject the next sources (single small slab)
template <typename T, int SCHEME, int TY, int RPW, int NST, int DIF, bool WIDE, int TAIL>
__global__ void __launch_bounds__((TY / RPW + 1) * 32) __maxnreg__(tma_max_regs(sizeof(T), SCHEME, TY, RPW, NST, DIF))
    fdtd_update_tma(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_old,
                    const __grid_constant__ CUtensorMap tm_cls, const ClassEntry<T>* __restrict__ g_table, int n_classes,
                    T* __restrict__ Pn, T lam2, T a_air, int X, int Y, int z_begin, int z_end, int chunk, int hints,
                    const DifArgs<T> dif, const WideArgs<T> wide, T* __restrict__ peer, const __grid_constant__ FusedParams fused_p,
                    const FusedSrcRec<T>* __restrict__ fused) {
  using G = TileGeom<T, TY>;
  constexpr int NW = TY / RPW;
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
  const int n = z_hi - z_lo;              // planes this CTA computes
  if (n <= 0) return;
  if (DIF) for (int i = threadIdx.x; i < dif.n_dif; i += blockDim.x) s_dif[i] = dif.table[i];

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; s++) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_classes; i += blockDim.x) s_table[i] = g_table[i];
  __syncthreads();

  if (warp == NW) {
    // ------------------------------- producer -------------------------------------------------
    if (lane == 0) {
      // P is shared with the neighbouring tiles (halo rows/columns) and re-read by the z-neighbour chunk;
      // P_old and the node bytes are touched exactly once per step
      const uint64_t pol_once = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      // load i brings P(z_lo-1+i) with halo and, from i >= 2, P_old / node byte of plane z_lo+i-2
      for (int i = 0; i < n + 2; i++) {
        const int slot = i % NST;
        const uint32_t round = (uint32_t)(i / NST);
        mbar_wait(&bar_empty[slot], (round & 1u) ^ 1u);
        unsigned char* st = smem_raw + (size_t)slot * G::STAGE_BYTES;
        const uint32_t bytes = (i >= 2) ? (uint32_t)(G::PT_BYTES + G::PO_BYTES + G::PS_BYTES) : (uint32_t)G::PT_BYTES;
        mbar_expect_tx(&bar_full[slot], bytes);
        if (hints) {
          if (hints & 2) tma_load_3d_hint(st + G::PT_OFF, &tm_p, x0 - G::HX, y0 - 1, z_lo - 1 + i, &bar_full[slot], pol_keep);
          else tma_load_3d(st + G::PT_OFF, &tm_p, x0 - G::HX, y0 - 1, z_lo - 1 + i, &bar_full[slot]);
          if (i >= 2) {
            tma_load_3d_hint(st + G::PO_OFF, &tm_old, x0, y0, z_lo + i - 2, &bar_full[slot], pol_once);
            tma_load_3d_hint(st + G::PS_OFF, &tm_cls, x0, y0, z_lo + i - 2, &bar_full[slot], pol_once);
          }
        } else {
          tma_load_3d(st + G::PT_OFF, &tm_p, x0 - G::HX, y0 - 1, z_lo - 1 + i, &bar_full[slot]);
          if (i >= 2) {
            tma_load_3d(st + G::PO_OFF, &tm_old, x0, y0, z_lo + i - 2, &bar_full[slot]);
            tma_load_3d(st + G::PS_OFF, &tm_cls, x0, y0, z_lo + i - 2, &bar_full[slot]);
          }
        }
      }
    }
    return;
  }

  // --------------------------------- consumers ---------------------------------------------------
  // Shared memory is addressed through 32-bit window addresses kept in registers.  The frequency-independent kernels
  // unroll the plane loop over the NST stages so that every stage offset and barrier parity is a compile-time constant or
  // one XOR; the filter kernels run it rolled, because their unrolled body does not fit the instruction cache
  // (tma_unroll above).
  const int r0 = warp * RPW;                 // first tile row of this warp
  const int xl = 4 * lane;                   // x offset inside the tile
  const int gx = x0 + xl;
  const bool x_ok = gx < X;                  // X is a multiple of 4 on this path
  const int64_t XY = (int64_t)X * Y;
  constexpr uint32_t ES = (uint32_t)sizeof(T);
  constexpr uint32_t ROW = (uint32_t)G::PW * ES;           // bytes per halo-tile row
  const uint32_t sm = smem_u32(smem_raw);
  const uint32_t bf = smem_u32(bar_full), be = smem_u32(bar_empty);
  const uint32_t a_ctr = sm + G::PT_OFF + ((uint32_t)(r0 + 1) * G::PW + G::HX + xl) * ES;   // centre of the warp's first row
  const uint32_t a_old = sm + G::PO_OFF + ((uint32_t)r0 * TX + xl) * ES;
  const uint32_t a_cls = sm + G::PS_OFF + (uint32_t)r0 * TX + xl;
  const uint32_t a_row = sm + G::PT_OFF + (uint32_t)(r0 + 1) * ROW;                          // x = -HX of the first row

  // filter boundaries (one row per warp): entries and prefetched states (update_math.cuh DifRow).  Started before the
  // first wait on the pipeline so that the entry -> state round trips overlap the first planes' TMA loads.
  static_assert(!DIF || RPW == 1, "filter boundaries use the one-row-per-warp tile shapes");
  constexpr int DMO = DIF ? DIF : 1;
  DifRow<T, DMO, WIDE> drow;
  if (DIF) drow.start(dif, z_lo, z_hi, y0 + r0, Y, lane);

  V4<T> down[RPW], cur[RPW], up[RPW];

  // prologue: plane z_lo-1 (centre only), then plane z_lo (kept resident for its xy-neighbours)
  mbar_wait_a(bf, 0);
#pragma unroll
  for (int k = 0; k < RPW; k++) lds4_a(a_ctr + k * ROW, down[k]);
  // stage 0 is NOT handed back here: a shared-memory load that has been issued but not yet performed is not
  // ordered before the mbarrier arrive, so the producer's next TMA write into the stage could overtake it
  // (seen on B200 with two CTAs per SM: `down` picked up bytes of plane z_lo+3).  Every release below comes
  // after a global store that depends on the loaded values, which orders it behind the loads.
  mbar_wait_a(bf + 8 * (1 % NST), (uint32_t)((1 / NST) & 1));
#pragma unroll
  for (int k = 0; k < RPW; k++) lds4_a(a_ctr + (1 % NST) * G::STAGE_BYTES + k * ROW, cur[k]);

  constexpr uint32_t AIR4 = CLS_AIR * 0x01010101u;
  T* out = Pn + (int64_t)z_lo * XY + (int64_t)(y0 + r0) * X + gx;   // this lane's four voxels of the warp's first row
  // The plane loop is unrolled UNR times.  UNR == NST (the default): every stage offset and barrier parity is a
  // compile-time constant or one XOR.  UNR < NST: the stage of plane jb (`base_rt`) is carried in a register and the
  // offsets cost a compare-and-subtract each -- a few integer instructions per plane for a loop body UNR / NST the size,
  // which matters where the unrolled body outgrows the instruction cache (tma_unroll()).
  constexpr int UNR = tma_unroll(sizeof(T), SCHEME, DIF, NST);
  uint32_t par = 0;                                                    // parity of the round plane jb belongs to
  uint32_t base_rt = 0;                                                // stage index of plane jb (UNR < NST only)

  for (int jb = 0; jb < n; jb += UNR) {
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const int j = jb + u;
      if (j >= n) break;
      const uint32_t base = (UNR == NST) ? 0u : base_rt;
      const uint32_t t1 = base + (uint32_t)(u + 1), t2 = base + (uint32_t)(u + 2);
      const uint32_t i1 = t1 >= (uint32_t)NST ? t1 - NST : t1;           // stage of plane z
      const uint32_t w2 = t2 >= (uint32_t)NST ? 1u : 0u;
      const uint32_t i2 = w2 ? t2 - NST : t2;                            // stage of plane z+1 (and P_old / classes of z)
      const uint32_t s2 = i2 * G::STAGE_BYTES, s1 = i1 * G::STAGE_BYTES;
      mbar_wait_a(bf + 8 * i2, par ^ w2);

      V4<T> old[RPW];
      uint32_t pw[RPW];
#pragma unroll
      for (int k = 0; k < RPW; k++) {
        lds4_a(a_ctr + s2 + k * ROW, up[k]);
        lds4_a(a_old + s2 + k * (TX * ES), old[k]);
        pw[k] = lds_u32_a(a_cls + s2 + k * TX);
      }
      V4<T> ym0, yp1;   // y-1 of the warp's first row, y+1 of its last row: from the shared tile of plane z
      lds4_a(a_ctr + s1 - ROW, ym0);
      lds4_a(a_ctr + s1 + RPW * ROW, yp1);

#pragma unroll
      for (int k = 0; k < RPW; k++) {
        const V4<T>& cc = cur[k];
        T xm_edge = __shfl_up_sync(0xffffffffu, cc.v[3], 1);
        T xp_edge = __shfl_down_sync(0xffffffffu, cc.v[0], 1);
        if (lane == 0) xm_edge = lds1_a(a_row + s1 + k * ROW + (G::HX - 1) * ES, (T)0);
        if (lane == 31) xp_edge = lds1_a(a_row + s1 + k * ROW + (G::HX + TX) * ES, (T)0);
        const bool active = x_ok && (y0 + r0 + k) < Y;
        V4<T> res;
        if (active) {
          const V4<T>& ym = (k == 0) ? ym0 : cur[k - 1 < 0 ? 0 : k - 1];
          const V4<T>& yp = (k == RPW - 1) ? yp1 : cur[k + 1 >= RPW ? RPW - 1 : k + 1];
          T S[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            T xm = (q == 0) ? xm_edge : cc.v[q - 1 < 0 ? 0 : q - 1];
            T xp = (q == 3) ? xp_edge : cc.v[q + 1 > 3 ? 3 : q + 1];
            S[q] = (SCHEME == SCH_CENTRED) ? sum6_centred<T>(up[k].v[q], down[k].v[q], yp.v[q], ym.v[q], xp, xm)
                                           : sum6_forward<T>(up[k].v[q], down[k].v[q], yp.v[q], ym.v[q], xp, xm);
          }
          if (pw[k] == AIR4) {
#pragma unroll
            for (int q = 0; q < 4; q++)
              res.v[q] = (SCHEME == SCH_CENTRED) ? voxel_centred_air<T>(cc.v[q], S[q], old[k].v[q], lam2, a_air)
                                                 : voxel_forward_air<T>(cc.v[q], S[q], old[k].v[q], lam2, a_air);
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const ClassEntry<T> ce = class_entry<T, WIDE>(s_table, wide, (pw[k] >> (8 * q)) & 0xffu, (out - Pn) + (int64_t)k * X + q);
              if (SCHEME == SCH_CENTRED) {
                T xm = (q == 0) ? xm_edge : cc.v[q - 1 < 0 ? 0 : q - 1];
                T xp = (q == 3) ? xp_edge : cc.v[q + 1 > 3 ? 3 : q + 1];
                res.v[q] = ((pw[k] >> (8 * q)) & 0xffu) == CLS_AIR
                               ? voxel_centred_air<T>(cc.v[q], S[q], old[k].v[q], lam2, a_air)
                               : voxel_centred_cls<T>(ce, cc.v[q], S[q], up[k].v[q], down[k].v[q], yp.v[q], ym.v[q], xp, xm,
                                                      old[k].v[q], lam2, a_air);
              } else {
                res.v[q] = voxel_forward_cls<T>(ce, cc.v[q], S[q], old[k].v[q], lam2);
              }
            }
          }
        }
        if (DIF) drow.apply(j, res.v, old[k].v, pw[k], active, lane, dif, s_dif, out + (int64_t)k * X, Pn);
        if (active) {
          stg4(out + (int64_t)k * X, res);
          // edge launch (one plane): the same registers go straight into the neighbour slab's halo plane -- a peer-mapped
          // store over NVLink when the neighbour lives on another GPU -- so compute and halo transfer are one launch and the
          // tiles that finish first travel while the others are still being computed
          if (TAIL == 1 && peer != nullptr) stg4(peer + (int64_t)(y0 + r0 + k) * X + gx, res);
        }
      }
      out += XY;
      // every read of stage s1 (plane z tile, plus P_old/node bytes of plane z-1 read last iteration) is done
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_a(be + 8 * i1);
        if (j == 0) mbar_arrive_a(be);   // plane z_lo-1, read in the prologue
      }
      if (DIF) drow.next(dif, j, n, z_lo, z_hi, y0 + r0, Y, lane);
#pragma unroll
      for (int k = 0; k < RPW; k++) { down[k] = cur[k]; cur[k] = up[k]; }
    }
    if (UNR == NST) par ^= 1u;
    else {
      base_rt += UNR;
      if (base_rt >= (uint32_t)NST) { base_rt -= NST; par ^= 1u; }
    }
  }
  // single slab: receivers of this step and sources of the next one, by the CTA that owns their voxel (tma_common.cuh)
  if (TAIL == 2) fused_srcrec<T>(fused_p, fused, Pn, X, Y, TY, z_begin, z_end, chunk, hints, NW * 32);
}

// builds the per-class table with the device arithmetic of update_math.cuh (one thread per class)
template <typename T, int SCHEME>
__global__ void build_class_table_kernel(const uint32_t* __restrict__ keys, int n_classes, UpdConst<T> c, ClassEntry<T>* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_classes) return;
  const uint32_t key = keys[i];
  if (SCHEME == SCH_INTERP) out[i] = make_class_entry_interp<T>(key & 0xffu, (key >> 8) & 0xffu, (key >> 16) & 0xfu, (key >> 20) & 0xfu, c);
  else out[i] = make_class_entry<T, SCHEME>(key & 0xffu, (key >> 8) & 0xffu, c);
}

// ------------------------------------------------------------------------------------------------------
// host side of the TMA path
// ------------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int encode3d(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* base, int X, int Y, int nz, int bx, int by) {
  EncodeTiledFn fn = get_encode_fn();
  PF_CHECK(fn != nullptr, PFDTD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)nz};
  cuuint64_t strides[2] = {(cuuint64_t)X * esize, (cuuint64_t)X * Y * esize};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PF_CHECK(r == CUDA_SUCCESS, PFDTD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) X=%d Y=%d nz=%d box=%dx%d esize=%d", (int)r, X,
           Y, nz, bx, by, esize);
  return PFDTD_OK;
}

// tile variants: index -> (TY, RPW, NST)
struct TileDef { int ty, rpw, nst; const char* name; };
const TileDef kTiles[] = {
    {8, 1, 4, "128x8 r1 s4"},
    {16, 2, 4, "128x16 r2 s4"},
    {16, 1, 4, "128x16 r1 s4"},
    {8, 1, 6, "128x8 r1 s6"},
    {16, 2, 3, "128x16 r2 s3"},
    {32, 2, 3, "128x32 r2 s3"},
    {7, 1, 5, "128x7 r1 s5"},
    {7, 1, 6, "128x7 r1 s6"},
    {8, 1, 5, "128x8 r1 s5"},
};
constexpr int kNumTiles = (int)(sizeof(kTiles) / sizeof(kTiles[0]));

template <typename T, int TY>
constexpr int stage_bytes() { return TileGeom<T, TY>::STAGE_BYTES; }

template <typename T, int SCHEME, int TY, int RPW, int NST, int DIF = 0, bool WIDE = false, int TAIL = 0>
int launch_tma_t(const UpdateArgs& a, const TmaMaps& m, int chunk, int* occupancy_out) {
  auto kern = fdtd_update_tma<T, SCHEME, TY, RPW, NST, DIF, WIDE, TAIL>;
  const int smem = NST * stage_bytes<T, TY>();
  static bool attr_set[64] = {false};   // per device
  int dev = 0;
  PF_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dev] = true;
  }
  const int threads = (TY / RPW + 1) * 32;
  if (occupancy_out) {
    PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occupancy_out, kern, threads, smem));
    return PFDTD_OK;
  }
  const int nplanes = a.z_end - a.z_begin;
  dim3 grid((a.X + TX - 1) / TX, (a.Y + TY - 1) / TY, (nplanes + chunk - 1) / chunk);
  const UpdConst<T> c = make_const<T>(a);
  kern<<<grid, threads, smem, a.stream>>>(m.p_halo, m.p_old, m.cls, (const ClassEntry<T>*)a.class_table, a.n_classes, (T*)a.Pn,
                                          c.lam2, c.a_air, a.X, a.Y, a.z_begin, a.z_end, chunk, a.tma_hints, make_dif<T>(a), make_wide<T>(a),
                                          (a.z_end - a.z_begin == 1) ? (T*)a.peer_plane : nullptr,
                                          a.fused_params, (const FusedSrcRec<T>*)a.fused_srcrec);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

// filter boundaries: one kernel per order on one shape per dtype -- fp32: 128x8 with six stages (72 registers, three
// CTAs per SM); fp64: 128x7 with five (eight warps per CTA leave 128 registers).  Wide meshes (position classes x
// material table) and the launches with a tail (edge planes, fused sources / receivers) run on the same shapes.
template <typename T, int SCHEME, bool WIDE, int TAIL>
int dispatch_order(const UpdateArgs& a, const TmaMaps& m, int chunk, int* occ) {
  constexpr int DTY = sizeof(T) == 4 ? 8 : 7, DNST = sizeof(T) == 4 ? 6 : 5;
  switch (a.dif_order) {
    case 0: return launch_tma_t<T, SCHEME, DTY, 1, DNST, 0, WIDE, TAIL>(a, m, chunk, occ);
    case 1: return launch_tma_t<T, SCHEME, DTY, 1, DNST, 1, WIDE, TAIL>(a, m, chunk, occ);
    case 2: return launch_tma_t<T, SCHEME, DTY, 1, DNST, 2, WIDE, TAIL>(a, m, chunk, occ);
    case 3: return launch_tma_t<T, SCHEME, DTY, 1, DNST, 3, WIDE, TAIL>(a, m, chunk, occ);
    case 4: return launch_tma_t<T, SCHEME, DTY, 1, DNST, 4, WIDE, TAIL>(a, m, chunk, occ);
  }
  set_error("filter order %d is not supported", a.dif_order);
  return PFDTD_ERR_INVALID;
}

template <typename T, int SCHEME>
int dispatch_tile(const UpdateArgs& a, const TmaMaps& m, int tile, int chunk, int* occ) {
  constexpr int DTY = sizeof(T) == 4 ? 8 : 7, DNST = sizeof(T) == 4 ? 6 : 5;
  if (a.tail == 1) return a.wide ? dispatch_order<T, SCHEME, true, 1>(a, m, chunk, occ) : dispatch_order<T, SCHEME, false, 1>(a, m, chunk, occ);
  if (a.tail == 2) {
    PF_CHECK(a.dif_order == 0 && !a.wide, PFDTD_ERR_INVALID, "fused sources / receivers are built for the frequency-independent kernels");
    return launch_tma_t<T, SCHEME, DTY, 1, DNST, 0, false, 2>(a, m, chunk, occ);
  }
  if (a.wide) return dispatch_order<T, SCHEME, true, 0>(a, m, chunk, occ);
  if (a.dif_order > 0) return dispatch_order<T, SCHEME, false, 0>(a, m, chunk, occ);
  switch (tile) {
    case 0: return launch_tma_t<T, SCHEME, 8, 1, 4>(a, m, chunk, occ);
    case 1: return launch_tma_t<T, SCHEME, 16, 2, 4>(a, m, chunk, occ);
    case 2: return launch_tma_t<T, SCHEME, 16, 1, 4>(a, m, chunk, occ);
    case 3: return launch_tma_t<T, SCHEME, 8, 1, 6>(a, m, chunk, occ);
    case 4: return launch_tma_t<T, SCHEME, 16, 2, 3>(a, m, chunk, occ);
    case 5: return launch_tma_t<T, SCHEME, 32, 2, 3>(a, m, chunk, occ);
    case 6: return launch_tma_t<T, SCHEME, 7, 1, 5>(a, m, chunk, occ);
    case 7: return launch_tma_t<T, SCHEME, 7, 1, 6>(a, m, chunk, occ);
    case 8: return launch_tma_t<T, SCHEME, 8, 1, 5>(a, m, chunk, occ);
  }
  set_error("unknown TMA tile variant %d", tile);
  return PFDTD_ERR_INVALID;
}

int dispatch(const UpdateArgs& a, const TmaMaps& m, int tile, int chunk, int* occ) {
  if (a.dtype == PFDTD_F32)
    return a.scheme == SCH_CENTRED ? dispatch_tile<float, SCH_CENTRED>(a, m, tile, chunk, occ)
                                   : dispatch_tile<float, SCH_FORWARD>(a, m, tile, chunk, occ);
  return a.scheme == SCH_CENTRED ? dispatch_tile<double, SCH_CENTRED>(a, m, tile, chunk, occ)
                                 : dispatch_tile<double, SCH_FORWARD>(a, m, tile, chunk, occ);
}

}  // namespace

bool tma_supported(int X, int Y, int dtype) {
  (void)dtype;
  // rows must be 16-byte multiples for the tensor map strides and the 128-bit stores
  return X % 16 == 0 && X >= 16 && Y >= 1 && get_encode_fn() != nullptr;
}

// the one tile shape per dtype / scheme that the filter kernels, wide meshes and launches with a tail are built for
int tma_tail_tile(int dtype, int scheme) { return dtype == PFDTD_F64 ? 6 : (scheme == SCH_INTERP ? 7 : 3); }

const char* tma_tile_name(int dtype, int tile) {
  (void)dtype;
  return (tile >= 0 && tile < kNumTiles) ? kTiles[tile].name : "?";
}

int tma_encode_maps(TmaMaps* out, int dtype, int tile, const void* P, const void* Pold, const uint8_t* cls, int X, int Y, int nz) {
  PF_CHECK(tile >= 0 && tile < kNumTiles, PFDTD_ERR_INVALID, "bad tile variant %d", tile);
  const int ty = kTiles[tile].ty;
  const int esize = dtype == PFDTD_F32 ? 4 : 8;
  const CUtensorMapDataType dt = dtype == PFDTD_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  const int hx = 16 / esize;
  PF_TRY(encode3d(&out->p_halo, dt, esize, P, X, Y, nz, TX + 2 * hx, ty + 2));
  PF_TRY(encode3d(&out->p_old, dt, esize, Pold, X, Y, nz, TX, ty));
  PF_TRY(encode3d(&out->cls, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, cls, X, Y, nz, TX, ty));
  return PFDTD_OK;
}

int tma_pick_config(int dtype, int scheme, int dif_order, bool wide, int X, int Y, int nplanes, int device, int64_t opt_tile,
                    int64_t opt_chunk, TmaConfig* out) {
  // Defaults (B200, profiles/r01_sweep.md): the 128x7 one-row-per-warp shapes -- eight warps per CTA divide evenly
  // over the four sub-partitions, which leaves 64 (fp32, four CTAs per SM) / 128 (fp64, two) registers per thread
  // and room for a fifth or sixth stage.  The filter kernels have one shape per dtype (dispatch_tile).
  int tile;
  if (dif_order > 0 || wide) tile = tma_tail_tile(dtype, scheme);
  else if (opt_tile > 0 && opt_tile <= kNumTiles) tile = (int)opt_tile - 1;
  else tile = (dtype == PFDTD_F32 && scheme == SCH_FORWARD) ? 7 : 6;
  if (Y <= 8 && kTiles[tile].ty > 8) tile = 0;
  UpdateArgs probe{};
  probe.dtype = dtype;
  probe.scheme = scheme;
  probe.dif_order = dif_order;
  probe.wide = wide;
  probe.tail = 0;
  TmaMaps dummy{};
  int occ = 0;
  if (scheme == SCH_INTERP) {
    TmaConfig pc{tile, 1, 0};
    PF_TRY(launch_update_interp_tma(probe, dummy, pc, &occ));
  } else {
    PF_TRY(dispatch(probe, dummy, tile, 1, &occ));
  }
  PF_CHECK(occ >= 1, PFDTD_ERR_CUDA, "TMA kernel variant %d does not fit on an SM", tile);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int64_t resident = (int64_t)occ * sms;
  const int64_t tiles = (int64_t)((X + TX - 1) / TX) * ((Y + kTiles[tile].ty - 1) / kTiles[tile].ty);
  int chunk = 0;
  if (opt_chunk > 0) {
    chunk = (int)opt_chunk;
  } else {
    // Measured on B200 (profiles/r01_sweep.md): when the launch is many waves deep, short chunks win even
    // though every chunk re-reads two planes (those re-reads hit L2, and the fine grain evens out the
    // SMs' finishing times); when the whole launch is only a wave or two, wave quantisation dominates.
    // (fp32 forward: 8-14 planes; fp64, the centred scheme and the 27-point kernels: 20-28)
    const bool light = dtype == PFDTD_F32 && scheme == SCH_FORWARD && dif_order == 0;
    const int lo = light ? 12 : 20, hi = light ? 16 : 28;
    const bool deep = tiles * (int64_t)((nplanes + lo - 1) / lo) >= 4 * resident;
    double best = -1;
    for (int gz = 1; gz <= nplanes; gz++) {
      int ch = (nplanes + gz - 1) / gz;
      if (deep && ch > hi) continue;
      if (ch < (deep ? lo : 2) && gz > 1) break;   // a launch of less than a wave is latency-bound: short chunks fill the machine
      int gz_eff = (nplanes + ch - 1) / ch;
      int64_t total = tiles * gz_eff;
      int64_t waves = (total + resident - 1) / resident;
      double fill = (double)total / (double)(waves * resident);
      double eff = deep ? fill : fill * (double)ch / (double)(ch + 2);
      if (eff > best + 1e-9) { best = eff; chunk = ch; }
    }
    if (chunk <= 0) chunk = nplanes;
  }
  out->tile = tile;
  out->occupancy = occ;
  out->chunk = std::max(1, std::min(chunk, std::max(nplanes, 1)));
  return PFDTD_OK;
}

int build_class_table(const UpdateArgs& a, const uint32_t* d_keys, int n_classes, void* d_table) {
  if (n_classes <= 0) return PFDTD_OK;
  const int th = 64, bl = (n_classes + th - 1) / th;
  if (a.scheme == SCH_INTERP) {
    if (a.dtype == PFDTD_F32) build_class_table_kernel<float, SCH_INTERP><<<bl, th, 0, a.stream>>>(d_keys, n_classes, make_const<float>(a), (ClassEntry<float>*)d_table);
    else build_class_table_kernel<double, SCH_INTERP><<<bl, th, 0, a.stream>>>(d_keys, n_classes, make_const<double>(a), (ClassEntry<double>*)d_table);
    PF_CUDA(cudaGetLastError());
    return PFDTD_OK;
  }
  if (a.dtype == PFDTD_F32) {
    if (a.scheme == SCH_CENTRED) build_class_table_kernel<float, SCH_CENTRED><<<bl, th, 0, a.stream>>>(d_keys, n_classes, make_const<float>(a), (ClassEntry<float>*)d_table);
    else build_class_table_kernel<float, SCH_FORWARD><<<bl, th, 0, a.stream>>>(d_keys, n_classes, make_const<float>(a), (ClassEntry<float>*)d_table);
  } else {
    if (a.scheme == SCH_CENTRED) build_class_table_kernel<double, SCH_CENTRED><<<bl, th, 0, a.stream>>>(d_keys, n_classes, make_const<double>(a), (ClassEntry<double>*)d_table);
    else build_class_table_kernel<double, SCH_FORWARD><<<bl, th, 0, a.stream>>>(d_keys, n_classes, make_const<double>(a), (ClassEntry<double>*)d_table);
  }
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

template <typename T, int SCHEME>
__global__ void build_dif_table_kernel(const uint32_t* __restrict__ keys, int n, UpdConst<T> c, int order, DifEntry<T>* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t key = keys[i];
  out[i] = make_dif_entry<T, SCHEME>(key & 0xffu, (key >> 8) & 0xffu, c, order);
}

int build_dif_table(const UpdateArgs& a, const uint32_t* d_keys, int n_dif, void* d_table) {
  if (n_dif <= 0) return PFDTD_OK;
  const int th = 64, bl = (n_dif + th - 1) / th;
  if (a.dtype == PFDTD_F32) {
    if (a.scheme == SCH_CENTRED) build_dif_table_kernel<float, SCH_CENTRED><<<bl, th, 0, a.stream>>>(d_keys, n_dif, make_const<float>(a), a.dif_order, (DifEntry<float>*)d_table);
    else build_dif_table_kernel<float, SCH_FORWARD><<<bl, th, 0, a.stream>>>(d_keys, n_dif, make_const<float>(a), a.dif_order, (DifEntry<float>*)d_table);
  } else {
    if (a.scheme == SCH_CENTRED) build_dif_table_kernel<double, SCH_CENTRED><<<bl, th, 0, a.stream>>>(d_keys, n_dif, make_const<double>(a), a.dif_order, (DifEntry<double>*)d_table);
    else build_dif_table_kernel<double, SCH_FORWARD><<<bl, th, 0, a.stream>>>(d_keys, n_dif, make_const<double>(a), a.dif_order, (DifEntry<double>*)d_table);
  }
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

size_t dif_entry_bytes(int dtype) { return dtype == PFDTD_F32 ? sizeof(DifEntry<float>) : sizeof(DifEntry<double>); }

size_t class_entry_bytes(int dtype) { return dtype == PFDTD_F32 ? sizeof(ClassEntry<float>) : sizeof(ClassEntry<double>); }

int launch_update_tma(const UpdateArgs& a, const TmaMaps& maps, const TmaConfig& cfg) {
  if (a.z_end <= a.z_begin) return PFDTD_OK;
  return dispatch(a, maps, cfg.tile, cfg.chunk, nullptr);
}

}  // namespace pfdtd
