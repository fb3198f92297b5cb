// tma_common.cuh -- device helpers shared by the TMA z-march kernels (update_kernels.cu: 7-point SRL,
// interp_kernels.cu: 27-point IISO/IWB): mbarrier + cp.async.bulk.tensor wrappers, 128-bit shared/global
// vector access, and the shared-memory stage geometry.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "pfdtd_internal.h"

namespace pfdtd {
namespace {

constexpr int TX = 128;  // voxels per tile row: 32 lanes x 4 voxels

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], "
      "[%5], %6;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---- the same operations on 32-bit shared-window addresses (hot loops keep integer addresses in registers
// instead of re-deriving them from generic pointers every iteration) ----
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

template <typename T> struct V4 { T v[4]; };

__device__ __forceinline__ void lds4_a(uint32_t a, V4<float>& o) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.v[0]), "=f"(o.v[1]), "=f"(o.v[2]), "=f"(o.v[3]) : "r"(a));
}
__device__ __forceinline__ void lds4_a(uint32_t a, V4<double>& o) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o.v[0]), "=d"(o.v[1]) : "r"(a));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(o.v[2]), "=d"(o.v[3]) : "r"(a));
}
__device__ __forceinline__ float lds1_a(uint32_t a, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds1_a(uint32_t a, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32_a(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

__device__ __forceinline__ void lds4(const float* p, V4<float>& o) {
  float4 t = *reinterpret_cast<const float4*>(p);
  o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[3] = t.w;
}
__device__ __forceinline__ void lds4(const double* p, V4<double>& o) {
  double2 a = *reinterpret_cast<const double2*>(p);
  double2 b = *reinterpret_cast<const double2*>(p + 2);
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = b.x; o.v[3] = b.y;
}
__device__ __forceinline__ void ldg4(const float* p, V4<float>& o) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[3] = t.w;
}
__device__ __forceinline__ void ldg4(const double* p, V4<double>& o) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = b.x; o.v[3] = b.y;
}
__device__ __forceinline__ void stg4(float* p, const V4<float>& o) {
  *reinterpret_cast<float4*>(p) = make_float4(o.v[0], o.v[1], o.v[2], o.v[3]);
}
__device__ __forceinline__ void stg4(double* p, const V4<double>& o) {
  *reinterpret_cast<double2*>(p) = make_double2(o.v[0], o.v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(o.v[2], o.v[3]);
}

// ---- halo hand-over between processes (one process per GPU) ------------------------------------------------------
// The edge launch of a slab stores its plane into the neighbour process's halo plane through a CUDA-IPC mapping
// (NVLink).  Ordering between the two processes goes through one word per direction in the RECEIVER's memory:
// the sender publishes "step q done" with a one-thread launch right behind its edge launches, the receiver's next
// step waits for it (both in srcrec_kernels.cu).  Flag block of a process (ints): [0] published by the lower
// neighbour, [1] by the upper one, [2] / [3] steps this process has published towards the lower / upper neighbour,
// [6] set when a wait timed out.  (A first version published from inside the edge launch -- last CTA done -> flag; its
// system-scope fence per CTA kept every edge CTA resident for a link round trip and tripled the edge launches' time.)
enum : int { HALO_FROM_LO = 0, HALO_FROM_HI = 1, HALO_SEQ = 2, HALO_CTR = 4, HALO_ERR = 6, HALO_FLAG_INTS = 8 };

// CTA -> (tile row, z chunk): the natural order.  (A remapped order -- the first / last tile rows of every chunk, whose warps
// pay a filter-state round trip per plane, launched first -- bought nothing, 0 ... -1.5 %, and the two remapped values kept
// alive across the plane loop cost fp64 IISO + DIF 2 a spill and 8 % at its 128-register budget: profiles/r02_dif_ab.md.)
__device__ __forceinline__ void cta_tile(int& by, int& bz) {
  by = blockIdx.y;
  bz = blockIdx.z;
}

// ---- sources and receivers inside the update launch (single slab) ---------------------------------------------------
// The reference's step is source(n) -> update -> receiver(n) (kernels3d.cu:93-104,164-173); receiver(n) and
// source(n+1) both act on the field the update has just written.  So the CTA that computed a receiver's voxel records
// it, and the CTA that computed a source's voxel overwrites (or adds to) what it stored with the next step's sample:
// no second launch per step.  Every CTA compares its tile with the item coordinates in the kernel parameters (no memory
// access); the few owners do the rest.  Each item carries its own step counter, advanced by its owner, so there is no
// grid-wide hand-over.  Out of line and fed only by kernel parameters: the tile is recomputed here instead of being
// kept alive across the march (every value that lives across the plane loop costs a register the loop does not have).
template <typename T>
__device__ __noinline__ void fused_srcrec(const FusedParams& fp, const FusedSrcRec<T>* __restrict__ sr, T* __restrict__ Pn, int X, int Y,
                                          int ty, int z_begin, int z_end, int chunk, int hints, int consumer_threads) {
  int cby, cbz;
  cta_tile(cby, cbz);
  const int x0 = blockIdx.x * TX, y0 = cby * ty;
  const int z_lo = z_begin + cbz * chunk, z_hi = min(z_lo + chunk, z_end);
  const int n_items = fp.n_src + fp.n_rec;
  unsigned mine = 0u;                                 // CTA-uniform: bit i = item i lies in this CTA's tile
  for (int i = 0; i < n_items; i++) {
    const int x = fp.xyz[i][0], y = fp.xyz[i][1], z = fp.xyz[i][2];
    if (x >= x0 && x < x0 + TX && y >= y0 && y < y0 + ty && z >= z_lo && z < z_hi) mine |= 1u << i;
  }
  if (mine == 0u) return;
  asm volatile("bar.sync 1, %0;" ::"r"(consumer_threads) : "memory");   // every store of this CTA's tile has been issued
  if (threadIdx.x != 0) return;
  const int64_t XY = (int64_t)X * Y;
  const int first_recordable = sr->d_step[1], last = sr->d_step[2];
  for (int i = fp.n_src; i < n_items; i++)            // receivers first: they see the field before the next injection
    if (mine & (1u << i)) {
      const int n = sr->item_step[i];
      if (n >= first_recordable && (long long)n < sr->rec_stride)
        sr->rec_out[(long long)sr->slot[i] * sr->rec_stride + n] = Pn[(int64_t)fp.xyz[i][2] * XY + (int64_t)fp.xyz[i][1] * X + fp.xyz[i][0]];
      sr->item_step[i] = n + 1;
    }
  for (int i = 0; i < fp.n_src; i++)                  // in source order: a later source at the same voxel wins, like the reference's loop
    if (mine & (1u << i)) {
      const int n = sr->item_step[i];
      if (n < last && (long long)(n + 1) < sr->src_stride) {
        T* q = Pn + (int64_t)fp.xyz[i][2] * XY + (int64_t)fp.xyz[i][1] * X + fp.xyz[i][0];
        const T v = sr->src_samples[(long long)sr->slot[i] * sr->src_stride + n + 1];
        if (sr->type[i] == 0 /* PFDTD_SRC_HARD */ || !sr->soft_accumulate) *q = v;
        else *q += v;
      }
      sr->item_step[i] = n + 1;
    }
}

constexpr int align128(int x) { return (x + 127) & ~127; }

template <typename T, int TY>
struct TileGeom {
  static constexpr int HX = 16 / (int)sizeof(T);           // x halo columns each side (16 B keeps rows 16-B aligned)
  static constexpr int PW = TX + 2 * HX;                    // halo tile row pitch (elements)
  static constexpr int PT_BYTES = (TY + 2) * PW * (int)sizeof(T);
  static constexpr int PO_BYTES = TY * TX * (int)sizeof(T);
  static constexpr int PS_BYTES = TY * TX;
  static constexpr int PT_OFF = 0;
  static constexpr int PO_OFF = align128(PT_BYTES);
  static constexpr int PS_OFF = PO_OFF + align128(PO_BYTES);
  static constexpr int STAGE_BYTES = PS_OFF + align128(PS_BYTES);
};


}  // namespace
}  // namespace pfdtd
