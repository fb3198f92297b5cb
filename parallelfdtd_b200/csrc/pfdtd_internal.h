// pfdtd_internal.h -- internal declarations shared by the translation units of libpfdtd_b200.so
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/pfdtd.h"

namespace pfdtd {

void set_error(const char* fmt, ...);

// equation families: the reference's SRL_FORWARD and SHARED share the forward-difference boundary
// (kernels3d.cu:485-606), SRL uses the centred-difference boundary (:608-665)
// SCH_INTERP: the 27-point compact explicit family (IISO, IWB) with the forward-difference boundary
enum : int { SCH_FORWARD = 0, SCH_CENTRED = 2, SCH_INTERP = 3 };

#define PF_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      pfdtd::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return PFDTD_ERR_CUDA;                                                                   \
    }                                                                                          \
  } while (0)

#define PF_CHECK(cond, code, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      pfdtd::set_error(__VA_ARGS__);       \
      return (code);                       \
    }                                      \
  } while (0)

#define PF_TRY(expr)              \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != PFDTD_OK) return rc__; \
  } while (0)

// ---- device block cache (pfdtd_api.cu) -----------------------------------------------------------
// The large device arrays of a solver (node volumes, pressure fields, filter states) are taken from / returned to a
// per-device cache of blocks of exactly the sizes seen before, so that a process that runs one simulation after another
// (a MATLAB or Python session) pays the driver's cudaMalloc / cudaFree -- milliseconds to tens of milliseconds per
// 512 MB block on a busy host, profiles/r02_final_setup_trace.log -- once.  Semantics of the driver calls are kept:
// dev_free synchronises the block's device before the block can be handed out again (cudaFree's implicit
// synchronisation), a reused block is zero-filled, and an allocation that fails empties the cache and retries.
// Blocks under 1 MiB, and pointers that did not come from dev_alloc (volumes a caller adopts out to the library), go
// straight to the driver.  PFDTD_CACHE_MB = most cached megabytes per device (default 16384; 0 = no cache).
int dev_alloc(void** d_ptr, size_t bytes);      // on the calling thread's current device
void dev_free(void* d_ptr);                     // nullptr is fine
void dev_cache_forget(void* d_ptr);             // this block is not to be recycled: dev_free will hand it to cudaFree
void dev_cache_release(int device);             // cudaFree everything cached for `device` (-1: every device)
size_t dev_cache_bytes(int device);

// ---- mesh preparation kernels (mesh_kernels.cu) ------------------------------------------------
// padWithZeros + toBilbao/toKowalczyk + calcBoundaries of the reference, on `stream`.
int launch_pad_with_zeros(const uint8_t* d_old, uint8_t* d_new, uint32_t dx, uint32_t dy, uint32_t dz,
                          uint32_t nx, uint32_t ny, uint32_t nz, int skip_z0, cudaStream_t stream);
int launch_translate_nodes(uint8_t* d_pos, uint8_t* d_mat, uint64_t n, int centred, unsigned long long* d_counts2,
                           cudaStream_t stream);
// padWithZeros of both volumes + translation + counts in one pass (16 bytes per thread when nx % 16 == 0)
int launch_prepare_nodes(const uint8_t* d_old_bid, const uint8_t* d_old_mat, uint8_t* d_pos, uint8_t* d_mat, uint32_t dx, uint32_t dy,
                         uint32_t dz, uint32_t nx, uint32_t ny, uint32_t nz, int skip_z0, int centred, unsigned long long* d_counts2,
                         cudaStream_t stream);
// node classes: collect the distinct node keys (pos | mat << 8 | K12 << 16 | K8 << 20) in an open-addressing
// table (d_table: cap slots preset to 0xffffffff), then write the class byte volume through it
int launch_mark_classes(const uint8_t* d_pos, const uint8_t* d_mat, uint64_t n, uint32_t air_key, uint32_t air_code, int interp,
                        uint32_t X, uint32_t Y, uint32_t Z, uint32_t* d_table, uint32_t cap, uint32_t* d_count, cudaStream_t stream);
int launch_assign_classes(const uint8_t* d_pos, const uint8_t* d_mat, uint64_t n, uint32_t air_key, uint32_t air_code, int interp,
                          uint32_t X, uint32_t Y, uint32_t Z, const uint32_t* d_table, const uint8_t* d_ids, uint32_t cap,
                          uint8_t* d_cls, cudaStream_t stream);

// ---- sources / receivers handled inside the update launch (tma_common.cuh fused_srcrec) ------------------------------
// The voxel coordinates travel as a kernel parameter (constant bank): "does my tile own one of them" costs a CTA a few
// compares and no memory access; only the owning CTAs touch the descriptor in device memory.
#define PFDTD_FUSED_MAX 16
struct FusedParams {
  int n_src, n_rec;                 // items [0, n_src) are sources, [n_src, n_src + n_rec) receivers; 0 / 0 = off
  int xyz[PFDTD_FUSED_MAX][3];      // voxel coordinates local to the slab
};
template <typename T>
struct FusedSrcRec {                // device memory, read by owning CTAs only
  int soft_accumulate, pad;
  long long rec_stride, src_stride;
  T* rec_out;                       // [n_rec_total][rec_stride]
  const T* src_samples;             // [n_src_total][src_stride]
  const int* d_step;                // [1] first recordable step, [2] last step of the enqueue
  int* item_step;                   // [n_src + n_rec] step each item is at (advanced by its owner CTA)
  int slot[PFDTD_FUSED_MAX];        // row in src_samples / rec_out
  int type[PFDTD_FUSED_MAX];        // source type
};

// ---- update kernels (update_kernels.cu) ---------------------------------------------------------
struct TmaMaps {           // tensor maps of one partition for one (cur,new) buffer assignment
  CUtensorMap p_halo;      // P (current field), box with x/y halo
  CUtensorMap p_old;       // field that is overwritten (past -> next), box without halo
  CUtensorMap cls;         // node class byte, box without halo
};

struct UpdateArgs {
  int dtype;               // PFDTD_F32 / PFDTD_F64
  int scheme;              // 0 forward equations, 2 centred equations
  const uint8_t* pos;
  const uint8_t* mat;
  const uint8_t* cls;      // node class byte volume (TMA kernel)
  const void* class_table; // device ClassEntry<T>[n_classes]
  int n_classes;
  // wide meshes (update_math.cuh WideArgs): per (lossy position class, material) tables in global memory; null = narrow
  bool wide;
  const void* wide_class_table;  // ClassEntry<T>[n_lossy][n_unique]
  const void* wide_dif_table;    // DifEntry<T>[n_lossy][n_unique]
  int tma_hints;           // bit0: evict_first on once-read operands, bit1: evict_last on P
  const void* P;           // current field (slab base)
  void* Pn;                // past field, overwritten with the next field
  const void* materials;   // device [n_coefs]
  uint32_t n_coefs;
  double params[4];        // lambda, lambda^2, 1/3, octave  (already rounded to the dtype)
  double dcoef[4];         // SCH_INTERP: d1 (axial), d2 (edge), d3 (corner), d4 (centre) of the compact scheme
  int matidx_as_written;
  // digital impedance filters (0 = off)
  int dif_order;
  void* dif_state;              // [dif_nb][dif_state_pad(order)] of the dtype
  const uint32_t* dif_rowbase;  // [nz][Y][ceil(X/128)][2]: index of the segment's first filter voxel, flags (update_math.cuh)
  const void* dif_table;        // DifEntry<T>[n_dif]
  uint32_t dif_nb;
  uint32_t dif_lo;
  int n_dif;
  int X, Y;
  int z_begin, z_end;      // local planes [z_begin, z_end) are updated
  // edge launches of a slab (exactly one plane): the freshly computed plane is also stored into this plane of the
  // neighbour slab -- its halo plane, on another GPU over NVLink when peer access exists -- by the same kernel
  void* peer_plane;
  // what the CTAs do after their march: 0 nothing, 1 edge launch (peer_plane), 2 fused sources / receivers.
  // Launches with a tail run on the tile shape of tma_tail_tile() and need tensor maps encoded for it.
  int tail;
  FusedParams fused_params;
  const void* fused_srcrec;     // device FusedSrcRec<T>
  cudaStream_t stream;
};

enum { KERNEL_AUTO = 0, KERNEL_TMA = 1, KERNEL_PLAIN = 2 };

struct TmaConfig { int tile; int chunk; int occupancy = 0; /* resident CTAs per SM (occupancy API) */ };

bool tma_supported(int X, int Y, int dtype);
int tma_encode_maps(TmaMaps* out, int dtype, int tile, const void* P, const void* Pold, const uint8_t* cls, int X, int Y, int nz);
int build_class_table(const UpdateArgs& a, const uint32_t* d_keys, int n_classes, void* d_table);
size_t class_entry_bytes(int dtype);
size_t dif_entry_bytes(int dtype);
int build_dif_table(const UpdateArgs& a, const uint32_t* d_keys /* keys of the lossy classes */, int n_dif, void* d_table);
// per 128-voxel row segment: number of voxels whose class id is >= dif_lo (bits 0..7) and the x offset of the
// first of them inside the segment (bits 8..14), planes [1, nz-1) only
inline int dif_state_pad(int order) { return order <= 1 ? 1 : (order == 2 ? 2 : 4); }
int launch_count_dif_segments(const uint8_t* d_cls, int X, int Y, int nz, uint32_t dif_lo, uint32_t* d_counts, cudaStream_t stream);
// counts [nz][n_cols] (launch_count_dif_segments) -> entries [nz][n_cols][2], total number of filter voxels to *d_nb;
// d_col_scratch: n_cols words
int launch_build_dif_entries(const uint32_t* d_counts, int n_cols, int nz, uint32_t* d_col_scratch, unsigned long long* d_nb,
                             uint32_t* d_entries, cudaStream_t stream);
int tma_pick_config(int dtype, int scheme, int dif_order, bool wide, int X, int Y, int nplanes, int device, int64_t opt_tile,
                    int64_t opt_chunk, TmaConfig* out);
int launch_update_tma(const UpdateArgs& a, const TmaMaps& maps, const TmaConfig& cfg);
int launch_update_plain(const UpdateArgs& a);
// 27-point kernels (interp_kernels.cu); the TMA variant shares TmaMaps / TmaConfig with the 7-point one
int launch_update_interp_tma(const UpdateArgs& a, const TmaMaps& maps, const TmaConfig& cfg, int* occupancy_out);
int launch_update_interp_plain(const UpdateArgs& a);
const char* tma_tile_name(int dtype, int tile);
int tma_tail_tile(int dtype, int scheme);

// ---- capture kernel (capture_kernels.cu): planes [z_lo, z_lo+nz) of a partition, orientation 1 (xz) or 2 (yz)
int launch_capture_slice(int dtype, const void* P, const uint8_t* pos, void* out_p, uint8_t* out_pos, uint32_t X, uint32_t Y,
                         uint32_t z_lo, uint32_t nz, uint32_t slice, int orientation, cudaStream_t stream);

// ---- voxeliser (voxelize_kernels.cu): closed triangle mesh -> device `bid` / material volumes (cudaMalloc'ed, with the
// slice + row + 1 bytes of slack the reference gives them), dims = ceil(max / dx) + 3 per axis
int voxelize_to_device(int device, const float* h_vertices, uint32_t n_vertices, const uint32_t* h_indices, uint32_t n_triangles,
                       const uint8_t* h_tri_material, float dx, uint8_t** d_bid_out, uint8_t** d_mat_out, uint32_t* vx_out, uint32_t* vy_out,
                       uint32_t* vz_out, uint64_t* launches);

// ---- source / receiver kernel (srcrec_kernels.cu) -------------------------------------------------
struct SrcRecArgs {
  int dtype;
  void* P;                       // current field of the partition
  int n_rec; const int64_t* rec_elem; const int32_t* rec_slot; void* rec_out; int64_t rec_stride;
  int n_src; const int64_t* src_elem; const int32_t* src_type; const int32_t* src_slot; const void* src_samples; int64_t src_stride;
  int* d_step;                   // device step counter n: record slot n-1 (if n>0 && do_record), inject sample n (if do_inject)
  int do_record, do_inject, soft_accumulate, advance;
  // wait, before anything else, until the neighbour processes have delivered the halo planes of the previous step
  const int* halo_flags;         // this process's flag block (tma_common.cuh) or null
  int wait_lo, wait_hi;
  cudaStream_t stream;
};
int launch_srcrec(const SrcRecArgs& a);
// After the edge launches of a step (same stream): tell the neighbour PROCESSES that their halo planes hold this step
// (tma_common.cuh: flag block layout).  remote_lo / remote_hi: the neighbours' flag words, either may be null.
int launch_halo_publish(int* local_flags, int* remote_lo, int* remote_hi, cudaStream_t stream);

}  // namespace pfdtd
