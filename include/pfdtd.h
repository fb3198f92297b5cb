/* =============================================================================
 * pfdtd.h -- C ABI of libpfdtd_b200.so: the B200-native replacement for the
 * time-stepping hot path of juuli/ParallelFDTD.
 *
 * Every entry point names the reference interface (file:line relative to the reference
 * root) it replaces.  Plain pointers and sizes only; no C++ or torch types.  All functions
 * return 0 on success and a non-zero PFDTD_ERR_* code on failure, with a human-readable
 * message available from pfdtd_last_error() (thread-local).  There is no CPU fallback: every
 * compute call needs a CUDA device and fails loudly without one.
 *
 * Layout contract (bit-exact with the reference): node volumes are uint8 arrays indexed
 * z*dimX*dimY + y*dimX + x (src/kernels/cudaMesh.h:247-249) after padding dimX,dimY,dimZ up to
 * multiples of the block size (src/kernels/cudaMesh.cu:253-304); partitions are z-slabs with
 * one-slice halos (src/kernels/cudaMesh.h:280-307).
 * ============================================================================= */
#ifndef PFDTD_H_
#define PFDTD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfdtd_solver pfdtd_solver;

/* precision of the pressure fields (reference: CudaMesh::double_, cudaMesh.h:87) */
enum { PFDTD_F32 = 0, PFDTD_F64 = 1 };

/* update scheme; 0..2 are the reference's `enum UpdateType {SRL_FORWARD, SHARED, SRL}`
 * (src/base/SimulationParameters.h:39); the interpolated compact schemes are appended so the
 * numeric values used by MEX / Python callers keep their meaning. */
enum { PFDTD_SRL_FORWARD = 0, PFDTD_SHARED = 1, PFDTD_SRL = 2, PFDTD_IISO = 3, PFDTD_IWB = 4 };

/* reference `enum SrcType {SRC_HARD, SRC_SOFT, SRC_TRANSPARENT}` (src/base/SrcRec.h:33) */
enum { PFDTD_SRC_HARD = 0, PFDTD_SRC_SOFT = 1, PFDTD_SRC_TRANSPARENT = 2 };

enum {
  PFDTD_OK = 0,
  PFDTD_ERR_INVALID = 1,   /* bad argument / call order */
  PFDTD_ERR_CUDA = 2,      /* a CUDA call failed (reference: cudasafe -> throw(-1), cudaUtils.h:47-51) */
  PFDTD_ERR_NO_DEVICE = 3, /* no usable CUDA device */
  PFDTD_ERR_RANGE = 4,     /* index out of range (reference: std::out_of_range from .at()) */
  PFDTD_ERR_INTERRUPTED = 5,
  PFDTD_ERR_COMM = 6       /* inter-process halo transport failed */
};

/* options for pfdtd_set_option */
enum {
  /* material index arithmetic of the forward kernel: 1 = as written in the reference
   * (`mat*20*+octave`, src/kernels/kernels3d.cu:513), 0 = intended `mat*20+octave`.  Default 1. */
  PFDTD_OPT_MATIDX_AS_WRITTEN = 1,
  /* SOFT/TRANSPARENT injection: 0 = as written (addSample overwrites, cudaMesh.h:362-366),
   * 1 = accumulate (what cudaMesh.h:497-508 documents).  Default 0. */
  PFDTD_OPT_SOFT_ACCUMULATE = 2,
  /* update kernel: 0 = auto, 1 = TMA z-march, 2 = plain per-voxel kernel.  Default 0. */
  PFDTD_OPT_KERNEL = 3,
  /* the volume handed to pfdtd_setup_mesh is global slices [first, first+vz) of a taller
   * domain owned by several processes: first global slice / global slice count. */
  PFDTD_OPT_GLOBAL_Z_FIRST = 4,
  PFDTD_OPT_GLOBAL_Z_DIM = 5,
  /* 1 = pad y with block.x in double precision like setupMeshDouble (cudaMesh.cu:106-111). Default 0. */
  PFDTD_OPT_DOUBLE_PAD_AS_WRITTEN = 6,
  /* capture the step loop of pfdtd_run in a CUDA graph (default 1) */
  PFDTD_OPT_USE_GRAPH = 7,
  /* overlap halo exchange with interior compute (default 1); 0 = exchange after the full update */
  PFDTD_OPT_OVERLAP = 8,
  /* z-planes per CTA chunk of the TMA kernel, 0 = auto */
  PFDTD_OPT_TMA_CHUNK = 9,
  /* tile variant of the TMA kernel, 0 = auto */
  PFDTD_OPT_TMA_TILE = 10,
  /* bracket every update-kernel launch of pfdtd_enqueue_steps with CUDA events so that
   * pfdtd_last_timing can report the kernels' own device time (disables graph replay) */
  PFDTD_OPT_TIME_KERNELS = 11,
  /* L2 cache hints on the TMA loads: bit0 = evict_first for operands read once per step (P_old,
   * node bytes), bit1 = evict_last for P (shared with neighbouring tiles).  Default 0. */
  PFDTD_OPT_TMA_HINTS = 12,
  /* Frequency-dependent boundaries (digital impedance filters; not in the reference -- its `dif_` members are a
   * stub, cudaMesh.h:88,122-123).  N > 0 (<= 4): row m of the material table is read as the IIR admittance of
   * material m, [b0 .. bN, a1 .. aN] (a0 = 1), the octave index is ignored, and every boundary node carries N
   * filter states updated in the same kernel pass.  N = 0 (default): the reference's scalar admittance
   * materials[m*20 + octave].  Order-0 filters reproduce that path bit for bit.  Set before pfdtd_make_partition. */
  PFDTD_OPT_DIF_ORDER = 13,
  /* slabs in one process: the edge launches store their plane straight into the neighbour slab's halo plane
   * (peer-mapped stores; default 1).  0 = copy the planes with cudaMemcpyPeerAsync after the edge launches.
   * One process per GPU: the same through CUDA-IPC mappings (pfdtd_comm_init); 0 = ncclSend/ncclRecv. */
  PFDTD_OPT_PEER_STORES = 14,
  /* single slab: record the receivers and inject the next step's sources inside the update launch instead of a
   * separate launch per step (the reference does both with per-element memcpys, kernels3d.cu:93-104,164-173).
   * 1 (default): where the step is launch-bound (slabs up to 2^24 voxels, frequency-independent boundaries, at most 16
   * sources + receivers); 2: for any slab size; 0: always the separate launch.  Same results either way. */
  PFDTD_OPT_FUSE_SRCREC = 15
};

typedef int (*pfdtd_interrupt_cb)(void);                      /* kernels3d.h: bool (*)(void) */
typedef void (*pfdtd_progress_cb)(int step, int max_step, float seconds_per_step);

/* ---- library ------------------------------------------------------------- */
const char* pfdtd_last_error(void);
const char* pfdtd_version(void);
/* App::queryDevices / initializeDevices (src/App.cpp:91-131) */
int pfdtd_device_count(int* out_count);
int pfdtd_device_mem_mb(int device, int* out_total_mb, int* out_free_mb);

/* ---- device memory helpers (src/kernels/cudaUtils.h:59-171: toDevice / valueToDevice / fromDevice / copyHostToDevice /
 * copyDeviceToHost / destroyMem / getCurrentDevice) -------------------------------------------------------------
 * The reference's callers (its tests, the voxelizer wrapper) create the `bid` / material volumes with these helpers and
 * hand them to CudaMesh::setupMesh, which adopts them (pfdtd_setup_mesh_device).  `device` -1 = the calling thread's
 * current device; otherwise the device is made current, as the reference's helpers do.  Blocking, like the
 * reference's. */
int pfdtd_current_device(int* device);
int pfdtd_device_alloc(int device, size_t bytes, void** d_ptr);
/* valueToDevice: `count` elements of `elem_size` (1, 2, 4 or 8) bytes set to *value, written on the device */
int pfdtd_device_fill(int device, void* d_ptr, size_t count, size_t elem_size, const void* value);
int pfdtd_device_upload(int device, void* d_dst, const void* h_src, size_t bytes);
int pfdtd_device_download(int device, void* h_dst, const void* d_src, size_t bytes);
int pfdtd_device_free(int device, void* d_ptr);
/* The library keeps the large device blocks of a solver it has destroyed (node volumes, pressure fields, filter states; at most
 * PFDTD_CACHE_MB megabytes per device, default 16384, 0 = keep nothing) and builds the next solver of the same size out of them
 * instead of going through cudaFree / cudaMalloc again.  pfdtd_device_mem_mb counts them as free; an allocation that does
 * not fit gives them back by itself.  This call returns them to the driver now (`device` -1: every device) -- what is left
 * of the reference's cudaDeviceReset before a run (matlab/device_reset.cpp:5-17, App::resetDevices src/App.cpp:114-120). */
int pfdtd_release_cached_memory(int device);

/* ---- solver lifetime ----------------------------------------------------- */
int pfdtd_create(pfdtd_solver** out);
/* CudaMesh::destroyPartitions (src/kernels/cudaMesh.h:160-182) + receiver buffers */
int pfdtd_destroy(pfdtd_solver* s);
int pfdtd_set_option(pfdtd_solver* s, int option, int64_t value);
int pfdtd_get_option(pfdtd_solver* s, int option, int64_t* value);
/* Interpolated schemes (PFDTD_IISO / PFDTD_IWB; not in the reference): override the four weights of the
 * 27-point compact explicit update, d[0] axial, d[1] edge, d[2] corner, d[3] centre; NULL restores the
 * scheme's own values.  (lam2, 0, 0, 2 - 6 lam2) is the reference's SRL_FORWARD equation.  Call before
 * pfdtd_setup_mesh. */
int pfdtd_set_scheme_coefficients(pfdtd_solver* s, const double* d4);

/* ---- mesh setup ---------------------------------------------------------- */
/* CudaMesh::setupMesh / setupMeshDouble (src/kernels/cudaMesh.cu:26-153): pad the voxelizer-style
 * `bid` (0..27) and material volumes to block multiples (padWithZeros, :253-326), translate to the
 * scheme's node byte (toBilbao :328-361 for element_type 0,1,3 / toKowalczyk :363-480 otherwise),
 * count air/boundary nodes (calcBoundaries :500-514).  HOST pointers; volumes are [vz][vy][vx].
 * `params` = 4 values [lambda, lambda^2, 1/3, octave] and `material_coefs` = [n_unique][20]
 * admittances, both of the solver dtype (SimulationParameters.cpp:379-396,
 * MaterialHandler.cpp:100-164); they are copied, not borrowed. */
int pfdtd_setup_mesh(pfdtd_solver* s, const uint8_t* h_bid, const uint8_t* h_mat,
                     uint32_t vx, uint32_t vy, uint32_t vz,
                     uint32_t block_x, uint32_t block_y, uint32_t block_z,
                     uint32_t element_type, int dtype,
                     const void* params, const void* material_coefs, uint32_t n_unique_materials);
/* same, but the two volumes are DEVICE pointers on `device` (-1 = the calling thread's current
 * device) and are adopted (freed by the library), exactly like the reference's setupMesh arguments
 * (cudaMesh.cu:290-291). */
int pfdtd_setup_mesh_device(pfdtd_solver* s, int device, uint8_t* d_bid, uint8_t* d_mat,
                            uint32_t vx, uint32_t vy, uint32_t vz,
                            uint32_t block_x, uint32_t block_y, uint32_t block_z,
                            uint32_t element_type, int dtype,
                            const void* params, const void* material_coefs, uint32_t n_unique_materials);
/* ---- voxelisation (the step before setupMesh in App::initializeMesh, src/App.cpp:181-190) ----
 * voxelizeGeometry (src/kernels/voxelizationUtils.cu:47-146; the reference hands this to the un-vendored third-party
 * Voxelizer): solid voxelisation of a closed triangle mesh on the device.  vertices [n_vertices][3] metres,
 * indices [n_triangles][3], triangle_material [n_triangles] unique-material index per triangle (NULL = all 0),
 * dx = voxel edge.  Output: voxelizer-style volumes [vz][vy][vx], `bid` 0 (solid) / 27 (air) / 1..26 (boundary by
 * air-neighbour set) and the material index of the nearest triangle for boundary voxels; voxel (i,j,k) samples the
 * point ((i-1)dx, (j-1)dx, (k-1)dx), dims = ceil(max/dx) + 3.  The _device form returns cudaMalloc'ed volumes on
 * `device` (-1 = current) that pfdtd_setup_mesh_device adopts; the host form copies them out (call
 * pfdtd_voxelize_dims first to size the buffers). */
int pfdtd_voxelize_dims(const float* vertices, uint32_t n_vertices, float dx, uint32_t* vx, uint32_t* vy, uint32_t* vz);
int pfdtd_voxelize_device(int device, const float* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_triangles,
                          const uint8_t* triangle_material, float dx, uint8_t** d_bid, uint8_t** d_mat,
                          uint32_t* vx, uint32_t* vy, uint32_t* vz);
int pfdtd_voxelize(const float* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_triangles,
                   const uint8_t* triangle_material, float dx, uint8_t* h_bid, uint8_t* h_mat);
/* CudaMesh::makePartition (src/kernels/cudaMesh.h:648-751). device_list may be NULL (= 0..n-1). */
int pfdtd_make_partition(pfdtd_solver* s, uint32_t n_partitions, const uint32_t* device_list);

/* ---- mesh queries (CudaMesh getters, src/kernels/cudaMesh.h:214-244) ------ */
int pfdtd_get_dims(pfdtd_solver* s, uint32_t* dim_x, uint32_t* dim_y, uint32_t* dim_z);
int pfdtd_get_counts(pfdtd_solver* s, uint64_t* n_elements, uint64_t* n_air, uint64_t* n_boundary);
int pfdtd_get_num_partitions(pfdtd_solver* s, uint32_t* n);
/* getFirstSliceIdx / getPartitionSize / getDeviceAt */
int pfdtd_get_partition(pfdtd_solver* s, uint32_t k, uint32_t* first_slice, uint32_t* n_slices, uint32_t* device);
/* CudaMesh::getPartitionIndexing (cudaMesh.h:280-307); host-only, needs no device. */
int pfdtd_partition_indexing(uint32_t dim_z, uint32_t n_partitions, uint32_t* first_slice, uint32_t* n_slices);
/* getElementIdxAndDevice (cudaMesh.h:251-266): first partition containing z; (-1,-1) if none. */
int pfdtd_get_element_idx_and_partition(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z,
                                        int* partition, int64_t* element);
/* copies of the node bytes of partition k (n_slices*dimX*dimY each) for bit-exact checks */
int pfdtd_export_partition_nodes(pfdtd_solver* s, uint32_t k, uint8_t* h_pos, uint8_t* h_mat);
/* current pressure field of partition k (host buffer of n_slices*dimX*dimY elements of the dtype);
 * which = 0 current, 1 past */
int pfdtd_export_partition_pressure(pfdtd_solver* s, uint32_t k, int which, void* h_out);

/* raw device pointers of partition k (CudaMesh::getPressurePtrAt / getPastPressurePtrAt /
 * getPositionIdxPtrAt / getMaterialIdxPtrAt, cudaMesh.h:184-212) for capture / visualisation code that runs
 * its own kernels on the fields; any out pointer may be NULL.  Valid until the next pfdtd_make_partition. */
int pfdtd_get_device_pointers(pfdtd_solver* s, uint32_t k, void** d_pressure, void** d_pressure_past,
                              uint8_t** d_position_idx, uint8_t** d_material_idx);

/* ---- captures (the step either side of launchFDTD3dStep in App::executeStep, src/App.cpp:421-431) ----
 * captureSliceFast (src/kernels/visualizationUtils.cu:127-254): pressure (+ optionally the position byte)
 * of one axis-aligned slice of the CURRENT field, gathered on the device(s) from the planes each partition
 * owns, only the slice copied to the host.  orientation 0: xy at z = slice -> [dimY][dimX];
 * 1: xz at y = slice -> [dimZ][dimX]; 2: yz at x = slice -> [dimZ][dimY].  h_pressure holds elements of
 * the solver dtype; h_position may be NULL.  A slice beyond the dimension is PFDTD_ERR_RANGE (the
 * reference logs and skips the capture). */
int pfdtd_capture_slice(pfdtd_solver* s, uint32_t slice, uint32_t orientation, void* h_pressure, uint8_t* h_position);
/* captureMesh (visualizationUtils.cu:111-125): the whole current field, [dimZ][dimY][dimX] of the dtype,
 * assembled from the planes each partition owns (the reference copies partition 0 only). */
int pfdtd_capture_mesh(pfdtd_solver* s, void* h_field);

/* ---- single samples (CudaMesh::setSample/addSample/getSample, cudaMesh.h:321-404,497-584) -- */
int pfdtd_set_sample(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z, double value);
int pfdtd_add_sample(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z, double value);
int pfdtd_get_sample(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z, double* value);
int pfdtd_set_sample_at(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t local_z, uint32_t partition, double value);
int pfdtd_get_sample_at(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t local_z, uint32_t partition, double* value);
/* CudaMesh::switchHalos (cudaMesh.h:432-463,755-763) / flipPressurePointers (:766-779) /
 * resetPressures (:782-791) as stand-alone operations */
int pfdtd_switch_halos(pfdtd_solver* s);
int pfdtd_flip_pressure_pointers(pfdtd_solver* s);
int pfdtd_reset_pressures(pfdtd_solver* s);

/* ---- sources / receivers -------------------------------------------------- */
/* Element coordinates are final voxel indices (SimulationParameters::getSourceElementCoordinates,
 * SimulationParameters.cpp:200-211).  samples = [n][n_steps] of the solver dtype, i.e. the values
 * SimulationParameters::getSourceSample[Double](i, step) returns (:139-159); they are uploaded once
 * and applied on the device (replaces the per-step H2D copies of kernels3d.cu:93-104). */
int pfdtd_set_sources(pfdtd_solver* s, uint32_t n, const int32_t* xyz, const int32_t* src_types,
                      const void* samples, uint32_t n_steps);
int pfdtd_set_receivers(pfdtd_solver* s, uint32_t n, const int32_t* xyz);

/* ---- time stepping --------------------------------------------------------- */
/* launchFDTD3d / launchFDTD3dDouble (src/kernels/kernels3d.cu:31-203, 205-374): run steps
 * 0..n_steps-1 with the per-step order source(step) -> update -> flip -> halo -> receiver(step);
 * h_response[r*n_steps + step] (solver dtype, caller-owned).  interrupt is polled and progress
 * called when step % 100 == 0 (kernels3d.h:36 PROGRESS_MOD); either may be NULL.
 * *seconds_per_step receives the measured wall time per step. */
int pfdtd_run(pfdtd_solver* s, uint32_t n_steps, void* h_response,
              pfdtd_interrupt_cb interrupt, pfdtd_progress_cb progress, float* seconds_per_step);
/* launchFDTD3dStep (src/kernels/kernels3d.cu:376-482): one step; h_response may be NULL, else
 * h_response[r*n_steps_total + step] is written.  direction = +1/-1 (time reversal rule :462-465). */
int pfdtd_step(pfdtd_solver* s, uint32_t step, int direction, void* h_response, uint32_t n_steps_total);

/* Asynchronous pieces used by benchmarks and by the multi-process driver: enqueue `n_steps`
 * steps starting at `first_step` on the solver's streams without synchronising; receiver
 * samples accumulate in device buffers sized by pfdtd_set_sources' n_steps (or
 * pfdtd_reserve_steps).  pfdtd_sync waits for all of the solver's streams;
 * pfdtd_fetch_responses copies [n_rec][n_steps] to the host. */
int pfdtd_reserve_steps(pfdtd_solver* s, uint32_t n_steps);
int pfdtd_enqueue_steps(pfdtd_solver* s, uint32_t first_step, uint32_t n_steps);
int pfdtd_sync(pfdtd_solver* s);
int pfdtd_fetch_responses(pfdtd_solver* s, void* h_response, uint32_t n_steps);
/* device time of the last pfdtd_enqueue_steps call (CUDA events on the compute stream of
 * partition 0 and, separately, of the update kernels only), valid after pfdtd_sync */
int pfdtd_last_timing(pfdtd_solver* s, float* total_ms, float* update_kernel_ms, uint32_t* n_update_launches);
/* the same per-launch event times split by kind: `bulk` = the full-slab or interior launches (the dominant kernel;
 * bulk_planes = z-planes they updated in total), `edge` = the one-plane launches next to a neighbour slab, which run
 * on their own stream concurrently with the interior launch.  Needs PFDTD_OPT_TIME_KERNELS.  (The reference has no
 * counterpart: it times whole steps on the host, kernels3d.cu:61-63,184-202.) */
int pfdtd_last_timing_detail(pfdtd_solver* s, float* bulk_kernel_ms, uint32_t* n_bulk_launches, uint64_t* bulk_planes,
                             float* edge_kernel_ms, uint32_t* n_edge_launches);

/* ---- multi-process z-slab decomposition (one process per GPU) --------------- */
/* The process owns global slices [GLOBAL_Z_FIRST, +vz) (options above) as ONE local partition
 * whose end planes are halos of neighbouring processes.  The halo transport is NCCL
 * point-to-point over NVLink, set up from a 128-byte ncclUniqueId that the caller distributes
 * (torch.distributed is used for exactly that).  rank/nranks order the slabs bottom to top. */
int pfdtd_comm_unique_id(uint8_t* out_id128);
int pfdtd_comm_init(pfdtd_solver* s, const uint8_t* id128, int rank, int nranks);
/* pfdtd_comm_init also maps the neighbour processes' slabs into this process (CUDA IPC, NVLink peer access): the edge
 * launches then store their plane straight into the neighbour's halo plane and hand over through a flag word, so
 * compute and halo transfer are one launch (replaces CudaMesh::switchHalos' copies, cudaMesh.h:432-463).  Interfaces
 * where the mapping is not possible keep ncclSend/ncclRecv.  pfdtd_comm_release unmaps the neighbours' memory; every
 * process must call it (followed by a barrier of the caller's) BEFORE any of them destroys its solver.
 * pfdtd_halo_transport names the transport the stepping loop will use. */
int pfdtd_comm_release(pfdtd_solver* s);
int pfdtd_halo_transport(pfdtd_solver* s, char* buf, size_t buflen);
/* CudaMesh::switchHalos (cudaMesh.h:432-463) alone, `reps` times back to back with nothing else running, timed with
 * CUDA events on the communication stream: the link time of one exchange when both sides are ready.  Collective:
 * every process calls it with the same `reps`. */
int pfdtd_time_halo_exchange(pfdtd_solver* s, uint32_t reps, float* ms_per_exchange);
/* NVLink halo time (ms) accumulated on the communication stream during the last enqueue */
int pfdtd_last_halo_ms(pfdtd_solver* s, float* halo_ms);

/* ---- introspection for tests / benchmarks ----------------------------------- */
/* name of the update kernel variant that pfdtd_enqueue_steps will launch */
int pfdtd_kernel_name(pfdtd_solver* s, char* buf, size_t buflen);
/* number of kernel launches issued by this solver since creation */
int pfdtd_launch_count(pfdtd_solver* s, uint64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* PFDTD_H_ */
