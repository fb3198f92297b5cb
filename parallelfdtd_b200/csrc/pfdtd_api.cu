// pfdtd_api.cu -- the C ABI (include/pfdtd.h): domain container, z-slab partitioning, halo
// exchange and the step loop.  Host-side replacement of the reference's CudaMesh
// (src/kernels/cudaMesh.{h,cu}) and launchFDTD3d* drivers (src/kernels/kernels3d.cu:31-482).
#include "pfdtd_internal.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <list>
#include <map>
#include <mutex>

#define PFDTD_DIF_MAX_ORDER_HOST 4

namespace pfdtd {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

// ---- device block cache (pfdtd_internal.h) ---------------------------------------------------------
namespace {
constexpr size_t kCacheMinBytes = (size_t)1 << 20;
constexpr int kCacheMaxDevices = 64;
struct DevCache {
  struct Block { void* p; size_t bytes; int dev; };
  std::mutex mu;
  std::map<void*, std::pair<size_t, int>> live;   // blocks handed out by dev_alloc: size, device
  std::list<Block> idle;                          // cached blocks, least recently returned first
  size_t idle_bytes[kCacheMaxDevices] = {};
  size_t cap = (size_t)16384 << 20;
  DevCache() {
    if (const char* e = getenv("PFDTD_CACHE_MB")) cap = (size_t)std::max<long long>(0, atoll(e)) << 20;
  }
};
DevCache& dev_cache() {
  static DevCache* c = new DevCache();   // never destroyed: the CUDA context may be gone by the time statics are
  return *c;
}
}  // namespace

void dev_cache_release(int device) {
  DevCache& c = dev_cache();
  std::vector<void*> drop;
  {
    std::lock_guard<std::mutex> g(c.mu);
    for (auto it = c.idle.begin(); it != c.idle.end();) {
      if (device < 0 || it->dev == device) {
        drop.push_back(it->p);
        c.idle_bytes[it->dev] -= it->bytes;
        it = c.idle.erase(it);
      } else {
        ++it;
      }
    }
  }
  for (void* q : drop) cudaFree(q);
}

void dev_cache_forget(void* d_ptr) {
  DevCache& c = dev_cache();
  std::lock_guard<std::mutex> g(c.mu);
  c.live.erase(d_ptr);
}

size_t dev_cache_bytes(int device) {
  DevCache& c = dev_cache();
  std::lock_guard<std::mutex> g(c.mu);
  return (device >= 0 && device < kCacheMaxDevices) ? c.idle_bytes[device] : 0;
}

int dev_alloc(void** d_ptr, size_t bytes) {
  *d_ptr = nullptr;
  DevCache& c = dev_cache();
  size_t want = bytes ? bytes : 1;
  int dev = 0;
  PF_CUDA(cudaGetDevice(&dev));
  const bool cached = c.cap > 0 && want >= kCacheMinBytes && dev < kCacheMaxDevices;
  if (cached) {
    want = (want + 511) & ~(size_t)511;
    void* hit = nullptr;
    {
      std::lock_guard<std::mutex> g(c.mu);
      for (auto it = c.idle.end(); it != c.idle.begin();) {   // most recently returned first
        --it;
        if (it->dev == dev && it->bytes == want) {
          hit = it->p;
          c.idle_bytes[dev] -= want;
          c.idle.erase(it);
          c.live[hit] = {want, dev};
          break;
        }
      }
    }
    if (hit) {
      // what the block held belongs to a solver that is gone: hand it out zero-filled, and finished
      cudaError_t e = cudaMemsetAsync(hit, 0, want, 0);
      if (e == cudaSuccess) e = cudaStreamSynchronize(0);
      if (e != cudaSuccess) { dev_free(hit); PF_CUDA(e); }
      *d_ptr = hit;
      return PFDTD_OK;
    }
  }
  cudaError_t e = cudaMalloc(d_ptr, want);
  if (e == cudaErrorMemoryAllocation) {   // the cache may be what is in the way
    cudaGetLastError();
    dev_cache_release(-1);
    e = cudaMalloc(d_ptr, want);
  }
  if (e != cudaSuccess) *d_ptr = nullptr;
  PF_CUDA(e);
  if (cached) {
    std::lock_guard<std::mutex> g(c.mu);
    c.live[*d_ptr] = {want, dev};
  }
  return PFDTD_OK;
}

void dev_free(void* d_ptr) {
  if (!d_ptr) return;
  DevCache& c = dev_cache();
  size_t bytes = 0;
  int dev = 0;
  {
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.live.find(d_ptr);
    if (it == c.live.end()) {
      for (const auto& b : c.idle) if (b.p == d_ptr) return;   // returned already: the block stays where it is
      bytes = 0;
    } else {
      bytes = it->second.first; dev = it->second.second; c.live.erase(it);
    }
  }
  if (bytes == 0 || bytes > c.cap) { cudaFree(d_ptr); return; }   // not ours (a caller's cudaMalloc), or larger than the cache
  // cudaFree waits for the device; whoever frees a block and allocates the next one relies on that
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != dev) cudaSetDevice(dev);
  const cudaError_t se = cudaDeviceSynchronize();
  if (cur != dev) cudaSetDevice(cur);
  if (se != cudaSuccess) { cudaFree(d_ptr); return; }             // a failed context: nothing to keep
  std::vector<void*> drop;
  {
    std::lock_guard<std::mutex> g(c.mu);
    for (auto it = c.idle.begin(); it != c.idle.end() && c.idle_bytes[dev] + bytes > c.cap;) {
      if (it->dev == dev) {
        drop.push_back(it->p);
        c.idle_bytes[dev] -= it->bytes;
        it = c.idle.erase(it);
      } else {
        ++it;
      }
    }
    c.idle.push_back({d_ptr, bytes, dev});
    c.idle_bytes[dev] += bytes;
  }
  for (void* q : drop) cudaFree(q);
}

// ---- NCCL, loaded lazily so that single-GPU use has no NCCL dependency ---------------------------
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, ...) = nullptr;   // (ncclComm_t*, int nranks, ncclUniqueId by value, int rank)
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId { char bytes[128]; };   // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.lib) return PFDTD_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  PF_CHECK(h != nullptr, PFDTD_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, ...))dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclRecv");
  g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  PF_CHECK(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd,
           PFDTD_ERR_COMM, "libnccl is missing required symbols");
  g_nccl.lib = h;
  return PFDTD_OK;
}
#define PF_NCCL(call)                                                                               \
  do {                                                                                              \
    int r__ = (call);                                                                               \
    if (r__ != 0) {                                                                                 \
      pfdtd::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                            \
                       g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");           \
      return PFDTD_ERR_COMM;                                                                        \
    }                                                                                               \
  } while (0)

// ---- containers -------------------------------------------------------------------------------------
struct Partition {
  int device = 0;
  int64_t first = 0, size = 0;          // slices held: [first, first+size) in the local volume
  uint8_t* pos = nullptr;
  uint8_t* mat = nullptr;
  uint8_t* cls = nullptr;               // node class byte (what the TMA kernel reads instead of pos + mat)
  void* class_table = nullptr;          // ClassEntry<T>[n_classes]
  uint32_t* d_class_keys = nullptr;
  // digital impedance filters
  void* dif_state = nullptr;            // [dif_nb][P]
  uint32_t* dif_rowbase = nullptr;      // [size][Y][ceil(X/128)]
  void* dif_table = nullptr;            // DifEntry<T>[n_lossy]
  // wide meshes: entries per (lossy position class, material), keys = position key | material << 8
  uint32_t* d_wide_keys = nullptr;      // [n_lossy][n_unique]
  void* wide_class_table = nullptr;     // ClassEntry<T>[n_lossy][n_unique]
  void* wide_dif_table = nullptr;       // DifEntry<T>[n_lossy][n_unique]
  uint32_t dif_nb = 0;
  bool owns_nodes = true;
  void* P[2] = {nullptr, nullptr};
  void* materials = nullptr;
  cudaStream_t s_main = nullptr, s_edge = nullptr;
  cudaEvent_t ev_src = nullptr, ev_int = nullptr, ev_edge = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  bool use_tma = false;
  TmaMaps maps[2];                      // maps[c]: current field = P[c], overwritten field = P[1-c]
  TmaMaps maps_tail[2];                 // the same for the tile shape of the launches with a tail (edge planes, fused sources / receivers)
  TmaConfig cfg_tail_full{0, 1, 0};     // full slab on that shape (fused launches)
  TmaConfig cfg_full{0, 1, 0}, cfg_int{0, 1, 0}, cfg_edge{0, 1, 0};
  int* d_step = nullptr;                // [0] step counter, [1] first recordable step, [2] last step of the current enqueue
  // sources / receivers handled by the update launch itself (single slab; tma_common.cuh fused_srcrec)
  bool fused = false;
  FusedParams fused_params{};           // item coordinates, passed to the launch by value
  void* d_fused = nullptr;              // FusedSrcRec<T>
  int* d_item_step = nullptr;           // [n_src + n_rec] per-item step counters
  // sources / receivers that live in this partition
  int n_src = 0, n_rec = 0;
  int64_t* d_src_elem = nullptr; int32_t* d_src_type = nullptr; int32_t* d_src_slot = nullptr;
  int64_t* d_rec_elem = nullptr; int32_t* d_rec_slot = nullptr;
  void* d_src_samples = nullptr;        // [n_src_total][src_steps]
  void* d_rec_out = nullptr;            // [n_rec_total][rec_cap]
  std::vector<int32_t> rec_slots;       // host copy of owned receiver slots
  uint32_t rec_cap_alloc = 0;           // steps d_rec_out has room for
  std::vector<cudaEvent_t> kev;         // kernel timing events (pairs)
  std::vector<int> kev_planes;          // planes updated by the launch each pair brackets
};

}  // namespace pfdtd

using namespace pfdtd;

struct pfdtd_solver {
  // options
  int64_t opt_matidx_as_written = 1, opt_soft_accumulate = 0, opt_kernel = KERNEL_AUTO, opt_global_z_first = 0,
          opt_global_z_dim = 0, opt_double_pad = 0, opt_use_graph = 1, opt_overlap = 1, opt_tma_chunk = 0, opt_tma_tile = 0,
          opt_time_kernels = 0, opt_tma_hints = 0, opt_dif_order = 0, opt_fuse_srcrec = 1;
  int dtype = PFDTD_F32;
  int element_type = 0;
  int scheme = SCH_FORWARD;
  uint32_t X = 0, Y = 0, Z = 0;          // padded dims of the local volume
  uint32_t bx = 32, by = 4, bz = 1;
  double params[4] = {0, 0, 0, 0};
  std::vector<unsigned char> materials_host;   // [n_unique][20] of dtype
  uint32_t n_unique = 0;
  uint64_t n_air = 0, n_boundary = 0;
  bool mesh_ready = false;
  int stage_device = 0;
  uint8_t* d_pos0 = nullptr;             // padded + translated node volumes before partitioning
  uint8_t* d_mat0 = nullptr;
  uint8_t* d_cls0 = nullptr;             // class byte volume before partitioning (null when > 256 classes)
  std::vector<uint32_t> class_keys;      // class id -> pos | mat << 8 | K12 << 16 | K8 << 20
  double dcoef[4] = {0, 0, 0, 0};        // interpolated schemes: d1..d4
  bool dcoef_user = false;
  bool tables_dirty = true;
  bool wide = false;                      // class byte = position class only; lossy voxels look their material up (WideArgs)
  uint32_t dif_lo = 2, n_lossy = 0;       // class ids >= dif_lo are lossy boundary classes
  std::vector<Partition> parts;
  int cur = 0;                           // index of the current field in Partition::P
  int past_direction = 1;                // launchFDTD3dStep's static (kernels3d.cu:386)
  // sources / receivers (host copies)
  std::vector<int32_t> src_xyz, src_type, rec_xyz;
  std::vector<unsigned char> src_samples;   // [n_src][src_steps] of dtype
  uint32_t n_src = 0, n_rec = 0, src_steps = 0, rec_cap = 0;
  bool srcrec_dirty = true;              // receiver (and source) tables must be rebuilt: recorded samples are dropped
  bool src_dirty = false;                // only the source tables changed: the recorded samples stay
  bool fuse_dirty = true;                // the fused source / receiver descriptor must be rebuilt (options changed)
  int dif_order_built = 0;               // filter order the partitions were built for (state layout, row-segment entries)
  // comm
  void* comm = nullptr;
  int rank = 0, nranks = 1;
  cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr;
  float last_total_ms = 0, last_kernel_ms = 0, last_halo_ms = 0, last_edge_ms = 0;
  uint32_t last_kernel_launches = 0, last_edge_launches = 0;
  uint64_t last_bulk_planes = 0;
  uint64_t launch_count = 0;
  std::vector<char> halo_sent_up, halo_sent_down;   // per step: interfaces served by the edge launches' peer stores
  std::vector<char> peer_ok;                        // [np*np]: partition a's device can address partition b's memory
  int64_t opt_peer_stores = 1;
  // neighbour PROCESSES whose slab memory is mapped here (CUDA IPC): their halo planes are written by this process's
  // edge launches, the hand-over goes through flag words (tma_common.cuh).  link[0] = lower neighbour, link[1] = upper.
  struct PeerLink {
    bool ok = false;            // both sides mapped each other and can split their step into edge + interior launches
    void* nb_P[2] = {nullptr, nullptr};
    int* nb_flags = nullptr;
    int64_t nb_size = 0;        // planes in the neighbour's slab
  } link[2];
  int* d_halo_flags = nullptr;  // this process's flag block (own 2 MiB allocation: exported whole)
  // graph
  cudaGraphExec_t graph_exec = nullptr;
  uint32_t graph_steps = 0;
};

namespace pfdtd {

static size_t esize(const pfdtd_solver* s) { return s->dtype == PFDTD_F32 ? 4 : 8; }

// PFDTD_TRACE_SETUP=1: wall-clock of the set-up phases on stderr (device-synchronised at every mark)
struct PhaseTrace {
  bool on = getenv("PFDTD_TRACE_SETUP") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[pfdtd setup] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

static void partition_indexing(uint32_t dim, uint32_t n, std::vector<int64_t>& first, std::vector<int64_t>& size) {
  // CudaMesh::getPartitionIndexing, src/kernels/cudaMesh.h:280-307
  first.resize(n);
  size.resize(n);
  int64_t ps = dim / n;
  for (uint32_t i = 0; i < n; i++) {
    int64_t s_inc = (i == 0) ? 0 : 1, e_inc = (i == n - 1) ? 0 : 1;
    int64_t cs = ps + s_inc + e_inc;
    if (i != 0 && i == n - 1) cs += (int64_t)dim - (int64_t)(i + 1) * ps;
    first[i] = (int64_t)i * ps - s_inc;
    size[i] = cs;
  }
}

// mappings of the neighbour processes' slabs and this process's flag block (pfdtd_comm_init)
static void close_links(pfdtd_solver* s) {
  for (auto& l : s->link) {
    for (void*& q : l.nb_P) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    if (l.nb_flags) { cudaIpcCloseMemHandle(l.nb_flags); l.nb_flags = nullptr; }
    l.ok = false;
  }
  if (s->d_halo_flags) { cudaFree(s->d_halo_flags); s->d_halo_flags = nullptr; }
}

static int free_partitions(pfdtd_solver* s) {
  close_links(s);
  if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
  for (auto& p : s->parts) {
    cudaSetDevice(p.device);
    if (p.s_main) cudaStreamSynchronize(p.s_main);
    if (p.s_edge) cudaStreamSynchronize(p.s_edge);
    if (p.owns_nodes) { dev_free(p.pos); dev_free(p.mat); dev_free(p.cls); }
    cudaFree(p.class_table); cudaFree(p.d_class_keys);
    dev_free(p.dif_state); dev_free(p.dif_rowbase); cudaFree(p.dif_table);
    cudaFree(p.d_wide_keys); cudaFree(p.wide_class_table); cudaFree(p.wide_dif_table);
    dev_free(p.P[0]); dev_free(p.P[1]); cudaFree(p.materials); cudaFree(p.d_step);
    cudaFree(p.d_fused); cudaFree(p.d_item_step);
    cudaFree(p.d_src_elem); cudaFree(p.d_src_type); cudaFree(p.d_src_slot); cudaFree(p.d_rec_elem); cudaFree(p.d_rec_slot);
    cudaFree(p.d_src_samples); cudaFree(p.d_rec_out);
    if (p.s_main) cudaStreamDestroy(p.s_main);
    if (p.s_edge) cudaStreamDestroy(p.s_edge);
    for (cudaEvent_t e : {p.ev_src, p.ev_int, p.ev_edge, p.ev_t0, p.ev_t1}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : p.kev) cudaEventDestroy(e);
  }
  s->parts.clear();
  return PFDTD_OK;
}

static bool has_lower_ext(const pfdtd_solver* s) { return s->comm && s->rank > 0; }
static bool has_upper_ext(const pfdtd_solver* s) { return s->comm && s->rank < s->nranks - 1; }

// ---- per-partition source / receiver tables ----------------------------------------------------------
// Three independent reasons to touch them:
//   srcrec_dirty  receivers (or the slab layout / communicator) changed: everything is rebuilt, recorded samples dropped
//   src_dirty     only the sources changed: source tables rebuilt, the receiver buffer with its samples stays
//   rec_cap grew  (pfdtd_reserve_steps, or enqueue_steps past the reserved length): the receiver buffer is re-allocated
//                 and the samples recorded so far are carried over row by row
static int upload_new(void** d, const void* h, size_t bytes) {
  *d = nullptr;
  if (bytes == 0) return PFDTD_OK;
  PF_CUDA(cudaMalloc(d, bytes));
  PF_CUDA(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
  return PFDTD_OK;
}

// Single slab with the TMA kernel and every source / receiver in an updated plane: the update launch records and injects
// (tma_common.cuh fused_srcrec).  Anything else keeps the separate launch.
static int build_fused(pfdtd_solver* s) {
  for (auto& p : s->parts) {
    p.fused = false;
    p.fused_params = FusedParams{};
    cudaFree(p.d_fused); cudaFree(p.d_item_step);
    p.d_fused = nullptr; p.d_item_step = nullptr;
  }
  if (s->parts.size() != 1 || s->comm || !s->opt_fuse_srcrec) return PFDTD_OK;
  Partition& p = s->parts[0];
  if (!p.use_tma || s->n_src + s->n_rec > PFDTD_FUSED_MAX || s->n_src + s->n_rec == 0) return PFDTD_OK;
  // Worth it where the step is launch-bound: a second launch costs a few microseconds per step, the tail costs every CTA
  // a few hundred cycles -- small slabs only (option value 2 forces it for any size); frequency-independent, narrow classes.
  if (s->opt_dif_order != 0 || s->wide) return PFDTD_OK;
  if (s->opt_fuse_srcrec < 2 && (uint64_t)s->X * s->Y * (uint64_t)p.size > (1ull << 24)) return PFDTD_OK;
  const int64_t zoff = s->opt_global_z_first;
  FusedParams fp{};
  FusedSrcRec<void> h{};
  int k = 0;
  for (int pass = 0; pass < 2; pass++) {
    const uint32_t n = pass == 0 ? s->n_src : s->n_rec;
    const std::vector<int32_t>& xyz = pass == 0 ? s->src_xyz : s->rec_xyz;
    for (uint32_t i = 0; i < n; i++, k++) {
      const int64_t z = (int64_t)xyz[3 * i + 2] - zoff - p.first;
      if (z < 1 || z > p.size - 2) return PFDTD_OK;             // a plane this slab never updates: separate launch
      fp.xyz[k][0] = xyz[3 * i]; fp.xyz[k][1] = xyz[3 * i + 1]; fp.xyz[k][2] = (int)z;
      h.slot[k] = (int)i;
      h.type[k] = pass == 0 ? s->src_type[i] : 0;
    }
  }
  fp.n_src = (int)s->n_src; fp.n_rec = (int)s->n_rec;
  PF_CUDA(cudaSetDevice(p.device));
  PF_CUDA(cudaMalloc(&p.d_item_step, PFDTD_FUSED_MAX * sizeof(int)));
  PF_CUDA(cudaMemset(p.d_item_step, 0, PFDTD_FUSED_MAX * sizeof(int)));
  h.soft_accumulate = (int)s->opt_soft_accumulate;
  h.rec_stride = p.rec_cap_alloc; h.src_stride = s->src_steps;
  h.rec_out = p.d_rec_out; h.src_samples = p.d_src_samples;
  h.d_step = p.d_step; h.item_step = p.d_item_step;
  PF_CUDA(cudaMalloc(&p.d_fused, sizeof(h)));
  PF_CUDA(cudaMemcpy(p.d_fused, &h, sizeof(h), cudaMemcpyHostToDevice));
  p.fused_params = fp;
  p.fused = true;
  return PFDTD_OK;
}

static int prepare_srcrec(pfdtd_solver* s) {
  PF_CHECK(!s->parts.empty(), PFDTD_ERR_INVALID, "make_partition must be called first");
  const int64_t XY = (int64_t)s->X * s->Y;
  const int64_t zoff = s->opt_global_z_first;
  if (s->rec_cap < s->src_steps) s->rec_cap = s->src_steps;
  if (s->rec_cap == 0) s->rec_cap = 1;
  bool changed = false;
  for (size_t k = 0; k < s->parts.size(); k++) {
    Partition& p = s->parts[k];
    const bool all = s->srcrec_dirty;
    const bool grow = !all && s->n_rec && p.rec_cap_alloc < s->rec_cap;
    if (!all && !s->src_dirty && !grow) continue;
    changed = true;
    PF_CUDA(cudaSetDevice(p.device));
    if (p.s_main) PF_CUDA(cudaStreamSynchronize(p.s_main));
    if (p.s_edge) PF_CUDA(cudaStreamSynchronize(p.s_edge));
    if (all || s->src_dirty) {
      for (void* q : {(void*)p.d_src_elem, (void*)p.d_src_type, (void*)p.d_src_slot, p.d_src_samples}) cudaFree(q);
      p.d_src_elem = nullptr; p.d_src_type = nullptr; p.d_src_slot = nullptr; p.d_src_samples = nullptr;
      std::vector<int64_t> se;
      std::vector<int32_t> st, ss;
      // sources: every partition that holds the slice, halo copies included (cudaMesh.h:321-338)
      for (uint32_t i = 0; i < s->n_src; i++) {
        int64_t z = (int64_t)s->src_xyz[3 * i + 2] - zoff;
        if (z < p.first || z > p.first + p.size - 1) continue;
        se.push_back((z - p.first) * XY + (int64_t)s->src_xyz[3 * i + 1] * s->X + s->src_xyz[3 * i]);
        st.push_back(s->src_type[i]);
        ss.push_back((int32_t)i);
      }
      p.n_src = (int)se.size();
      PF_TRY(upload_new((void**)&p.d_src_elem, se.data(), se.size() * 8));
      PF_TRY(upload_new((void**)&p.d_src_type, st.data(), st.size() * 4));
      PF_TRY(upload_new((void**)&p.d_src_slot, ss.data(), ss.size() * 4));
      if (p.n_src) PF_TRY(upload_new(&p.d_src_samples, s->src_samples.data(), s->src_samples.size()));
    }
    if (all) {
      for (void* q : {(void*)p.d_rec_elem, (void*)p.d_rec_slot, p.d_rec_out}) cudaFree(q);
      p.d_rec_elem = nullptr; p.d_rec_slot = nullptr; p.d_rec_out = nullptr;
      std::vector<int64_t> re;
      std::vector<int32_t> rs;
      // receivers: first partition containing the slice (cudaMesh.h:251-266); across processes the
      // lower process owns the two shared slices
      for (uint32_t i = 0; i < s->n_rec; i++) {
        int64_t z = (int64_t)s->rec_xyz[3 * i + 2] - zoff;
        if (z < p.first || z > p.first + p.size - 1) continue;
        if (k > 0 && z <= s->parts[k - 1].first + s->parts[k - 1].size - 1) continue;
        if (k == 0 && zoff > 0 && z <= 1) continue;
        re.push_back((z - p.first) * XY + (int64_t)s->rec_xyz[3 * i + 1] * s->X + s->rec_xyz[3 * i]);
        rs.push_back((int32_t)i);
      }
      p.n_rec = (int)re.size();
      p.rec_slots = rs;
      PF_TRY(upload_new((void**)&p.d_rec_elem, re.data(), re.size() * 8));
      PF_TRY(upload_new((void**)&p.d_rec_slot, rs.data(), rs.size() * 4));
      p.rec_cap_alloc = 0;
      if (s->n_rec) {
        const size_t bytes = (size_t)s->n_rec * s->rec_cap * esize(s);
        PF_CUDA(cudaMalloc(&p.d_rec_out, bytes));
        PF_CUDA(cudaMemset(p.d_rec_out, 0, bytes));
        p.rec_cap_alloc = s->rec_cap;
      }
    } else if (grow) {
      void* bigger = nullptr;
      const size_t es = esize(s);
      const size_t bytes = (size_t)s->n_rec * s->rec_cap * es;
      PF_CUDA(cudaMalloc(&bigger, bytes));
      PF_CUDA(cudaMemset(bigger, 0, bytes));
      if (p.d_rec_out && p.rec_cap_alloc)
        PF_CUDA(cudaMemcpy2D(bigger, (size_t)s->rec_cap * es, p.d_rec_out, (size_t)p.rec_cap_alloc * es, (size_t)p.rec_cap_alloc * es,
                             s->n_rec, cudaMemcpyDeviceToDevice));
      cudaFree(p.d_rec_out);
      p.d_rec_out = bigger;
      p.rec_cap_alloc = s->rec_cap;
    }
  }
  if (changed || s->fuse_dirty) PF_TRY(build_fused(s));
  s->fuse_dirty = false;
  s->srcrec_dirty = false;
  s->src_dirty = false;
  if (changed && s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }   // the graph holds the old pointers
  return PFDTD_OK;
}

static UpdateArgs make_update_args(pfdtd_solver* s, Partition& p, int z_begin, int z_end, cudaStream_t st) {
  UpdateArgs a{};
  a.dtype = s->dtype;
  a.scheme = s->scheme;
  a.pos = p.pos;
  a.mat = p.mat;
  a.cls = p.cls;
  a.class_table = p.class_table;
  a.wide = s->wide;
  a.wide_class_table = p.wide_class_table;
  a.wide_dif_table = p.wide_dif_table;
  a.n_classes = (int)s->class_keys.size();
  a.tma_hints = (int)s->opt_tma_hints;
  a.peer_plane = nullptr;
  a.tail = 0;
  a.fused_srcrec = nullptr;
  a.fused_params = FusedParams{};
  a.dif_order = p.dif_rowbase ? (int)s->opt_dif_order : 0;
  a.dif_state = p.dif_state;
  a.dif_rowbase = p.dif_rowbase;
  a.dif_table = p.dif_table;
  a.dif_nb = p.dif_nb;
  a.dif_lo = s->dif_lo;
  a.n_dif = (int)s->n_lossy;
  a.P = p.P[s->cur];
  a.Pn = p.P[1 - s->cur];
  a.materials = p.materials;
  a.n_coefs = s->n_unique * 20;
  for (int i = 0; i < 4; i++) a.params[i] = s->params[i];
  for (int i = 0; i < 4; i++) a.dcoef[i] = s->dcoef[i];
  a.matidx_as_written = (int)s->opt_matidx_as_written;
  a.X = (int)s->X;
  a.Y = (int)s->Y;
  a.z_begin = z_begin;
  a.z_end = z_end;
  a.stream = st;
  return a;
}

// (Re)build the per-class coefficient tables: they depend on params, materials and the material-index option.
static int ensure_class_tables(pfdtd_solver* s) {
  if (!s->tables_dirty) return PFDTD_OK;
  for (auto& p : s->parts) {
    if (!p.class_table) continue;
    PF_CUDA(cudaSetDevice(p.device));
    UpdateArgs a = make_update_args(s, p, 0, 0, p.s_main);
    a.dif_order = (int)s->opt_dif_order;     // class entries take b0 as the scalar admittance when filters are on
    PF_TRY(build_class_table(a, p.d_class_keys, (int)s->class_keys.size(), p.class_table));
    if (p.dif_table) PF_TRY(build_dif_table(a, p.d_class_keys + s->dif_lo, (int)s->n_lossy, p.dif_table));
    if (p.wide_class_table) {   // the same builders over the (class, material) key list
      const int nw = (int)(s->n_lossy * s->n_unique);
      PF_TRY(build_class_table(a, p.d_wide_keys, nw, p.wide_class_table));
      if (p.wide_dif_table) PF_TRY(build_dif_table(a, p.d_wide_keys, nw, p.wide_dif_table));
    }
    PF_CUDA(cudaStreamSynchronize(p.s_main));
    s->launch_count++;
  }
  s->tables_dirty = false;
  return PFDTD_OK;
}

static int launch_update(pfdtd_solver* s, Partition& p, int z_begin, int z_end, const TmaConfig& cfg, cudaStream_t st,
                         bool timed, void* peer_plane = nullptr, bool fused = false) {
  if (z_end <= z_begin) return PFDTD_OK;
  UpdateArgs a = make_update_args(s, p, z_begin, z_end, st);
  if (fused) { a.fused_srcrec = p.d_fused; a.fused_params = p.fused_params; a.tail = 2; }
  else if (peer_plane != nullptr && z_end - z_begin == 1 && p.use_tma) a.tail = 1;
  const TmaMaps& maps = a.tail ? p.maps_tail[s->cur] : p.maps[s->cur];
  TmaConfig cfg_used = cfg;
  if (a.tail == 1) cfg_used.tile = tma_tail_tile(s->dtype, s->scheme);
  a.peer_plane = peer_plane;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed) {
    PF_CUDA(cudaEventCreate(&e0));
    PF_CUDA(cudaEventCreate(&e1));
    p.kev.push_back(e0);
    p.kev.push_back(e1);
    p.kev_planes.push_back(z_end - z_begin);
    PF_CUDA(cudaEventRecord(e0, st));
  }
  if (s->scheme == SCH_INTERP) {
    if (p.use_tma) PF_TRY(launch_update_interp_tma(a, maps, cfg_used, nullptr));
    else PF_TRY(launch_update_interp_plain(a));
  } else if (p.use_tma) PF_TRY(launch_update_tma(a, maps, cfg_used));
  else PF_TRY(launch_update_plain(a));
  if (timed) PF_CUDA(cudaEventRecord(e1, st));
  s->launch_count++;
  return PFDTD_OK;
}

// peer-mapped halo hand-over with the lower / upper neighbour PROCESS is in use for the stepping loop
static bool ipc_side(const pfdtd_solver* s, int side) {
  if (!s->comm || s->parts.size() != 1 || !s->opt_peer_stores || !s->opt_overlap || !s->link[side].ok) return false;
  return side == 0 ? s->rank > 0 : s->rank < s->nranks - 1;
}

static int launch_srcrec_for(pfdtd_solver* s, Partition& p, cudaStream_t st, int record, int inject, int advance) {
  const bool wait_lo = ipc_side(s, 0), wait_hi = ipc_side(s, 1);
  if (p.n_src == 0 && p.n_rec == 0 && !advance && !wait_lo && !wait_hi) return PFDTD_OK;
  SrcRecArgs a{};
  a.halo_flags = (wait_lo || wait_hi) ? s->d_halo_flags : nullptr;
  a.wait_lo = wait_lo; a.wait_hi = wait_hi;
  a.dtype = s->dtype;
  a.P = p.P[s->cur];
  a.n_rec = p.n_rec; a.rec_elem = p.d_rec_elem; a.rec_slot = p.d_rec_slot; a.rec_out = p.d_rec_out; a.rec_stride = p.rec_cap_alloc;
  a.n_src = p.n_src; a.src_elem = p.d_src_elem; a.src_type = p.d_src_type; a.src_slot = p.d_src_slot;
  a.src_samples = p.d_src_samples; a.src_stride = s->src_steps;
  a.d_step = p.d_step;
  a.do_record = record; a.do_inject = inject; a.soft_accumulate = (int)s->opt_soft_accumulate; a.advance = advance;
  a.stream = st;
  PF_TRY(launch_srcrec(a));
  s->launch_count++;
  return PFDTD_OK;
}

// Halo exchange of field index `b` for every interface this process touches, enqueued on the edge
// streams (CudaMesh::switchHalos, cudaMesh.h:432-463: slab_i[last-1] -> slab_{i+1}[0],
// slab_{i+1}[1] -> slab_i[last]).  Local interfaces: peer copies.  Process boundaries: NCCL send/recv.
static int enqueue_halo_local_up(pfdtd_solver* s, size_t k, int b) {   // copies issued by partition k towards k+1 ... and k+1 -> k
  Partition& lo = s->parts[k];
  Partition& hi = s->parts[k + 1];
  const size_t plane = (size_t)s->X * s->Y * esize(s);
  char* src1 = (char*)lo.P[b] + (size_t)(lo.size - 2) * plane;
  char* dst1 = (char*)hi.P[b];
  PF_CUDA(cudaSetDevice(lo.device));
  PF_CUDA(cudaMemcpyPeerAsync(dst1, hi.device, src1, lo.device, plane, lo.s_edge));
  return PFDTD_OK;
}
static int enqueue_halo_local_down(pfdtd_solver* s, size_t k, int b) {  // partition k+1's plane 1 -> partition k's last plane
  Partition& lo = s->parts[k];
  Partition& hi = s->parts[k + 1];
  const size_t plane = (size_t)s->X * s->Y * esize(s);
  char* src2 = (char*)hi.P[b] + plane;
  char* dst2 = (char*)lo.P[b] + (size_t)(lo.size - 1) * plane;
  PF_CUDA(cudaSetDevice(hi.device));
  PF_CUDA(cudaMemcpyPeerAsync(dst2, lo.device, src2, hi.device, plane, hi.s_edge));
  return PFDTD_OK;
}
static int enqueue_halo_external(pfdtd_solver* s, int b, bool want_lo = true, bool want_hi = true) {
  if (!s->comm) return PFDTD_OK;
  const size_t plane = (size_t)s->X * s->Y * esize(s);
  const bool lo_ext = has_lower_ext(s) && want_lo, hi_ext = has_upper_ext(s) && want_hi;
  if (!lo_ext && !hi_ext) return PFDTD_OK;
  Partition& p = s->parts.front();        // pfdtd_comm_init: exactly one local partition
  PF_CUDA(cudaSetDevice(p.device));
  PF_NCCL(g_nccl.GroupStart());
  if (lo_ext) {
    PF_NCCL(g_nccl.Send((char*)p.P[b] + plane, plane, 1 /*ncclUint8*/, s->rank - 1, s->comm, p.s_edge));
    PF_NCCL(g_nccl.Recv((char*)p.P[b], plane, 1, s->rank - 1, s->comm, p.s_edge));
  }
  if (hi_ext) {
    PF_NCCL(g_nccl.Send((char*)p.P[b] + (size_t)(p.size - 2) * plane, plane, 1, s->rank + 1, s->comm, p.s_edge));
    PF_NCCL(g_nccl.Recv((char*)p.P[b] + (size_t)(p.size - 1) * plane, plane, 1, s->rank + 1, s->comm, p.s_edge));
  }
  PF_NCCL(g_nccl.GroupEnd());
  return PFDTD_OK;
}

static bool can_store_to(const pfdtd_solver* s, size_t a, size_t b) {
  const size_t np = s->parts.size();
  return s->opt_peer_stores && s->peer_ok.size() == np * np && s->peer_ok[a * np + b];
}

// One time step for all partitions: source(n) -> update -> flip -> halo, receiver(n) is recorded by the
// next step's srcrec launch (or the final flush).  Edge planes are computed first on the edge stream,
// their halo transfer overlaps the interior update on the main stream.
static int enqueue_one_step(pfdtd_solver* s, bool timed, bool time_halo) {
  const size_t np = s->parts.size();
  const int c = s->cur;
  const bool single = (np == 1 && !s->comm);
  if (single) {
    Partition& p = s->parts[0];
    PF_CUDA(cudaSetDevice(p.device));
    // fused: the update launch records this step's receivers, injects the next step's sources and advances the step
    if (!p.fused) PF_TRY(launch_srcrec_for(s, p, p.s_main, 1, 1, 1));
    PF_TRY(launch_update(s, p, 1, (int)p.size - 1, p.fused ? p.cfg_tail_full : p.cfg_full, p.s_main, timed, nullptr, p.fused));
    s->cur = 1 - c;
    return PFDTD_OK;
  }
  const size_t plane_bytes = (size_t)s->X * s->Y * esize(s);
  s->halo_sent_up.assign(np, 0);      // interface k|k+1: partition k's top plane already stored into k+1's plane 0
  s->halo_sent_down.assign(np, 0);    //                  partition k+1's plane 1 already stored into k's last plane
  bool ext_sent_lo = false, ext_sent_hi = false;   // process boundaries served by the edge launches (CUDA IPC mapping)
  // phase 1: sources / receivers on the MAIN stream, right behind the previous interior launch: the critical path of a
  // step is interior -> sources/receivers -> interior on one stream, the edge planes and the halo traffic hang off it.
  for (size_t k = 0; k < np; k++) {
    Partition& p = s->parts[k];
    PF_CUDA(cudaSetDevice(p.device));
    // everything this step reads has arrived: own edge planes and halo traffic of the previous step, and the
    // neighbours' edge work (they wrote this slab's halo planes; across processes the flag wait in srcrec does this)
    PF_CUDA(cudaStreamWaitEvent(p.s_main, p.ev_edge, 0));
    if (k > 0) PF_CUDA(cudaStreamWaitEvent(p.s_main, s->parts[k - 1].ev_edge, 0));
    if (k + 1 < np) PF_CUDA(cudaStreamWaitEvent(p.s_main, s->parts[k + 1].ev_edge, 0));
    PF_TRY(launch_srcrec_for(s, p, p.s_main, 1, 1, 1));
    PF_CUDA(cudaEventRecord(p.ev_src, p.s_main));
  }
  for (size_t k = 0; k < np; k++) {
    Partition& p = s->parts[k];
    PF_CUDA(cudaSetDevice(p.device));
    const bool lo_nb = (k > 0) || has_lower_ext(s);
    const bool hi_nb = (k + 1 < np) || has_upper_ext(s);
    const int zb = 1, ze = (int)p.size - 1;   // updated planes [zb, ze)
    int ib = zb, ie = ze;
    const bool split = s->opt_overlap && (ze - zb) >= 4;
    if (split) {
      // local neighbours whose memory this device can address get their halo plane from the edge launch itself
      // (peer-mapped stores); otherwise the plane is copied / sent below
      void* peer_lo = nullptr;
      void* peer_hi = nullptr;
      if (p.use_tma && k > 0 && can_store_to(s, k, k - 1)) {
        Partition& nb = s->parts[k - 1];
        peer_lo = (char*)nb.P[1 - c] + (size_t)(nb.size - 1) * plane_bytes;
        s->halo_sent_down[k - 1] = 1;
      }
      if (p.use_tma && k + 1 < np && can_store_to(s, k, k + 1)) {
        peer_hi = s->parts[k + 1].P[1 - c];
        s->halo_sent_up[k] = 1;
      }
      // neighbour PROCESSES whose slab is mapped here: same stores, then the flag hand-over instead of a stream event
      int *sig_lo = nullptr, *sig_hi = nullptr;
      if (p.use_tma && k == 0 && ipc_side(s, 0)) {
        peer_lo = (char*)s->link[0].nb_P[1 - c] + (size_t)(s->link[0].nb_size - 1) * plane_bytes;
        sig_lo = s->link[0].nb_flags + 1 /* HALO_FROM_HI: this process is its upper neighbour */;
        ext_sent_lo = true;
      }
      if (p.use_tma && k + 1 == np && ipc_side(s, 1)) {
        peer_hi = s->link[1].nb_P[1 - c];
        sig_hi = s->link[1].nb_flags + 0 /* HALO_FROM_LO */;
        ext_sent_hi = true;
      }
      // edge planes on the edge stream (it waits for the sources, and through them for the previous interior launch)
      PF_CUDA(cudaStreamWaitEvent(p.s_edge, p.ev_src, 0));
      if (lo_nb) { PF_TRY(launch_update(s, p, zb, zb + 1, p.cfg_edge, p.s_edge, timed, peer_lo)); ib = zb + 1; }
      if (hi_nb) { PF_TRY(launch_update(s, p, ze - 1, ze, p.cfg_edge, p.s_edge, timed, peer_hi)); ie = ze - 1; }
      if (sig_lo || sig_hi) { PF_TRY(launch_halo_publish(s->d_halo_flags, sig_lo, sig_hi, p.s_edge)); s->launch_count++; }
      // interior on the main stream, concurrently with the edge launches and the halo traffic below
      PF_TRY(launch_update(s, p, ib, ie, p.cfg_int, p.s_main, timed));
      PF_CUDA(cudaEventRecord(p.ev_int, p.s_main));
    } else {
      PF_TRY(launch_update(s, p, zb, ze, p.cfg_full, p.s_main, timed));
      PF_CUDA(cudaEventRecord(p.ev_int, p.s_main));
      PF_CUDA(cudaStreamWaitEvent(p.s_edge, p.ev_int, 0));   // the halo traffic below reads what this launch wrote
    }
  }
  // phase 2: halo traffic of the new field (index 1-c) on the edge streams
  if (time_halo && s->ev_h0) { PF_CUDA(cudaSetDevice(s->parts[0].device)); PF_CUDA(cudaEventRecord(s->ev_h0, s->parts[0].s_edge)); }
  for (size_t k = 0; k + 1 < np; k++) {
    if (!s->halo_sent_up[k]) PF_TRY(enqueue_halo_local_up(s, k, 1 - c));
    if (!s->halo_sent_down[k]) PF_TRY(enqueue_halo_local_down(s, k, 1 - c));
  }
  PF_TRY(enqueue_halo_external(s, 1 - c, !ext_sent_lo, !ext_sent_hi));
  if (time_halo && s->ev_h1) { PF_CUDA(cudaSetDevice(s->parts[0].device)); PF_CUDA(cudaEventRecord(s->ev_h1, s->parts[0].s_edge)); }
  for (size_t k = 0; k < np; k++) {
    Partition& p = s->parts[k];
    PF_CUDA(cudaSetDevice(p.device));
    PF_CUDA(cudaEventRecord(p.ev_edge, p.s_edge));
  }
  s->cur = 1 - c;
  return PFDTD_OK;
}

static int set_step_counters(pfdtd_solver* s, int step, int first_recordable, int last_step) {
  int h[3] = {step, first_recordable, last_step};
  for (auto& p : s->parts) {
    PF_CUDA(cudaSetDevice(p.device));
    cudaStream_t st = p.s_main;   // the stream the source / receiver launches run on
    PF_CUDA(cudaMemcpyAsync(p.d_step, h, sizeof(h), cudaMemcpyHostToDevice, st));
    if (p.d_item_step) {
      int items[PFDTD_FUSED_MAX];
      for (int& v : items) v = step;
      PF_CUDA(cudaMemcpyAsync(p.d_item_step, items, sizeof(items), cudaMemcpyHostToDevice, st));
    }
  }
  return PFDTD_OK;
}

static int sync_all(pfdtd_solver* s) {
  for (auto& p : s->parts) {
    PF_CUDA(cudaSetDevice(p.device));
    PF_CUDA(cudaStreamSynchronize(p.s_edge));
    PF_CUDA(cudaStreamSynchronize(p.s_main));
  }
  return PFDTD_OK;
}

static int elem_of(pfdtd_solver* s, uint32_t x, uint32_t y, int64_t zl, size_t k, int64_t* e) {
  PF_CHECK(x < s->X && y < s->Y, PFDTD_ERR_RANGE, "coordinate (%u,%u) outside mesh %ux%u", x, y, s->X, s->Y);
  *e = (zl - s->parts[k].first) * (int64_t)s->X * s->Y + (int64_t)y * s->X + x;
  return PFDTD_OK;
}

// valueToDevice of the reference fills on the host and copies (cudaUtils.h:84-99); here the value is written on the device
template <typename W>
__global__ void fill_words(W* __restrict__ d, W v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = v;
}

// ---- CUDA-IPC links to the neighbour processes ------------------------------------------------------------------
// Every process exports its two field buffers and its flag block, the neighbours map them (NVLink peer access), and
// both sides confirm.  The handles travel over the NCCL communicator that exists anyway.  An interface uses the
// mapping only when BOTH sides mapped each other and both can split their step into edge + interior launches with
// the TMA kernels; otherwise that interface keeps ncclSend/ncclRecv.
struct LinkHello {
  cudaIpcMemHandle_t p0, p1, flags;
  int64_t size;
  int32_t can_peer, pad;
};

static int exchange_links(pfdtd_solver* s) {
  close_links(s);
  if (s->nranks < 2 || getenv("PFDTD_NO_IPC")) return PFDTD_OK;
  Partition& p = s->parts[0];
  PF_CUDA(cudaSetDevice(p.device));
  PF_CUDA(cudaMalloc(&s->d_halo_flags, 2u << 20));
  PF_CUDA(cudaMemset(s->d_halo_flags, 0, 2u << 20));
  LinkHello mine{};
  bool exportable = cudaIpcGetMemHandle(&mine.p0, p.P[0]) == cudaSuccess && cudaIpcGetMemHandle(&mine.p1, p.P[1]) == cudaSuccess &&
                    cudaIpcGetMemHandle(&mine.flags, s->d_halo_flags) == cudaSuccess;
  if (!exportable) cudaGetLastError();
  // a field another process maps stays out of the block cache: it goes back to the driver with its solver, as it
  // always did, instead of turning up inside a later solver while a neighbour may still hold the old mapping
  if (exportable) { dev_cache_forget(p.P[0]); dev_cache_forget(p.P[1]); }
  mine.size = p.size;
  mine.can_peer = (exportable && p.use_tma && (int)p.size - 2 >= 4) ? 1 : 0;
  LinkHello* d_hello = nullptr;     // [0] mine, [1] from the lower neighbour, [2] from the upper one
  int* d_ack = nullptr;             // [0] mine towards lower, [1] mine towards upper, [2] from lower, [3] from upper
  PF_CUDA(cudaMalloc(&d_hello, 3 * sizeof(LinkHello)));
  PF_CUDA(cudaMalloc(&d_ack, 4 * sizeof(int)));
  PF_CUDA(cudaMemset(d_hello, 0, 3 * sizeof(LinkHello)));
  PF_CUDA(cudaMemset(d_ack, 0, 4 * sizeof(int)));
  PF_CUDA(cudaMemcpy(d_hello, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  const bool lo = has_lower_ext(s), hi = has_upper_ext(s);
  auto cleanup = [&]() { cudaFree(d_hello); cudaFree(d_ack); };
  auto round = [&](const void* send_lo, const void* send_hi, void* recv_lo, void* recv_hi, size_t bytes) -> int {
    PF_NCCL(g_nccl.GroupStart());
    if (lo) { PF_NCCL(g_nccl.Send(send_lo, bytes, 1, s->rank - 1, s->comm, p.s_edge)); PF_NCCL(g_nccl.Recv(recv_lo, bytes, 1, s->rank - 1, s->comm, p.s_edge)); }
    if (hi) { PF_NCCL(g_nccl.Send(send_hi, bytes, 1, s->rank + 1, s->comm, p.s_edge)); PF_NCCL(g_nccl.Recv(recv_hi, bytes, 1, s->rank + 1, s->comm, p.s_edge)); }
    PF_NCCL(g_nccl.GroupEnd());
    PF_CUDA(cudaStreamSynchronize(p.s_edge));
    return PFDTD_OK;
  };
  int rc = round(d_hello, d_hello, d_hello + 1, d_hello + 2, sizeof(LinkHello));
  if (rc != PFDTD_OK) { cleanup(); return rc; }
  LinkHello got[3];
  PF_CUDA(cudaMemcpy(got, d_hello, sizeof(got), cudaMemcpyDeviceToHost));
  int ack[4] = {0, 0, 0, 0};
  for (int side = 0; side < 2; side++) {
    if (!(side == 0 ? lo : hi)) continue;
    const LinkHello& nb = got[1 + side];
    auto& l = s->link[side];
    bool mapped = mine.can_peer && nb.can_peer;
    if (mapped) {
      mapped = cudaIpcOpenMemHandle(&l.nb_P[0], nb.p0, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
               cudaIpcOpenMemHandle(&l.nb_P[1], nb.p1, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
               cudaIpcOpenMemHandle((void**)&l.nb_flags, nb.flags, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      if (!mapped) {
        cudaGetLastError();
        for (void*& q : l.nb_P) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
        if (l.nb_flags) { cudaIpcCloseMemHandle(l.nb_flags); l.nb_flags = nullptr; }
      }
    }
    l.nb_size = nb.size;
    ack[side] = mapped ? 1 : 0;
  }
  PF_CUDA(cudaMemcpy(d_ack, ack, 2 * sizeof(int), cudaMemcpyHostToDevice));
  rc = round(d_ack + 0, d_ack + 1, d_ack + 2, d_ack + 3, sizeof(int));
  if (rc != PFDTD_OK) { cleanup(); return rc; }
  PF_CUDA(cudaMemcpy(ack, d_ack, sizeof(ack), cudaMemcpyDeviceToHost));
  for (int side = 0; side < 2; side++) s->link[side].ok = (side == 0 ? lo : hi) && ack[side] && ack[2 + side];
  cleanup();
  return PFDTD_OK;
}

static int select_device(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    set_error("no CUDA device: libpfdtd_b200 has no CPU fallback");
    return PFDTD_ERR_NO_DEVICE;
  }
  if (device >= 0) {
    PF_CHECK(device < ndev, PFDTD_ERR_RANGE, "device %d of %d", device, ndev);
    PF_CUDA(cudaSetDevice(device));
  }
  return PFDTD_OK;
}

}  // namespace pfdtd

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" {

const char* pfdtd_last_error(void) { return g_last_error.c_str(); }
const char* pfdtd_version(void) { return "pfdtd-b200 0.1 (sm_100a)"; }

int pfdtd_device_count(int* out_count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    *out_count = 0;
    return PFDTD_ERR_NO_DEVICE;
  }
  *out_count = n;
  return PFDTD_OK;
}

int pfdtd_device_mem_mb(int device, int* out_total_mb, int* out_free_mb) {
  PF_CUDA(cudaSetDevice(device));
  size_t f = 0, t = 0;
  PF_CUDA(cudaMemGetInfo(&f, &t));
  f += dev_cache_bytes(device);   // blocks the library holds for reuse are given up when an allocation needs them
  if (out_total_mb) *out_total_mb = (int)(t >> 20);
  if (out_free_mb) *out_free_mb = (int)(f >> 20);
  return PFDTD_OK;
}

int pfdtd_current_device(int* device) {
  PF_CHECK(device != nullptr, PFDTD_ERR_INVALID, "null out pointer");
  PF_TRY(select_device(-1));
  PF_CUDA(cudaGetDevice(device));
  return PFDTD_OK;
}

int pfdtd_device_alloc(int device, size_t bytes, void** d_ptr) {
  PF_CHECK(d_ptr != nullptr, PFDTD_ERR_INVALID, "null out pointer");
  *d_ptr = nullptr;
  PF_TRY(select_device(device));
  return dev_alloc(d_ptr, bytes);
}

int pfdtd_device_fill(int device, void* d_ptr, size_t count, size_t elem_size, const void* value) {
  PF_CHECK(d_ptr != nullptr && value != nullptr, PFDTD_ERR_INVALID, "null pointer");
  PF_CHECK(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8, PFDTD_ERR_INVALID, "element size %zu", elem_size);
  PF_TRY(select_device(device));
  if (count == 0) return PFDTD_OK;
  const unsigned int blocks = (unsigned int)std::min<size_t>((count + 255) / 256, 148 * 8);
  if (elem_size == 1) {
    PF_CUDA(cudaMemset(d_ptr, *(const uint8_t*)value, count));
  } else if (elem_size == 2) {
    uint16_t v; memcpy(&v, value, 2);
    fill_words<uint16_t><<<blocks, 256>>>((uint16_t*)d_ptr, v, count);
  } else if (elem_size == 4) {
    uint32_t v; memcpy(&v, value, 4);
    fill_words<uint32_t><<<blocks, 256>>>((uint32_t*)d_ptr, v, count);
  } else {
    uint64_t v; memcpy(&v, value, 8);
    fill_words<uint64_t><<<blocks, 256>>>((uint64_t*)d_ptr, v, count);
  }
  PF_CUDA(cudaGetLastError());
  PF_CUDA(cudaDeviceSynchronize());
  return PFDTD_OK;
}

int pfdtd_device_upload(int device, void* d_dst, const void* h_src, size_t bytes) {
  PF_CHECK(bytes == 0 || (d_dst != nullptr && h_src != nullptr), PFDTD_ERR_INVALID, "null pointer");
  PF_TRY(select_device(device));
  if (bytes) PF_CUDA(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
  return PFDTD_OK;
}

int pfdtd_device_download(int device, void* h_dst, const void* d_src, size_t bytes) {
  PF_CHECK(bytes == 0 || (h_dst != nullptr && d_src != nullptr), PFDTD_ERR_INVALID, "null pointer");
  PF_TRY(select_device(device));
  if (bytes) PF_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
  return PFDTD_OK;
}

int pfdtd_device_free(int device, void* d_ptr) {
  if (!d_ptr) return PFDTD_OK;
  PF_TRY(select_device(device));
  dev_free(d_ptr);
  return PFDTD_OK;
}

int pfdtd_release_cached_memory(int device) {
  PF_CHECK(device >= -1, PFDTD_ERR_INVALID, "device %d", device);
  dev_cache_release(device);   // no CUDA call unless something is cached
  return PFDTD_OK;
}

int pfdtd_create(pfdtd_solver** out) {
  PF_CHECK(out != nullptr, PFDTD_ERR_INVALID, "null out pointer");
  *out = new pfdtd_solver();
  // tuning knobs for sweeps: defaults of PFDTD_OPT_TMA_TILE / PFDTD_OPT_TMA_CHUNK (pfdtd_set_option still overrides)
  if (const char* e = getenv("PFDTD_TMA_TILE")) (*out)->opt_tma_tile = atoll(e);
  if (const char* e = getenv("PFDTD_TMA_CHUNK")) (*out)->opt_tma_chunk = atoll(e);
  return PFDTD_OK;
}

int pfdtd_destroy(pfdtd_solver* s) {
  if (!s) return PFDTD_OK;
  if (!s->parts.empty()) sync_all(s);
  free_partitions(s);
  if (s->d_pos0) { cudaSetDevice(s->stage_device); dev_free(s->d_pos0); dev_free(s->d_mat0); dev_free(s->d_cls0); }
  // s->comm is owned by the process-wide communicator cache (pfdtd_comm_init)
  if (s->ev_h0) cudaEventDestroy(s->ev_h0);
  if (s->ev_h1) cudaEventDestroy(s->ev_h1);
  delete s;
  return PFDTD_OK;
}

static int64_t* option_slot(pfdtd_solver* s, int option) {
  switch (option) {
    case PFDTD_OPT_MATIDX_AS_WRITTEN: return &s->opt_matidx_as_written;
    case PFDTD_OPT_SOFT_ACCUMULATE: return &s->opt_soft_accumulate;
    case PFDTD_OPT_KERNEL: return &s->opt_kernel;
    case PFDTD_OPT_GLOBAL_Z_FIRST: return &s->opt_global_z_first;
    case PFDTD_OPT_GLOBAL_Z_DIM: return &s->opt_global_z_dim;
    case PFDTD_OPT_DOUBLE_PAD_AS_WRITTEN: return &s->opt_double_pad;
    case PFDTD_OPT_USE_GRAPH: return &s->opt_use_graph;
    case PFDTD_OPT_OVERLAP: return &s->opt_overlap;
    case PFDTD_OPT_TMA_CHUNK: return &s->opt_tma_chunk;
    case PFDTD_OPT_TMA_TILE: return &s->opt_tma_tile;
    case PFDTD_OPT_TIME_KERNELS: return &s->opt_time_kernels;
    case PFDTD_OPT_TMA_HINTS: return &s->opt_tma_hints;
    case PFDTD_OPT_DIF_ORDER: return &s->opt_dif_order;
    case PFDTD_OPT_PEER_STORES: return &s->opt_peer_stores;
    case PFDTD_OPT_FUSE_SRCREC: return &s->opt_fuse_srcrec;
  }
  return nullptr;
}

int pfdtd_set_option(pfdtd_solver* s, int option, int64_t value) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  int64_t* slot = option_slot(s, option);
  PF_CHECK(slot, PFDTD_ERR_INVALID, "unknown option %d", option);
  // the filter states and row-segment entries of the partitions are laid out for one order
  PF_CHECK(!(option == PFDTD_OPT_DIF_ORDER && !s->parts.empty() && value != s->dif_order_built), PFDTD_ERR_INVALID,
           "PFDTD_OPT_DIF_ORDER cannot change (%d -> %lld) while partitions exist: set it before pfdtd_make_partition",
           s->dif_order_built, (long long)value);
  *slot = value;
  s->tables_dirty = true;
  s->fuse_dirty = true;
  if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
  return PFDTD_OK;
}

int pfdtd_set_scheme_coefficients(pfdtd_solver* s, const double* d4) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  if (d4) { for (int i = 0; i < 4; i++) s->dcoef[i] = d4[i]; s->dcoef_user = true; }
  else s->dcoef_user = false;
  s->tables_dirty = true;
  if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
  return PFDTD_OK;
}

int pfdtd_get_option(pfdtd_solver* s, int option, int64_t* value) {
  PF_CHECK(s && value, PFDTD_ERR_INVALID, "null argument");
  int64_t* slot = option_slot(s, option);
  PF_CHECK(slot, PFDTD_ERR_INVALID, "unknown option %d", option);
  *value = *slot;
  return PFDTD_OK;
}

int pfdtd_setup_mesh_device(pfdtd_solver* s, int device, uint8_t* d_bid, uint8_t* d_mat, uint32_t vx, uint32_t vy, uint32_t vz,
                            uint32_t block_x, uint32_t block_y, uint32_t block_z, uint32_t element_type, int dtype,
                            const void* params, const void* material_coefs, uint32_t n_unique_materials) {
  PF_CHECK(s && d_bid && d_mat && params && material_coefs, PFDTD_ERR_INVALID, "null argument");
  PF_CHECK(dtype == PFDTD_F32 || dtype == PFDTD_F64, PFDTD_ERR_INVALID, "bad dtype %d", dtype);
  PF_CHECK(vx && vy && vz && block_x && block_y && block_z, PFDTD_ERR_INVALID, "zero dimension");
  PF_CHECK(n_unique_materials >= 1, PFDTD_ERR_INVALID, "at least one material is required");
  PF_CHECK(element_type <= PFDTD_IWB, PFDTD_ERR_INVALID, "unknown update type %u", element_type);
  free_partitions(s);
  if (device < 0) PF_CUDA(cudaGetDevice(&device));
  PF_CUDA(cudaSetDevice(device));
  if (s->d_pos0) { dev_free(s->d_pos0); dev_free(s->d_mat0); dev_free(s->d_cls0); s->d_pos0 = s->d_mat0 = s->d_cls0 = nullptr; }
  s->dtype = dtype;
  s->element_type = (int)element_type;
  // scheme choice as in setupMesh: types 0,1,(3) -> Bilbao/forward, else Kowalczyk/centred (cudaMesh.cu:70-73,134-137)
  s->scheme = (element_type == 0 || element_type == 1) ? SCH_FORWARD : (element_type == PFDTD_SRL ? SCH_CENTRED : SCH_INTERP);
  s->bx = block_x; s->by = block_y; s->bz = block_z;
  // padWithZeros (cudaMesh.cu:253-304); setupMeshDouble pads y with block.x (cudaMesh.cu:106-111) when asked to
  uint32_t pby = (dtype == PFDTD_F64 && s->opt_double_pad) ? block_x : block_y;
  auto padded = [](uint32_t d, uint32_t b) { return d % b ? d + (b - d % b) : d; };
  const uint32_t nx = padded(vx, block_x), ny = padded(vy, pby), nz = padded(vz, block_z);
  const uint64_t n_new = (uint64_t)nx * ny * nz;
  // everything allocated below is released on every early return; the input volumes are adopted either way
  // (the reference frees them inside padWithZeros, cudaMesh.cu:290-291)
  struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* q : ptrs) dev_free(q); }
    void keep(void* q) { ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), q), ptrs.end()); }
  } scratch;
  scratch.ptrs.push_back(d_bid);
  scratch.ptrs.push_back(d_mat);
  uint8_t *np = nullptr, *nm = nullptr;
  unsigned long long* d_counts = nullptr;
  PF_TRY(dev_alloc((void**)&np, n_new));
  scratch.ptrs.push_back(np);
  PF_TRY(dev_alloc((void**)&nm, n_new));
  scratch.ptrs.push_back(nm);
  PF_CUDA(cudaMalloc(&d_counts, 2 * sizeof(unsigned long long)));
  scratch.ptrs.push_back(d_counts);
  PF_CUDA(cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), 0));
  PhaseTrace tr;
  tr.mark("alloc node volumes");
  const int skip_z0 = (s->opt_global_z_first == 0);   // only the global z=0 plane is dropped by the reference's copy
  PF_TRY(launch_prepare_nodes(d_bid, d_mat, np, nm, vx, vy, vz, nx, ny, nz, skip_z0, s->scheme == SCH_CENTRED, d_counts, 0));
  s->launch_count += (nx % 16 == 0) ? 1 : 3;
  unsigned long long h_counts[2] = {0, 0};
  PF_CUDA(cudaMemcpy(h_counts, d_counts, sizeof(h_counts), cudaMemcpyDeviceToHost));
  tr.mark("pad + translate + count");
  scratch.keep(d_bid); scratch.keep(d_mat);
  dev_free(d_bid);   // adopted, like the reference (cudaMesh.cu:290-291)
  dev_free(d_mat);
  tr.mark("free input volumes");
  // node classes: distinct node keys (position byte, material byte [, K12, K8]) -> one class byte per voxel
  {
    const uint32_t air_code = s->scheme == SCH_CENTRED ? 0x80u : 0x86u;
    const int interp = s->scheme == SCH_INTERP;
    const uint32_t air_key = interp ? (air_code | (12u << 16) | (8u << 20)) : air_code;
    const uint32_t cap = 65536;
    uint32_t* d_table = nullptr;
    uint32_t* d_count = nullptr;
    uint8_t* d_ids = nullptr;
    PF_CUDA(cudaMalloc(&d_table, cap * sizeof(uint32_t)));
    scratch.ptrs.push_back(d_table);
    PF_CUDA(cudaMalloc(&d_count, sizeof(uint32_t)));
    scratch.ptrs.push_back(d_count);
    PF_CUDA(cudaMalloc(&d_ids, cap));
    scratch.ptrs.push_back(d_ids);
    // Narrow first: one class per (position, material) pair.  When that needs more than a byte can name (or more lossy
    // classes than the kernels' shared filter table holds), classes are formed from the position part alone and the
    // material is looked up per boundary voxel (update_math.cuh WideArgs) -- "wide" mode.
    const bool centred_bytes = s->scheme == SCH_CENTRED;
    auto lossy = [centred_bytes](uint32_t k) { const uint32_t p = k & 0xffu; return centred_bytes ? (p & 7u) != 0u : (p & 0x7fu) < 6u; };
    s->d_cls0 = nullptr;
    s->wide = false;
    s->class_keys.assign({0u, air_key});                // class 0: solid, class 1: air
    s->dif_lo = 2; s->n_lossy = 0;
    for (int attempt = 0; attempt < 2 && s->d_cls0 == nullptr; attempt++) {
      const uint8_t* key_mat = attempt == 0 ? nm : nullptr;
      PF_CUDA(cudaMemset(d_table, 0xff, cap * sizeof(uint32_t)));
      PF_CUDA(cudaMemset(d_count, 0, sizeof(uint32_t)));
      PF_TRY(launch_mark_classes(np, key_mat, n_new, air_key, air_code, interp, nx, ny, nz, d_table, cap, d_count, 0));
      s->launch_count++;
      tr.mark("mark classes");
      std::vector<uint32_t> table(cap);
      uint32_t count = 0;
      PF_CUDA(cudaMemcpy(table.data(), d_table, cap * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      PF_CUDA(cudaMemcpy(&count, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
      std::vector<uint32_t> found;
      for (uint32_t k : table) if (k != 0xffffffffu) found.push_back(k);
      // lossless classes first, lossy boundary classes (K < 6 / a direction flag set) last: "is a filter boundary"
      // becomes one unsigned compare on the class byte
      std::sort(found.begin(), found.end(), [&](uint32_t a, uint32_t b) { return lossy(a) != lossy(b) ? lossy(b) : a < b; });
      uint32_t n_lossless = 0;
      for (uint32_t k : found) if (!lossy(k)) n_lossless++;
      const uint32_t n_lossy = (uint32_t)found.size() - n_lossless;
      // (the shared filter table only matters when filters are on: PFDTD_OPT_DIF_ORDER is read here, set it before setup_mesh)
      if (!(count + 2 <= 256 && count < cap / 2 && (n_lossy <= 64 || s->opt_dif_order == 0))) continue;       // does not fit this way
      std::vector<uint8_t> ids(cap, 0);
      for (uint32_t slot = 0; slot < cap; slot++)
        if (table[slot] != 0xffffffffu)
          ids[slot] = (uint8_t)(2 + (std::find(found.begin(), found.end(), table[slot]) - found.begin()));
      PF_CUDA(cudaMemcpy(d_ids, ids.data(), cap, cudaMemcpyHostToDevice));
      PF_TRY(dev_alloc((void**)&s->d_cls0, n_new));
      scratch.ptrs.push_back(s->d_cls0);
      PF_TRY(launch_assign_classes(np, key_mat, n_new, air_key, air_code, interp, nx, ny, nz, d_table, d_ids, cap, s->d_cls0, 0));
      PF_CUDA(cudaDeviceSynchronize());
      s->launch_count++;
      tr.mark("assign classes");
      for (uint32_t k : found) s->class_keys.push_back(k);
      s->dif_lo = 2 + n_lossless;
      s->n_lossy = n_lossy;
      s->wide = attempt == 1;
    }
  }
  tr.mark("class tables (host)");
  scratch.keep(np); scratch.keep(nm); scratch.keep(s->d_cls0);   // what the solver keeps; the rest goes with `scratch`
  s->d_pos0 = np; s->d_mat0 = nm;
  s->stage_device = device;
  s->X = nx; s->Y = ny; s->Z = nz;
  s->n_air = h_counts[0]; s->n_boundary = h_counts[1];
  if (dtype == PFDTD_F32) for (int i = 0; i < 4; i++) s->params[i] = (double)((const float*)params)[i];
  else for (int i = 0; i < 4; i++) s->params[i] = ((const double*)params)[i];
  if (s->scheme == SCH_INTERP && !s->dcoef_user) {
    // compact explicit family written so that the standard Courant numbers give exact
    // binary fractions: IISO (a=1/6, b=0): d = lam2*(1/3, 1/6, 0), d4 = 2 - 4 lam2;  IWB (a=1/4, b=1/16):
    // d = lam2*(1/4, 1/8, 1/16), d4 = 2 - 3.5 lam2
    const double l2 = s->params[1];
    if (element_type == PFDTD_IISO) { s->dcoef[0] = l2 / 3; s->dcoef[1] = l2 / 6; s->dcoef[2] = 0; s->dcoef[3] = 2 - 4 * l2; }
    else { s->dcoef[0] = l2 / 4; s->dcoef[1] = l2 / 8; s->dcoef[2] = l2 / 16; s->dcoef[3] = 2 - 3.5 * l2; }
  }
  s->n_unique = n_unique_materials;
  s->materials_host.assign((const unsigned char*)material_coefs,
                           (const unsigned char*)material_coefs + (size_t)n_unique_materials * 20 * esize(s));
  s->mesh_ready = true;
  s->cur = 0;
  s->past_direction = 1;
  s->srcrec_dirty = true;
  return PFDTD_OK;
}

int pfdtd_setup_mesh(pfdtd_solver* s, const uint8_t* h_bid, const uint8_t* h_mat, uint32_t vx, uint32_t vy, uint32_t vz,
                     uint32_t block_x, uint32_t block_y, uint32_t block_z, uint32_t element_type, int dtype, const void* params,
                     const void* material_coefs, uint32_t n_unique_materials) {
  PF_CHECK(s && h_bid && h_mat, PFDTD_ERR_INVALID, "null argument");
  int ndev = 0;
  PF_TRY(pfdtd_device_count(&ndev));
  PF_CHECK(ndev > 0, PFDTD_ERR_NO_DEVICE, "no CUDA device: libpfdtd_b200 has no CPU fallback");
  int device = 0;
  PF_CUDA(cudaGetDevice(&device));
  const size_t n = (size_t)vx * vy * vz;
  uint8_t *db = nullptr, *dm = nullptr;
  PhaseTrace tr;
  int rc = dev_alloc((void**)&db, n);
  if (rc == PFDTD_OK) rc = dev_alloc((void**)&dm, n);
  tr.mark("alloc input volumes");
  if (rc == PFDTD_OK) {
    cudaError_t e = cudaMemcpy(db, h_bid, n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dm, h_mat, n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("pfdtd_setup_mesh: H2D of the node volumes failed: %s", cudaGetErrorString(e)); rc = PFDTD_ERR_CUDA; }
  }
  if (rc != PFDTD_OK) { dev_free(db); dev_free(dm); return rc; }
  tr.mark("H2D bid + mat");
  return pfdtd_setup_mesh_device(s, device, db, dm, vx, vy, vz, block_x, block_y, block_z, element_type, dtype, params,
                                 material_coefs, n_unique_materials);
}

int pfdtd_voxelize_dims(const float* vertices, uint32_t n_vertices, float dx, uint32_t* vx, uint32_t* vy, uint32_t* vz) {
  PF_CHECK(vertices && n_vertices > 0 && vx && vy && vz, PFDTD_ERR_INVALID, "null argument");
  PF_CHECK(dx > 0.f, PFDTD_ERR_INVALID, "voxel size must be positive");
  float mx[3] = {vertices[0], vertices[1], vertices[2]};
  for (uint32_t v = 1; v < n_vertices; v++)
    for (int a = 0; a < 3; a++) mx[a] = std::max(mx[a], vertices[3 * v + a]);
  *vx = (uint32_t)std::ceil(mx[0] / dx) + 3; *vy = (uint32_t)std::ceil(mx[1] / dx) + 3; *vz = (uint32_t)std::ceil(mx[2] / dx) + 3;
  return PFDTD_OK;
}

int pfdtd_voxelize_device(int device, const float* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_triangles,
                          const uint8_t* triangle_material, float dx, uint8_t** d_bid, uint8_t** d_mat, uint32_t* vx, uint32_t* vy,
                          uint32_t* vz) {
  PF_CHECK(d_bid && d_mat && vx && vy && vz, PFDTD_ERR_INVALID, "null argument");
  int n = 0;
  PF_TRY(pfdtd_device_count(&n));
  PF_CHECK(n > 0, PFDTD_ERR_NO_DEVICE, "no CUDA device: the voxeliser has no CPU path");
  return voxelize_to_device(device, vertices, n_vertices, indices, n_triangles, triangle_material, dx, d_bid, d_mat, vx, vy, vz, nullptr);
}

int pfdtd_voxelize(const float* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_triangles,
                   const uint8_t* triangle_material, float dx, uint8_t* h_bid, uint8_t* h_mat) {
  PF_CHECK(h_bid && h_mat, PFDTD_ERR_INVALID, "null argument");
  uint8_t *d_bid = nullptr, *d_mat = nullptr;
  uint32_t vx = 0, vy = 0, vz = 0;
  PF_TRY(pfdtd_voxelize_device(-1, vertices, n_vertices, indices, n_triangles, triangle_material, dx, &d_bid, &d_mat, &vx, &vy, &vz));
  const size_t nvox = (size_t)vx * vy * vz;
  cudaError_t e = cudaMemcpy(h_bid, d_bid, nvox, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(h_mat, d_mat, nvox, cudaMemcpyDeviceToHost);
  cudaFree(d_bid);
  cudaFree(d_mat);
  PF_CUDA(e);
  return PFDTD_OK;
}

int pfdtd_partition_indexing(uint32_t dim_z, uint32_t n_partitions, uint32_t* first_slice, uint32_t* n_slices) {
  PF_CHECK(n_partitions >= 1 && first_slice && n_slices, PFDTD_ERR_INVALID, "bad argument");
  std::vector<int64_t> f, sz;
  partition_indexing(dim_z, n_partitions, f, sz);
  for (uint32_t i = 0; i < n_partitions; i++) { first_slice[i] = (uint32_t)f[i]; n_slices[i] = (uint32_t)sz[i]; }
  return PFDTD_OK;
}

int pfdtd_make_partition(pfdtd_solver* s, uint32_t n_partitions, const uint32_t* device_list) {
  PF_CHECK(s && s->mesh_ready, PFDTD_ERR_INVALID, "setup_mesh must be called before make_partition");
  PF_CHECK(n_partitions >= 1, PFDTD_ERR_INVALID, "need at least one partition");
  PF_CHECK(s->Z / n_partitions >= 1, PFDTD_ERR_INVALID, "more partitions (%u) than slices (%u)", n_partitions, s->Z);
  free_partitions(s);
  int ndev = 0;
  PF_TRY(pfdtd_device_count(&ndev));
  std::vector<int64_t> first, size;
  partition_indexing(s->Z, n_partitions, first, size);
  PhaseTrace tr;
  const size_t XY = (size_t)s->X * s->Y;
  const size_t es = esize(s);
  s->parts.resize(n_partitions);
  for (uint32_t k = 0; k < n_partitions; k++) {
    Partition& p = s->parts[k];
    p.device = device_list ? (int)device_list[k] : (int)k;
    PF_CHECK(p.device < ndev, PFDTD_ERR_INVALID, "partition %u wants device %d but only %d devices exist", k, p.device, ndev);
    p.first = first[k];
    p.size = size[k];
  }
  // peer access between devices that share an interface (NVLink P2P; the reference never enables it).  peer_ok says
  // whether a kernel running for partition a may store into partition b's memory (same device, or peer access on).
  s->peer_ok.assign((size_t)n_partitions * n_partitions, 0);
  for (uint32_t k = 0; k < n_partitions; k++) s->peer_ok[(size_t)k * n_partitions + k] = 1;
  for (uint32_t k = 0; k + 1 < n_partitions; k++) {
    const int a = s->parts[k].device, b = s->parts[k + 1].device;
    bool ab = a == b, ba = a == b;
    if (a != b) {
      int can_ab = 0, can_ba = 0;
      cudaDeviceCanAccessPeer(&can_ab, a, b);
      cudaDeviceCanAccessPeer(&can_ba, b, a);
      if (can_ab) {
        cudaSetDevice(a);
        const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
        ab = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
        if (e != cudaSuccess) cudaGetLastError();
      }
      if (can_ba) {
        cudaSetDevice(b);
        const cudaError_t e = cudaDeviceEnablePeerAccess(a, 0);
        ba = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
        if (e != cudaSuccess) cudaGetLastError();
      }
    }
    s->peer_ok[(size_t)k * n_partitions + (k + 1)] = ab;
    s->peer_ok[(size_t)(k + 1) * n_partitions + k] = ba;
  }
  for (uint32_t k = 0; k < n_partitions; k++) {
    Partition& p = s->parts[k];
    PF_CUDA(cudaSetDevice(p.device));
    const size_t nelem = (size_t)p.size * XY;
    if (n_partitions == 1 && p.device == s->stage_device) {
      p.pos = s->d_pos0; p.mat = s->d_mat0; p.cls = s->d_cls0; p.owns_nodes = true;   // adopt (cudaMesh.h:697-700)
      s->d_pos0 = s->d_mat0 = s->d_cls0 = nullptr;
    } else {
      PF_TRY(dev_alloc((void**)&p.pos, nelem));
      PF_TRY(dev_alloc((void**)&p.mat, nelem));
      PF_CUDA(cudaMemcpyPeer(p.pos, p.device, s->d_pos0 + (size_t)p.first * XY, s->stage_device, nelem));
      PF_CUDA(cudaMemcpyPeer(p.mat, p.device, s->d_mat0 + (size_t)p.first * XY, s->stage_device, nelem));
      if (s->d_cls0) {
        PF_TRY(dev_alloc((void**)&p.cls, nelem));
        PF_CUDA(cudaMemcpyPeer(p.cls, p.device, s->d_cls0 + (size_t)p.first * XY, s->stage_device, nelem));
      }
    }
    if (p.cls) {
      const size_t nc = s->class_keys.size();
      PF_CUDA(cudaMalloc(&p.class_table, nc * class_entry_bytes(s->dtype)));
      PF_CUDA(cudaMalloc(&p.d_class_keys, nc * sizeof(uint32_t)));
      PF_CUDA(cudaMemcpy(p.d_class_keys, s->class_keys.data(), nc * sizeof(uint32_t), cudaMemcpyHostToDevice));
      if (s->wide && s->n_lossy) {
        std::vector<uint32_t> wk((size_t)s->n_lossy * s->n_unique);
        for (uint32_t i = 0; i < s->n_lossy; i++)
          for (uint32_t m = 0; m < s->n_unique; m++) wk[(size_t)i * s->n_unique + m] = s->class_keys[s->dif_lo + i] | (m << 8);
        PF_CUDA(cudaMalloc(&p.d_wide_keys, wk.size() * sizeof(uint32_t)));
        PF_CUDA(cudaMemcpy(p.d_wide_keys, wk.data(), wk.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        PF_CUDA(cudaMalloc(&p.wide_class_table, wk.size() * class_entry_bytes(s->dtype)));
        if (s->opt_dif_order > 0) PF_CUDA(cudaMalloc(&p.wide_dif_table, wk.size() * dif_entry_bytes(s->dtype)));
      }
    }
    for (int b = 0; b < 2; b++) {
      PF_TRY(dev_alloc(&p.P[b], nelem * es));
      PF_CUDA(cudaMemset(p.P[b], 0, nelem * es));
    }
    tr.mark("partition: nodes + fields");
    PF_CUDA(cudaMalloc(&p.materials, s->materials_host.size()));
    PF_CUDA(cudaMemcpy(p.materials, s->materials_host.data(), s->materials_host.size(), cudaMemcpyHostToDevice));
    PF_CUDA(cudaMalloc(&p.d_step, 4 * sizeof(int)));
    PF_CUDA(cudaMemset(p.d_step, 0, 4 * sizeof(int)));
    int lo_pri = 0, hi_pri = 0;
    PF_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    PF_CUDA(cudaStreamCreateWithPriority(&p.s_main, cudaStreamNonBlocking, lo_pri));
    PF_CUDA(cudaStreamCreateWithPriority(&p.s_edge, cudaStreamNonBlocking, hi_pri));
    for (cudaEvent_t* e : {&p.ev_src, &p.ev_int, &p.ev_edge}) PF_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    PF_CUDA(cudaEventCreate(&p.ev_t0));
    PF_CUDA(cudaEventCreate(&p.ev_t1));
    // prime the cross-stream events so the first step's waits are satisfied
    PF_CUDA(cudaEventRecord(p.ev_int, p.s_main));
    PF_CUDA(cudaEventRecord(p.ev_edge, p.s_edge));
    // kernel choice
    const int nplanes = (int)p.size - 2;
    p.use_tma = (s->opt_kernel != KERNEL_PLAIN) && tma_supported((int)s->X, (int)s->Y, s->dtype) && nplanes >= 1 && p.cls != nullptr;
    PF_CHECK(!(s->scheme == SCH_INTERP && p.cls == nullptr), PFDTD_ERR_INVALID,
             "the interpolated schemes need <= 256 node classes (this mesh has more)");
    PF_CHECK(!(s->opt_kernel == KERNEL_TMA && !p.use_tma), PFDTD_ERR_INVALID,
             "TMA kernel requested but mesh %ux%u (slab of %lld slices, %zu node classes) is not supported by it", s->X, s->Y,
             (long long)p.size, s->class_keys.size());
    if (s->opt_dif_order > 0) {
      PF_CHECK(s->opt_dif_order <= PFDTD_DIF_MAX_ORDER_HOST, PFDTD_ERR_INVALID, "filter order %lld > %d", (long long)s->opt_dif_order,
               PFDTD_DIF_MAX_ORDER_HOST);
      PF_CHECK(p.use_tma, PFDTD_ERR_INVALID, "filter (DIF) boundaries need the TMA kernel (X %% 16 == 0 and <= 256 node classes)");
      PF_CHECK(s->n_lossy <= 64, PFDTD_ERR_INVALID,
               "filter (DIF) boundaries: %u lossy node classes do not fit the kernels' table of 64 -- set PFDTD_OPT_DIF_ORDER before "
               "pfdtd_setup_mesh so that the mesh is classified by position and material separately", s->n_lossy);
      const int segs = ((int)s->X + 127) / 128;
      const size_t n_seg = (size_t)p.size * s->Y * segs;
      // per-segment counts -> entries, on the device (three small kernels); only the total comes back
      const int n_cols = (int)s->Y * segs;
      uint32_t* d_counts = nullptr;
      uint32_t* d_cols = nullptr;
      unsigned long long* d_nb = nullptr;
      PF_TRY(dev_alloc((void**)&p.dif_rowbase, n_seg * 2 * sizeof(uint32_t)));
      PF_TRY(dev_alloc((void**)&d_counts, n_seg * sizeof(uint32_t)));
      PF_CUDA(cudaMalloc(&d_cols, (size_t)n_cols * sizeof(uint32_t)));
      PF_CUDA(cudaMalloc(&d_nb, sizeof(unsigned long long)));
      PF_TRY(launch_count_dif_segments(p.cls, (int)s->X, (int)s->Y, (int)p.size, s->dif_lo, d_counts, 0));
      PF_TRY(launch_build_dif_entries(d_counts, n_cols, (int)p.size, d_cols, d_nb, p.dif_rowbase, 0));
      unsigned long long run = 0;
      PF_CUDA(cudaMemcpy(&run, d_nb, sizeof(run), cudaMemcpyDeviceToHost));
      dev_free(d_counts);
      PF_CUDA(cudaFree(d_cols));
      PF_CUDA(cudaFree(d_nb));
      PF_CHECK(run < 0x7fffffffull, PFDTD_ERR_INVALID, "too many boundary voxels in one slab");
      p.dif_nb = (uint32_t)run;
      s->launch_count += 3;
      const size_t sb = (size_t)std::max<uint32_t>(p.dif_nb, 1) * dif_state_pad((int)s->opt_dif_order) * es;
      PF_TRY(dev_alloc(&p.dif_state, sb));
      PF_CUDA(cudaMemset(p.dif_state, 0, sb));
      PF_CUDA(cudaMalloc(&p.dif_table, std::max<uint32_t>(s->n_lossy, 1) * dif_entry_bytes(s->dtype)));
      s->launch_count++;
    }
    tr.mark("partition: streams, filters");
    if (p.use_tma) {
      int64_t want_tile = s->opt_tma_tile;
      PF_TRY(tma_pick_config(s->dtype, s->scheme, (int)s->opt_dif_order, s->wide, (int)s->X, (int)s->Y, nplanes, p.device, want_tile, s->opt_tma_chunk,
                             &p.cfg_full));
      PF_TRY(tma_pick_config(s->dtype, s->scheme, (int)s->opt_dif_order, s->wide, (int)s->X, (int)s->Y, std::max(1, nplanes - 2), p.device,
                             p.cfg_full.tile + 1, s->opt_tma_chunk, &p.cfg_int));
      p.cfg_edge = TmaConfig{p.cfg_full.tile, 1, p.cfg_full.occupancy};
      const int tail_tile = tma_tail_tile(s->dtype, s->scheme);
      PF_TRY(tma_pick_config(s->dtype, s->scheme, (int)s->opt_dif_order, s->wide, (int)s->X, (int)s->Y, nplanes, p.device, tail_tile + 1,
                             s->opt_tma_chunk, &p.cfg_tail_full));
      for (int c = 0; c < 2; c++) {
        PF_TRY(tma_encode_maps(&p.maps[c], s->dtype, p.cfg_full.tile, p.P[c], p.P[1 - c], p.cls, (int)s->X, (int)s->Y, (int)p.size));
        PF_TRY(tma_encode_maps(&p.maps_tail[c], s->dtype, tail_tile, p.P[c], p.P[1 - c], p.cls, (int)s->X, (int)s->Y, (int)p.size));
      }
    }
  }
  tr.mark("partition: tile config + maps");
  if (s->d_pos0) {   // free the staging volumes (cudaMesh.h:704-707)
    PF_CUDA(cudaSetDevice(s->stage_device));
    dev_free(s->d_pos0);
    dev_free(s->d_mat0);
    dev_free(s->d_cls0);
    s->d_pos0 = s->d_mat0 = s->d_cls0 = nullptr;
  }
  s->tables_dirty = true;
  s->dif_order_built = (int)s->opt_dif_order;
  s->mesh_ready = false;   // staging consumed; setup_mesh again before re-partitioning
  s->cur = 0;
  s->srcrec_dirty = true;
  return PFDTD_OK;
}

int pfdtd_get_dims(pfdtd_solver* s, uint32_t* dim_x, uint32_t* dim_y, uint32_t* dim_z) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  if (dim_x) *dim_x = s->X;
  if (dim_y) *dim_y = s->Y;
  if (dim_z) *dim_z = s->Z;
  return PFDTD_OK;
}

int pfdtd_get_counts(pfdtd_solver* s, uint64_t* n_elements, uint64_t* n_air, uint64_t* n_boundary) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  if (n_elements) *n_elements = (uint64_t)s->X * s->Y * s->Z;
  if (n_air) *n_air = s->n_air;
  if (n_boundary) *n_boundary = s->n_boundary;
  return PFDTD_OK;
}

int pfdtd_get_num_partitions(pfdtd_solver* s, uint32_t* n) {
  PF_CHECK(s && n, PFDTD_ERR_INVALID, "null argument");
  *n = (uint32_t)s->parts.size();
  return PFDTD_OK;
}

int pfdtd_get_partition(pfdtd_solver* s, uint32_t k, uint32_t* first_slice, uint32_t* n_slices, uint32_t* device) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  PF_CHECK(k < s->parts.size(), PFDTD_ERR_RANGE, "partition %u out of range (%zu)", k, s->parts.size());
  if (first_slice) *first_slice = (uint32_t)s->parts[k].first;
  if (n_slices) *n_slices = (uint32_t)s->parts[k].size;
  if (device) *device = (uint32_t)s->parts[k].device;
  return PFDTD_OK;
}

int pfdtd_get_element_idx_and_partition(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z, int* partition, int64_t* element) {
  PF_CHECK(s && partition && element, PFDTD_ERR_INVALID, "null argument");
  *partition = -1;
  *element = -1;
  for (size_t k = 0; k < s->parts.size(); k++) {
    const Partition& p = s->parts[k];
    if ((int64_t)z > p.first + p.size - 1) continue;
    if ((int64_t)z < p.first) break;
    *element = ((int64_t)z - p.first) * (int64_t)s->X * s->Y + (int64_t)y * s->X + x;
    *partition = (int)k;
    break;
  }
  return PFDTD_OK;
}

int pfdtd_export_partition_nodes(pfdtd_solver* s, uint32_t k, uint8_t* h_pos, uint8_t* h_mat) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  PF_CHECK(k < s->parts.size(), PFDTD_ERR_RANGE, "partition %u out of range", k);
  Partition& p = s->parts[k];
  PF_CUDA(cudaSetDevice(p.device));
  const size_t n = (size_t)p.size * s->X * s->Y;
  if (h_pos) PF_CUDA(cudaMemcpy(h_pos, p.pos, n, cudaMemcpyDeviceToHost));
  if (h_mat) PF_CUDA(cudaMemcpy(h_mat, p.mat, n, cudaMemcpyDeviceToHost));
  return PFDTD_OK;
}

int pfdtd_export_partition_pressure(pfdtd_solver* s, uint32_t k, int which, void* h_out) {
  PF_CHECK(s && h_out, PFDTD_ERR_INVALID, "null argument");
  PF_CHECK(k < s->parts.size(), PFDTD_ERR_RANGE, "partition %u out of range", k);
  PF_TRY(sync_all(s));
  Partition& p = s->parts[k];
  PF_CUDA(cudaSetDevice(p.device));
  const size_t n = (size_t)p.size * s->X * s->Y * esize(s);
  PF_CUDA(cudaMemcpy(h_out, p.P[which ? 1 - s->cur : s->cur], n, cudaMemcpyDeviceToHost));
  return PFDTD_OK;
}

// planes of partition k that belong to it (halo planes belong to the neighbours): local [lo, hi)
static void owned_planes(const pfdtd_solver* s, size_t k, int64_t* lo, int64_t* hi) {
  *lo = k == 0 ? 0 : 1;
  *hi = k + 1 == s->parts.size() ? s->parts[k].size : s->parts[k].size - 1;
}

int pfdtd_capture_slice(pfdtd_solver* s, uint32_t slice, uint32_t orientation, void* h_pressure, uint8_t* h_position) {
  PF_CHECK(s && h_pressure, PFDTD_ERR_INVALID, "null argument");
  PF_CHECK(!s->parts.empty(), PFDTD_ERR_INVALID, "pfdtd_capture_slice before pfdtd_make_partition");
  PF_CHECK(orientation <= 2, PFDTD_ERR_INVALID, "orientation %u (0 xy, 1 xz, 2 yz)", orientation);
  const uint32_t lim = orientation == 0 ? s->Z : (orientation == 1 ? s->Y : s->X);
  PF_CHECK(slice < lim, PFDTD_ERR_RANGE, "slice %u out of bounds %u", slice, lim);
  PF_TRY(sync_all(s));
  const size_t es = esize(s);
  const size_t XY = (size_t)s->X * s->Y;
  if (orientation == 0) {                                  // one contiguous plane of the first partition that holds z
    for (size_t k = 0; k < s->parts.size(); k++) {
      Partition& p = s->parts[k];
      if ((int64_t)slice < p.first || (int64_t)slice >= p.first + p.size) continue;
      PF_CUDA(cudaSetDevice(p.device));
      const size_t off = (size_t)(slice - p.first) * XY;
      PF_CUDA(cudaMemcpy(h_pressure, (const char*)p.P[s->cur] + off * es, XY * es, cudaMemcpyDeviceToHost));
      if (h_position) PF_CUDA(cudaMemcpy(h_position, p.pos + off, XY, cudaMemcpyDeviceToHost));
      return PFDTD_OK;
    }
    PF_CHECK(false, PFDTD_ERR_RANGE, "slice %u is in no partition", slice);
  }
  const size_t w = orientation == 1 ? s->X : s->Y;
  for (size_t k = 0; k < s->parts.size(); k++) {
    Partition& p = s->parts[k];
    int64_t lo, hi;
    owned_planes(s, k, &lo, &hi);
    if (hi <= lo) continue;
    PF_CUDA(cudaSetDevice(p.device));
    const size_t n = w * (size_t)(hi - lo);
    void* d_p = nullptr;
    uint8_t* d_pos = nullptr;
    PF_CUDA(cudaMalloc(&d_p, n * es));
    if (h_position) PF_CUDA(cudaMalloc((void**)&d_pos, n));
    int rc = launch_capture_slice(s->dtype, p.P[s->cur], p.pos, d_p, d_pos, s->X, s->Y, (uint32_t)lo, (uint32_t)(hi - lo), slice,
                                  (int)orientation, p.s_main);
    s->launch_count++;
    cudaError_t e = rc == PFDTD_OK ? cudaStreamSynchronize(p.s_main) : cudaSuccess;
    const size_t row0 = (size_t)(p.first + lo) * w;
    if (rc == PFDTD_OK && e == cudaSuccess) e = cudaMemcpy((char*)h_pressure + row0 * es, d_p, n * es, cudaMemcpyDeviceToHost);
    if (rc == PFDTD_OK && e == cudaSuccess && h_position) e = cudaMemcpy(h_position + row0, d_pos, n, cudaMemcpyDeviceToHost);
    cudaFree(d_p);
    if (d_pos) cudaFree(d_pos);
    if (rc != PFDTD_OK) return rc;
    PF_CUDA(e);
  }
  return PFDTD_OK;
}

int pfdtd_capture_mesh(pfdtd_solver* s, void* h_field) {
  PF_CHECK(s && h_field, PFDTD_ERR_INVALID, "null argument");
  PF_CHECK(!s->parts.empty(), PFDTD_ERR_INVALID, "pfdtd_capture_mesh before pfdtd_make_partition");
  PF_TRY(sync_all(s));
  const size_t es = esize(s);
  const size_t XY = (size_t)s->X * s->Y;
  for (size_t k = 0; k < s->parts.size(); k++) {
    Partition& p = s->parts[k];
    int64_t lo, hi;
    owned_planes(s, k, &lo, &hi);
    if (hi <= lo) continue;
    PF_CUDA(cudaSetDevice(p.device));
    PF_CUDA(cudaMemcpy((char*)h_field + (size_t)(p.first + lo) * XY * es, (const char*)p.P[s->cur] + (size_t)lo * XY * es,
                       (size_t)(hi - lo) * XY * es, cudaMemcpyDeviceToHost));
  }
  return PFDTD_OK;
}

int pfdtd_get_device_pointers(pfdtd_solver* s, uint32_t k, void** d_pressure, void** d_pressure_past, uint8_t** d_position_idx,
                              uint8_t** d_material_idx) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  PF_CHECK(k < s->parts.size(), PFDTD_ERR_RANGE, "partition %u out of range", k);
  PF_TRY(sync_all(s));
  Partition& p = s->parts[k];
  if (d_pressure) *d_pressure = p.P[s->cur];
  if (d_pressure_past) *d_pressure_past = p.P[1 - s->cur];
  if (d_position_idx) *d_position_idx = p.pos;
  if (d_material_idx) *d_material_idx = p.mat;
  return PFDTD_OK;
}

// ---- single samples --------------------------------------------------------------------------------------
static int write_sample(pfdtd_solver* s, size_t k, int64_t e, double value, bool add) {
  Partition& p = s->parts[k];
  PF_CUDA(cudaSetDevice(p.device));
  char* dst = (char*)p.P[s->cur] + (size_t)e * esize(s);
  if (s->dtype == PFDTD_F32) {
    float v = (float)value;
    if (add && s->opt_soft_accumulate) { float c = 0; PF_CUDA(cudaMemcpy(&c, dst, 4, cudaMemcpyDeviceToHost)); v += c; }
    PF_CUDA(cudaMemcpy(dst, &v, 4, cudaMemcpyHostToDevice));
  } else {
    double v = value;
    if (add && s->opt_soft_accumulate) { double c = 0; PF_CUDA(cudaMemcpy(&c, dst, 8, cudaMemcpyDeviceToHost)); v += c; }
    PF_CUDA(cudaMemcpy(dst, &v, 8, cudaMemcpyHostToDevice));
  }
  return PFDTD_OK;
}

static int set_or_add(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z, double value, bool add) {
  PF_CHECK(s && !s->parts.empty(), PFDTD_ERR_INVALID, "no partitions");
  PF_TRY(sync_all(s));
  // every partition containing the slice (cudaMesh.h:321-338 / 349-369)
  for (size_t k = 0; k < s->parts.size(); k++) {
    const Partition& p = s->parts[k];
    if ((int64_t)z > p.first + p.size - 1) continue;
    if ((int64_t)z < p.first) break;
    int64_t e;
    PF_TRY(elem_of(s, x, y, z, k, &e));
    PF_TRY(write_sample(s, k, e, value, add));
  }
  return PFDTD_OK;
}

int pfdtd_set_sample(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z, double value) { return set_or_add(s, x, y, z, value, false); }
int pfdtd_add_sample(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z, double value) { return set_or_add(s, x, y, z, value, true); }

static int read_sample(pfdtd_solver* s, size_t k, int64_t e, double* value) {
  Partition& p = s->parts[k];
  PF_CUDA(cudaSetDevice(p.device));
  const char* src = (const char*)p.P[s->cur] + (size_t)e * esize(s);
  if (s->dtype == PFDTD_F32) { float v = 0; PF_CUDA(cudaMemcpy(&v, src, 4, cudaMemcpyDeviceToHost)); *value = v; }
  else { PF_CUDA(cudaMemcpy(value, src, 8, cudaMemcpyDeviceToHost)); }
  return PFDTD_OK;
}

int pfdtd_get_sample(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t z, double* value) {
  PF_CHECK(s && value && !s->parts.empty(), PFDTD_ERR_INVALID, "bad argument");
  PF_TRY(sync_all(s));
  *value = 0.0;
  int k = -1;
  int64_t e = -1;
  PF_TRY(pfdtd_get_element_idx_and_partition(s, x, y, z, &k, &e));
  if (k < 0) return PFDTD_OK;   // the reference returns 0 for a slice nobody holds (cudaMesh.h:387-404)
  PF_CHECK(x < s->X && y < s->Y, PFDTD_ERR_RANGE, "coordinate outside mesh");
  return read_sample(s, (size_t)k, e, value);
}

int pfdtd_set_sample_at(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t local_z, uint32_t partition, double value) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  PF_CHECK(partition < s->parts.size(), PFDTD_ERR_RANGE, "partition %u out of range", partition);
  PF_CHECK(x < s->X && y < s->Y && (int64_t)local_z < s->parts[partition].size, PFDTD_ERR_RANGE, "coordinate outside partition");
  PF_TRY(sync_all(s));
  return write_sample(s, partition, (int64_t)local_z * s->X * s->Y + (int64_t)y * s->X + x, value, false);
}

int pfdtd_get_sample_at(pfdtd_solver* s, uint32_t x, uint32_t y, uint32_t local_z, uint32_t partition, double* value) {
  PF_CHECK(s && value, PFDTD_ERR_INVALID, "null argument");
  PF_CHECK(partition < s->parts.size(), PFDTD_ERR_RANGE, "partition %u out of range", partition);
  PF_CHECK(x < s->X && y < s->Y && (int64_t)local_z < s->parts[partition].size, PFDTD_ERR_RANGE, "coordinate outside partition");
  PF_TRY(sync_all(s));
  return read_sample(s, partition, (int64_t)local_z * s->X * s->Y + (int64_t)y * s->X + x, value);
}

int pfdtd_switch_halos(pfdtd_solver* s) {
  PF_CHECK(s && !s->parts.empty(), PFDTD_ERR_INVALID, "no partitions");
  PF_TRY(sync_all(s));
  for (size_t k = 0; k + 1 < s->parts.size(); k++) {
    PF_TRY(enqueue_halo_local_up(s, k, s->cur));
    PF_TRY(enqueue_halo_local_down(s, k, s->cur));
  }
  PF_TRY(enqueue_halo_external(s, s->cur));
  return sync_all(s);
}

int pfdtd_flip_pressure_pointers(pfdtd_solver* s) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  s->cur = 1 - s->cur;
  return PFDTD_OK;
}

int pfdtd_reset_pressures(pfdtd_solver* s) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  PF_TRY(sync_all(s));
  for (auto& p : s->parts) {
    PF_CUDA(cudaSetDevice(p.device));
    const size_t n = (size_t)p.size * s->X * s->Y * esize(s);
    PF_CUDA(cudaMemset(p.P[0], 0, n));
    PF_CUDA(cudaMemset(p.P[1], 0, n));
    if (p.dif_state) PF_CUDA(cudaMemset(p.dif_state, 0, (size_t)std::max<uint32_t>(p.dif_nb, 1) * dif_state_pad((int)s->opt_dif_order) * esize(s)));
  }
  return PFDTD_OK;
}

// ---- sources / receivers ----------------------------------------------------------------------------------
// height of the whole domain: the local mesh, or the taller domain this process owns a slab of (PFDTD_OPT_GLOBAL_Z_DIM)
static int64_t global_z_dim(const pfdtd_solver* s) { return s->opt_global_z_dim > 0 ? (int64_t)s->opt_global_z_dim : (int64_t)s->Z; }

int pfdtd_set_sources(pfdtd_solver* s, uint32_t n, const int32_t* xyz, const int32_t* src_types, const void* samples,
                      uint32_t n_steps) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  PF_CHECK(n == 0 || (xyz && src_types && samples), PFDTD_ERR_INVALID, "null source arrays");
  for (uint32_t i = 0; i < n; i++)
    PF_CHECK(xyz[3 * i] >= 0 && xyz[3 * i + 1] >= 0 && xyz[3 * i + 2] >= 0 &&
                 (s->X == 0 || ((uint32_t)xyz[3 * i] < s->X && (uint32_t)xyz[3 * i + 1] < s->Y && (int64_t)xyz[3 * i + 2] < global_z_dim(s))),
             PFDTD_ERR_RANGE, "source %u at (%d,%d,%d) is outside the mesh", i, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  s->n_src = n;
  s->src_steps = n_steps;
  s->src_xyz.assign(xyz, xyz + 3 * (size_t)n);
  s->src_type.assign(src_types, src_types + n);
  s->src_samples.assign((const unsigned char*)samples, (const unsigned char*)samples + (size_t)n * n_steps * esize(s));
  s->src_dirty = true;
  return PFDTD_OK;
}

int pfdtd_set_receivers(pfdtd_solver* s, uint32_t n, const int32_t* xyz) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  PF_CHECK(n == 0 || xyz, PFDTD_ERR_INVALID, "null receiver array");
  for (uint32_t i = 0; i < n; i++)
    PF_CHECK(xyz[3 * i] >= 0 && xyz[3 * i + 1] >= 0 && xyz[3 * i + 2] >= 0 &&
                 (s->X == 0 || ((uint32_t)xyz[3 * i] < s->X && (uint32_t)xyz[3 * i + 1] < s->Y && (int64_t)xyz[3 * i + 2] < global_z_dim(s))),
             PFDTD_ERR_RANGE, "receiver %u at (%d,%d,%d) is outside the mesh", i, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  s->n_rec = n;
  s->rec_xyz.assign(xyz, xyz + 3 * (size_t)n);
  s->srcrec_dirty = true;
  return PFDTD_OK;
}

int pfdtd_reserve_steps(pfdtd_solver* s, uint32_t n_steps) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  if (n_steps > s->rec_cap) s->rec_cap = n_steps;   // prepare_srcrec grows the buffers and keeps what they hold
  return PFDTD_OK;
}

// ---- stepping ----------------------------------------------------------------------------------------------
int pfdtd_enqueue_steps(pfdtd_solver* s, uint32_t first_step, uint32_t n_steps) {
  PF_CHECK(s && !s->parts.empty(), PFDTD_ERR_INVALID, "make_partition must be called before stepping");
  if (first_step + n_steps > s->rec_cap) PF_TRY(pfdtd_reserve_steps(s, first_step + n_steps));
  PF_TRY(prepare_srcrec(s));
  PF_TRY(ensure_class_tables(s));
  const bool single = (s->parts.size() == 1 && !s->comm);
  const bool timed = s->opt_time_kernels != 0;
  for (auto& p : s->parts) {
    for (cudaEvent_t e : p.kev) cudaEventDestroy(e);
    p.kev.clear();
    p.kev_planes.clear();
  }
  if (s->comm && !s->ev_h0) {
    PF_CUDA(cudaSetDevice(s->parts[0].device));
    PF_CUDA(cudaEventCreate(&s->ev_h0));
    PF_CUDA(cudaEventCreate(&s->ev_h1));
  }
  PF_TRY(set_step_counters(s, (int)first_step, (int)first_step, (int)(first_step + n_steps) - 1));
  for (auto& p : s->parts) {
    PF_CUDA(cudaSetDevice(p.device));
    PF_CUDA(cudaEventRecord(p.ev_t0, p.s_main));
  }
  uint32_t done = 0;
  const bool fused = single && s->parts[0].fused;
  if (fused && n_steps > 0) {   // sources of the first step; every later injection is done by the update launches
    Partition& p = s->parts[0];
    PF_CUDA(cudaSetDevice(p.device));
    PF_TRY(launch_srcrec_for(s, p, p.s_main, 0, 1, 0));
  }
  const uint64_t per_step = fused ? 1 : 2;   // launches per step of a single slab
  // CUDA-graph replay of two-step blocks (single partition, untimed): the launch-bound regime of small meshes
  if (single && s->opt_use_graph && !timed && n_steps >= 8) {
    Partition& p = s->parts[0];
    PF_CUDA(cudaSetDevice(p.device));
    if (s->cur != 0) { PF_TRY(enqueue_one_step(s, false, false)); done++; }
    if (!s->graph_exec) {
      cudaGraph_t g = nullptr;
      PF_CUDA(cudaStreamBeginCapture(p.s_main, cudaStreamCaptureModeThreadLocal));
      int rc = enqueue_one_step(s, false, false);
      if (rc == PFDTD_OK) rc = enqueue_one_step(s, false, false);
      cudaError_t ce = cudaStreamEndCapture(p.s_main, &g);
      s->launch_count -= 2 * per_step;   // capture recorded, nothing ran
      if (rc != PFDTD_OK) { if (g) cudaGraphDestroy(g); return rc; }
      PF_CUDA(ce);
      PF_CUDA(cudaGraphInstantiate(&s->graph_exec, g, 0));
      PF_CUDA(cudaGraphDestroy(g));
    }
    while (n_steps - done >= 2) {
      PF_CUDA(cudaGraphLaunch(s->graph_exec, p.s_main));
      s->launch_count += 2 * per_step;
      done += 2;
    }
  }
  for (; done < n_steps; done++) PF_TRY(enqueue_one_step(s, timed, s->comm != nullptr && done + 1 == n_steps));
  // flush: record the receivers of the last step
  for (auto& p : s->parts) {
    PF_CUDA(cudaSetDevice(p.device));
    if (single) {
      if (!fused) PF_TRY(launch_srcrec_for(s, p, p.s_main, 1, 0, 0));   // fused: the last update launch has recorded them
      PF_CUDA(cudaEventRecord(p.ev_t1, p.s_main));
    } else {
      // the last step's edge planes and halo traffic, and the neighbours' halo copies into this partition's end planes
      PF_CUDA(cudaStreamWaitEvent(p.s_main, p.ev_edge, 0));
      const size_t k = (size_t)(&p - &s->parts[0]);
      if (k > 0) PF_CUDA(cudaStreamWaitEvent(p.s_main, s->parts[k - 1].ev_edge, 0));
      if (k + 1 < s->parts.size()) PF_CUDA(cudaStreamWaitEvent(p.s_main, s->parts[k + 1].ev_edge, 0));
      PF_TRY(launch_srcrec_for(s, p, p.s_main, 1, 0, 0));
      PF_CUDA(cudaEventRecord(p.ev_t1, p.s_main));
    }
  }
  return PFDTD_OK;
}

int pfdtd_sync(pfdtd_solver* s) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  PF_TRY(sync_all(s));
  if (s->d_halo_flags) {   // a halo wait that timed out (srcrec_kernels.cu halo_wait): a neighbour process never delivered
    int err = 0;
    PF_CUDA(cudaSetDevice(s->parts[0].device));
    PF_CUDA(cudaMemcpy(&err, s->d_halo_flags + 6 /* HALO_ERR */, sizeof(int), cudaMemcpyDeviceToHost));
    PF_CHECK(err == 0, PFDTD_ERR_COMM, "halo hand-over timed out: a neighbour process did not deliver its plane (results are invalid)");
  }
  // Update launches are timed one by one on their own streams.  A slab with neighbours runs three per step: the two
  // one-plane edge launches (edge stream) and the interior launch (main stream), which overlap -- so the sum of all
  // three is not a time anything waited for.  Reported: the bulk launches (everything that is not a one-plane edge
  // launch: the dominant kernel, and what a roofline is computed from) and the edge launches separately.
  float total = 0, bulk = 0, edge = 0;
  uint32_t n_bulk = 0, n_edge = 0;
  uint64_t bulk_planes = 0;
  for (auto& p : s->parts) {
    PF_CUDA(cudaSetDevice(p.device));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.ev_t0, p.ev_t1) == cudaSuccess) total = std::max(total, ms);
    else cudaGetLastError();
    float bsum = 0, esum = 0;
    uint32_t nb = 0, ne = 0;
    uint64_t planes = 0;
    const bool has_edges = s->parts.size() > 1 || s->comm;
    for (size_t i = 0; i + 1 < p.kev.size(); i += 2) {
      float k = 0;
      if (cudaEventElapsedTime(&k, p.kev[i], p.kev[i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
      const int pl = i / 2 < p.kev_planes.size() ? p.kev_planes[i / 2] : 0;
      if (has_edges && pl == 1) { esum += k; ne++; }
      else { bsum += k; nb++; planes += (uint64_t)pl; }
    }
    if (bsum >= bulk) { bulk = bsum; n_bulk = nb; bulk_planes = planes; }
    if (esum >= edge) { edge = esum; n_edge = ne; }
  }
  s->last_total_ms = total;
  s->last_kernel_ms = bulk;
  s->last_kernel_launches = n_bulk;
  s->last_edge_ms = edge;
  s->last_edge_launches = n_edge;
  s->last_bulk_planes = bulk_planes;
  if (s->ev_h0) {
    float h = 0;
    cudaSetDevice(s->parts[0].device);
    if (cudaEventElapsedTime(&h, s->ev_h0, s->ev_h1) == cudaSuccess) s->last_halo_ms = h;
    else cudaGetLastError();
  }
  return PFDTD_OK;
}

int pfdtd_last_timing(pfdtd_solver* s, float* total_ms, float* update_kernel_ms, uint32_t* n_update_launches) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  if (total_ms) *total_ms = s->last_total_ms;
  if (update_kernel_ms) *update_kernel_ms = s->last_kernel_ms;
  if (n_update_launches) *n_update_launches = s->last_kernel_launches;
  return PFDTD_OK;
}

int pfdtd_last_timing_detail(pfdtd_solver* s, float* bulk_kernel_ms, uint32_t* n_bulk_launches, uint64_t* bulk_planes, float* edge_kernel_ms,
                             uint32_t* n_edge_launches) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  if (bulk_kernel_ms) *bulk_kernel_ms = s->last_kernel_ms;
  if (n_bulk_launches) *n_bulk_launches = s->last_kernel_launches;
  if (bulk_planes) *bulk_planes = s->last_bulk_planes;
  if (edge_kernel_ms) *edge_kernel_ms = s->last_edge_ms;
  if (n_edge_launches) *n_edge_launches = s->last_edge_launches;
  return PFDTD_OK;
}

// The halo exchange alone (CudaMesh::switchHalos, cudaMesh.h:432-463), `reps` times back to back on the edge stream(s)
// with nothing else running, timed with CUDA events on that stream: what one exchange of two planes per interface
// costs on the link when both sides are ready.  Collective across processes (every rank must call it with the same
// `reps`, after a barrier).  The fields are exchanged onto themselves: results do not change.
int pfdtd_time_halo_exchange(pfdtd_solver* s, uint32_t reps, float* ms_per_exchange) {
  PF_CHECK(s && ms_per_exchange && !s->parts.empty(), PFDTD_ERR_INVALID, "bad argument");
  PF_CHECK(reps >= 1, PFDTD_ERR_INVALID, "reps must be >= 1");
  *ms_per_exchange = 0.f;
  if (s->parts.size() == 1 && !s->comm) return PFDTD_OK;
  PF_TRY(sync_all(s));
  Partition& p0 = s->parts[0];
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  PF_CUDA(cudaSetDevice(p0.device));
  PF_CUDA(cudaEventCreate(&e0));
  PF_CUDA(cudaEventCreate(&e1));
  int rc = PFDTD_OK;
  for (uint32_t r = 0; r <= reps && rc == PFDTD_OK; r++) {      // exchange 0 is a warm-up (connection set-up)
    if (r == 1) { cudaSetDevice(p0.device); cudaEventRecord(e0, p0.s_edge); }
    for (size_t k = 0; k + 1 < s->parts.size() && rc == PFDTD_OK; k++) {
      rc = enqueue_halo_local_up(s, k, s->cur);
      if (rc == PFDTD_OK) rc = enqueue_halo_local_down(s, k, s->cur);
    }
    if (rc == PFDTD_OK) rc = enqueue_halo_external(s, s->cur);
    if (r == 0 && rc == PFDTD_OK) rc = sync_all(s);
  }
  if (rc == PFDTD_OK) { cudaSetDevice(p0.device); cudaEventRecord(e1, p0.s_edge); rc = sync_all(s); }
  float ms = 0;
  if (rc == PFDTD_OK && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) *ms_per_exchange = ms / (float)reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int pfdtd_last_halo_ms(pfdtd_solver* s, float* halo_ms) {
  PF_CHECK(s && halo_ms, PFDTD_ERR_INVALID, "null argument");
  *halo_ms = s->last_halo_ms;
  return PFDTD_OK;
}

int pfdtd_fetch_responses(pfdtd_solver* s, void* h_response, uint32_t n_steps) {
  PF_CHECK(s && (h_response || s->n_rec == 0), PFDTD_ERR_INVALID, "null response buffer");
  PF_CHECK(n_steps <= s->rec_cap || s->n_rec == 0, PFDTD_ERR_RANGE, "asked for %u steps but only %u were reserved", n_steps,
           s->rec_cap);
  PF_TRY(sync_all(s));
  const size_t es = esize(s);
  if (s->n_rec) memset(h_response, 0, (size_t)s->n_rec * n_steps * es);
  for (auto& p : s->parts) {
    if (!p.d_rec_out || p.rec_slots.empty()) continue;
    PF_CUDA(cudaSetDevice(p.device));
    const uint32_t have = std::min(n_steps, p.rec_cap_alloc);   // steps reserved after the last enqueue hold nothing yet
    for (int32_t slot : p.rec_slots)
      PF_CUDA(cudaMemcpy((char*)h_response + (size_t)slot * n_steps * es, (char*)p.d_rec_out + (size_t)slot * p.rec_cap_alloc * es,
                         (size_t)have * es, cudaMemcpyDeviceToHost));
  }
  return PFDTD_OK;
}

int pfdtd_run(pfdtd_solver* s, uint32_t n_steps, void* h_response, pfdtd_interrupt_cb interrupt, pfdtd_progress_cb progress,
              float* seconds_per_step) {
  PF_CHECK(s && !s->parts.empty(), PFDTD_ERR_INVALID, "make_partition must be called before run");
  auto t0 = std::chrono::steady_clock::now();
  PF_TRY(pfdtd_reserve_steps(s, n_steps));
  // Steps are enqueued in blocks that end on multiples of PROGRESS_MOD (kernels3d.h:36), so the progress callback keeps
  // its cadence without a device synchronisation per step.  The reference polls the interrupt callback every step
  // (kernels3d.cu:86-89); here, with a callback installed, a block is cut so that it holds about 25 ms of device work
  // (measured on the previous block) and the queue is drained before the next poll: cancelling takes tens of
  // milliseconds whatever the mesh size, at the price of one synchronisation per block.
  const uint32_t cadence = 100;
  uint32_t done = 0;
  bool interrupted = false;
  auto tb = t0;
  double ms_per_step = 0.0;
  while (done < n_steps) {
    if (interrupt && interrupt()) { interrupted = true; break; }
    uint32_t nb = std::min(cadence - done % cadence, n_steps - done);
    if (interrupt && ms_per_step > 0.0) nb = std::min<uint32_t>(nb, (uint32_t)std::max(1.0, 25.0 / ms_per_step));
    else if (interrupt && done == 0) nb = std::min<uint32_t>(nb, 8);        // a first short block to learn the step time
    PF_TRY(pfdtd_enqueue_steps(s, done, nb));
    if (interrupt) {
      PF_TRY(pfdtd_sync(s));
      ms_per_step = s->last_total_ms / (double)nb;
    }
    done += nb;
    if (progress && (done % cadence == 0 || done == n_steps)) {
      PF_TRY(sync_all(s));
      auto tn = std::chrono::steady_clock::now();
      const uint32_t since = done % cadence == 0 ? cadence : done % cadence;
      progress((int)(done - since), (int)n_steps, (float)(std::chrono::duration<double>(tn - tb).count() / since));
      tb = tn;
    }
  }
  PF_TRY(pfdtd_sync(s));
  if (h_response && s->n_rec) PF_TRY(pfdtd_fetch_responses(s, h_response, n_steps));
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (seconds_per_step) *seconds_per_step = (float)(secs / std::max(1u, done));
  if (interrupted) { set_error("interrupted at step %u", done); return PFDTD_ERR_INTERRUPTED; }
  return PFDTD_OK;
}

int pfdtd_step(pfdtd_solver* s, uint32_t step, int direction, void* h_response, uint32_t n_steps_total) {
  PF_CHECK(s && !s->parts.empty(), PFDTD_ERR_INVALID, "make_partition must be called before step");
  PF_CHECK(direction == 1 || direction == -1, PFDTD_ERR_INVALID, "direction must be +1 or -1");
  // the reference writes h_return_ptr[rec * numSteps + step] unchecked (kernels3d.cu:447-455)
  PF_CHECK(!h_response || s->n_rec == 0 || step < n_steps_total, PFDTD_ERR_RANGE, "step %u beyond the %u-step response buffer", step,
           n_steps_total);
  PF_TRY(prepare_srcrec(s));
  PF_TRY(ensure_class_tables(s));
  PF_TRY(sync_all(s));
  PF_TRY(set_step_counters(s, (int)step, 0x7fffffff, (int)step));
  const bool single = (s->parts.size() == 1 && !s->comm);
  for (auto& p : s->parts) {
    PF_CUDA(cudaSetDevice(p.device));
    cudaStream_t st = p.s_main;
    PF_TRY(launch_srcrec_for(s, p, st, 0, 1, 0));
    PF_TRY(launch_update(s, p, 1, (int)p.size - 1, p.cfg_full, st, false));
  }
  PF_TRY(sync_all(s));
  // time-reversal rule of launchFDTD3dStep (kernels3d.cu:462-465)
  if (s->past_direction == direction) s->cur = 1 - s->cur;
  s->past_direction = direction;
  PF_TRY(pfdtd_switch_halos(s));
  if (h_response) {
    for (uint32_t r = 0; r < s->n_rec; r++) {
      double v = 0;
      int64_t z = (int64_t)s->rec_xyz[3 * r + 2] - s->opt_global_z_first;
      if (z >= 0) PF_TRY(pfdtd_get_sample(s, (uint32_t)s->rec_xyz[3 * r], (uint32_t)s->rec_xyz[3 * r + 1], (uint32_t)z, &v));
      if (s->dtype == PFDTD_F32) ((float*)h_response)[(size_t)r * n_steps_total + step] = (float)v;
      else ((double*)h_response)[(size_t)r * n_steps_total + step] = v;
    }
  }
  return PFDTD_OK;
}

// ---- multi-process ------------------------------------------------------------------------------------------
int pfdtd_comm_unique_id(uint8_t* out_id128) {
  PF_CHECK(out_id128, PFDTD_ERR_INVALID, "null argument");
  PF_TRY(nccl_load());
  NcclId id;
  memset(&id, 0, sizeof(id));
  PF_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out_id128, id.bytes, 128);
  return PFDTD_OK;
}

int pfdtd_comm_init(pfdtd_solver* s, const uint8_t* id128, int rank, int nranks) {
  PF_CHECK(s && id128, PFDTD_ERR_INVALID, "null argument");
  PF_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, PFDTD_ERR_INVALID, "bad rank %d of %d", rank, nranks);
  PF_CHECK(!s->parts.empty(), PFDTD_ERR_INVALID, "make_partition must be called before comm_init");
  // one communicator on one device: the process owns exactly one slab (one process per GPU)
  PF_CHECK(s->parts.size() == 1, PFDTD_ERR_INVALID, "pfdtd_comm_init needs exactly one local partition (this solver has %zu)",
           s->parts.size());
  PF_TRY(nccl_load());
  NcclId id;
  memcpy(id.bytes, id128, 128);
  PF_CUDA(cudaSetDevice(s->parts[0].device));
  // communicators are process-lifetime objects keyed by (id, rank, nranks): a second solver created with the
  // same id (e.g. the next simulation of the same job) reuses the communicator instead of paying
  // ncclCommInitRank again
  static std::map<std::string, void*> cache;
  static std::mutex cache_mu;
  std::string key((const char*)id128, 128);
  key += "/" + std::to_string(rank) + "/" + std::to_string(nranks) + "/" + std::to_string(s->parts[0].device);
  void* comm = nullptr;
  {
    std::lock_guard<std::mutex> g(cache_mu);
    auto it = cache.find(key);
    if (it != cache.end()) comm = it->second;
  }
  if (!comm) {
    PF_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    std::lock_guard<std::mutex> g(cache_mu);
    cache[key] = comm;
  }
  s->comm = comm;
  s->rank = rank;
  s->nranks = nranks;
  s->srcrec_dirty = true;
  if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
  return exchange_links(s);
}

int pfdtd_comm_release(pfdtd_solver* s) {
  PF_CHECK(s, PFDTD_ERR_INVALID, "null solver");
  if (!s->parts.empty()) PF_TRY(sync_all(s));
  for (auto& l : s->link) {
    for (void*& q : l.nb_P) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    if (l.nb_flags) { cudaIpcCloseMemHandle(l.nb_flags); l.nb_flags = nullptr; }
    l.ok = false;
  }
  return PFDTD_OK;
}

int pfdtd_halo_transport(pfdtd_solver* s, char* buf, size_t buflen) {
  PF_CHECK(s && buf && buflen > 0, PFDTD_ERR_INVALID, "bad argument");
  if (s->parts.size() > 1) {
    bool all = s->opt_peer_stores != 0;
    for (size_t k = 0; all && k + 1 < s->parts.size(); k++) all = can_store_to(s, k, k + 1) && can_store_to(s, k + 1, k);
    snprintf(buf, buflen, "%s", all ? "in-process: peer-mapped stores from the edge launches" : "in-process: cudaMemcpyPeerAsync");
  } else if (!s->comm || s->nranks == 1) {
    snprintf(buf, buflen, "none");
  } else {
    const bool lo = s->rank == 0 || ipc_side(s, 0), hi = s->rank == s->nranks - 1 || ipc_side(s, 1);
    snprintf(buf, buflen, "%s", (lo && hi) ? "peer-mapped stores from the edge launches (CUDA IPC over NVLink), flag hand-over"
                                           : ((ipc_side(s, 0) || ipc_side(s, 1)) ? "mixed: peer-mapped stores / NCCL send-recv" : "NCCL send-recv"));
  }
  return PFDTD_OK;
}

// ---- introspection --------------------------------------------------------------------------------------------
int pfdtd_kernel_name(pfdtd_solver* s, char* buf, size_t buflen) {
  PF_CHECK(s && buf && buflen > 0, PFDTD_ERR_INVALID, "bad argument");
  if (s->parts.empty()) { snprintf(buf, buflen, "none"); return PFDTD_OK; }
  const Partition& p = s->parts[0];
  const char* sch = s->scheme == SCH_CENTRED ? "centred" : s->scheme == SCH_INTERP ? (s->dcoef[2] != 0 ? "interp27+corners" : "interp27") : "forward";
  const char* fam = s->scheme == SCH_INTERP ? "fdtd_update_interp" : "fdtd_update";
  if (p.use_tma)
    snprintf(buf, buflen, "%s_tma<%s,%s> tile %s chunk %d, %d CTAs/SM%s", fam, s->dtype == PFDTD_F32 ? "f32" : "f64", sch,
             tma_tile_name(s->dtype, p.cfg_full.tile), p.cfg_full.chunk, p.cfg_full.occupancy,
             s->wide ? ", position classes x material table" : "");
  else
    snprintf(buf, buflen, "%s_plain<%s,%s>", fam, s->dtype == PFDTD_F32 ? "f32" : "f64", sch);
  return PFDTD_OK;
}

int pfdtd_launch_count(pfdtd_solver* s, uint64_t* n) {
  PF_CHECK(s && n, PFDTD_ERR_INVALID, "null argument");
  *n = s->launch_count;
  return PFDTD_OK;
}

}  // extern "C"
