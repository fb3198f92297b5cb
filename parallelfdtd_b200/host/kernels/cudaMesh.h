// cudaMesh.h -- `class CudaMesh` with the reference's public interface (reference src/kernels/cudaMesh.h:62-840),
// implemented as a thin owner of a `pfdtd_solver*` (C ABI, include/pfdtd.h).  What the reference does
// with per-call cudaMemcpy / cudaMemcpyPeer / cudaDeviceSynchronize happens inside libpfdtd_b200.so on
// streams; this class only keeps the call shapes so App, the MEX gateway and the reference's tests
// compile against it unchanged.
#pragma once
#include <cstdlib>
#include <vector>
#include "cudaUtils.h"

// masks of the node bytes (reference cudaMesh.h:35-44)
#define INSIDE_SWITCH 7
#define FORWARD_POSITION_MASK 0X7F
#define CENTERED_MASK 0X80
#define DIR_X 0X01
#define DIR_Y 0X02
#define DIR_Z 0X04
#define SIGN_X 0X10
#define SIGN_Y 0X20
#define SIGN_Z 0X40

class CudaMesh {
 public:
  CudaMesh() : solver_(0), double_(false), block_(make_uint3(32, 4, 1)) { pfdtd_safe(pfdtd_create(&solver_), "CudaMesh"); }
  ~CudaMesh() { if (solver_) pfdtd_destroy(solver_); }
  CudaMesh(const CudaMesh&) = delete;
  CudaMesh& operator=(const CudaMesh&) = delete;

  pfdtd_solver* handle() { return solver_; }
  void setOption(int option, long long value) { pfdtd_safe(pfdtd_set_option(solver_, option, value), "CudaMesh::setOption"); }

  // ---- setup (reference cudaMesh.cu:26-153).  d_* are DEVICE volumes from the voxelizer and are adopted.
  void setupMesh(unsigned char* d_position_ptr, unsigned char* d_material_ptr, unsigned int number_of_unique_materials,
                 float* material_coefficients, float* parameter_ptr, uint3 voxelization_dim, uint3 block_size,
                 unsigned int element_type) {
    double_ = false; block_ = block_size;
    pfdtd_safe(pfdtd_setup_mesh_device(solver_, -1, d_position_ptr, d_material_ptr, voxelization_dim.x, voxelization_dim.y,
                                       voxelization_dim.z, block_size.x, block_size.y, block_size.z, element_type, PFDTD_F32,
                                       parameter_ptr, material_coefficients, number_of_unique_materials), "CudaMesh::setupMesh");
  }
  void setupMeshDouble(unsigned char* d_position_ptr, unsigned char* d_material_ptr, unsigned int number_of_unique_materials,
                       double* material_coefficients, double* parameter_ptr, uint3 voxelization_dim, uint3 block_size,
                       unsigned int element_type) {
    double_ = true; block_ = block_size;
    pfdtd_safe(pfdtd_setup_mesh_device(solver_, -1, d_position_ptr, d_material_ptr, voxelization_dim.x, voxelization_dim.y,
                                       voxelization_dim.z, block_size.x, block_size.y, block_size.z, element_type, PFDTD_F64,
                                       parameter_ptr, material_coefficients, number_of_unique_materials), "CudaMesh::setupMeshDouble");
  }
  // same from HOST volumes (synthetic geometry, tests; the reference's tests upload with toDevice first)
  void setupMeshHost(const unsigned char* h_position, const unsigned char* h_material, unsigned int number_of_unique_materials,
                     const void* material_coefficients, const void* parameter_ptr, uint3 dim, uint3 block_size,
                     unsigned int element_type) {
    block_ = block_size;
    pfdtd_safe(pfdtd_setup_mesh(solver_, h_position, h_material, dim.x, dim.y, dim.z, block_size.x, block_size.y, block_size.z,
                                element_type, double_ ? PFDTD_F64 : PFDTD_F32, parameter_ptr, material_coefficients,
                                number_of_unique_materials), "CudaMesh::setupMeshHost");
  }
  // reference cudaMesh.h:648-751
  void makePartition(unsigned int number_of_partitions, std::vector<unsigned int> device_list = std::vector<unsigned int>()) {
    bound_params_ = 0;
    pfdtd_safe(pfdtd_make_partition(solver_, number_of_partitions, device_list.empty() ? 0 : &device_list[0]), "CudaMesh::makePartition");
  }
  void destroyPartitions() { pfdtd_destroy(solver_); solver_ = 0; bound_params_ = 0; pfdtd_safe(pfdtd_create(&solver_), "CudaMesh::destroyPartitions"); }
  // Which (SimulationParameters, generation) the source / receiver tables inside the solver were built from
  // (launchFDTD3dStep rebuilds them when either changes; the solver drops them on makePartition).
  bool boundTo(const void* params, unsigned long long generation) const { return bound_params_ == params && bound_generation_ == generation; }
  void setBound(const void* params, unsigned long long generation) { bound_params_ = params; bound_generation_ = generation; }

  // ---- getters (reference cudaMesh.h:184-244)
  unsigned int getNumberOfPartitions() { unsigned int n = 0; pfdtd_get_num_partitions(solver_, &n); return n; }
  unsigned int getPartitionSize() { return getPartitionSize(0); }
  unsigned int getPartitionSize(int partition) { unsigned int f, n, d; part(partition, &f, &n, &d); return n; }
  unsigned int getFirstSliceIdx(int partition) { unsigned int f, n, d; part(partition, &f, &n, &d); return f; }
  unsigned int getDeviceAt(int i) { unsigned int f, n, d; part(i, &f, &n, &d); return d; }
  unsigned int getNumberOfElementsAt(unsigned int partition) { return getPartitionSize((int)partition) * getDimXY(); }
  unsigned int getBlockX() { return block_.x; }
  unsigned int getBlockY() { return block_.y; }
  unsigned int getBlockZ() { return block_.z; }
  unsigned int getDimX() { unsigned int x, y, z; pfdtd_get_dims(solver_, &x, &y, &z); return x; }
  unsigned int getDimY() { unsigned int x, y, z; pfdtd_get_dims(solver_, &x, &y, &z); return y; }
  unsigned int getDimZ() { unsigned int x, y, z; pfdtd_get_dims(solver_, &x, &y, &z); return z; }
  unsigned int getDimXY() { unsigned int x, y, z; pfdtd_get_dims(solver_, &x, &y, &z); return x * y; }
  unsigned int getGridDimX() { return getDimX() / block_.x; }
  unsigned int getGridDimY() { return getDimY() / block_.y; }
  unsigned int getGridDimZ() { return getDimZ() / block_.z; }
  unsigned int getNumberOfElements() { unsigned long long n = 0; counts(&n, 0, 0); return (unsigned int)n; }
  unsigned long long getNumberOfElements64() { unsigned long long n = 0; counts(&n, 0, 0); return n; }
  unsigned int getNumberOfAirElements() { unsigned long long n = 0; counts(0, &n, 0); return (unsigned int)n; }
  unsigned int getNumberOfBoundaryElements() { unsigned long long n = 0; counts(0, 0, &n); return (unsigned int)n; }
  bool isDouble() const { return double_; }
  void setDouble(bool is_double) { double_ = is_double; }

  // device pointers for capture / visualisation kernels (reference cudaMesh.h:184-212)
  float* getPressurePtrAt(unsigned int k) { void* p = 0; ptrs(k, &p, 0, 0, 0); return (float*)p; }
  float* getPastPressurePtrAt(unsigned int k) { void* p = 0; ptrs(k, 0, &p, 0, 0); return (float*)p; }
  double* getPressureDoublePtrAt(unsigned int k) { void* p = 0; ptrs(k, &p, 0, 0, 0); return (double*)p; }
  double* getPastPressureDoublePtrAt(unsigned int k) { void* p = 0; ptrs(k, 0, &p, 0, 0); return (double*)p; }
  unsigned char* getPositionIdxPtrAt(unsigned int k) { unsigned char* p = 0; ptrs(k, 0, 0, &p, 0); return p; }
  unsigned char* getMaterialIdxPtrAt(unsigned int k) { unsigned char* p = 0; ptrs(k, 0, 0, 0, &p); return p; }

  // ---- indexing (reference cudaMesh.h:247-307)
  unsigned int getElementIndex(unsigned int x, unsigned int y, unsigned int z) { return z * getDimXY() + y * getDimX() + x; }
  void getElementIdxAndDevice(unsigned int x, unsigned int y, unsigned int z, int* dev_i, int* elem_idx) {
    int64_t e = -1; int p = -1;
    pfdtd_safe(pfdtd_get_element_idx_and_partition(solver_, x, y, z, &p, &e), "CudaMesh::getElementIdxAndDevice");
    *dev_i = p; *elem_idx = (int)e;
  }
  std::vector<std::vector<unsigned int> > getPartitionIndexing(int num_parts, int dim) {
    std::vector<unsigned int> first(num_parts), size(num_parts);
    pfdtd_safe(pfdtd_partition_indexing((unsigned int)dim, (unsigned int)num_parts, &first[0], &size[0]), "CudaMesh::getPartitionIndexing");
    std::vector<std::vector<unsigned int> > ret(num_parts);
    for (int i = 0; i < num_parts; i++) { ret[i].resize(size[i]); for (unsigned int j = 0; j < size[i]; j++) ret[i][j] = first[i] + j; }
    return ret;
  }

  // ---- single samples (reference cudaMesh.h:497-584)
  template <typename T> void setSample(T sample, unsigned int x, unsigned int y, unsigned int z) {
    pfdtd_safe(pfdtd_set_sample(solver_, x, y, z, (double)sample), "CudaMesh::setSample"); }
  template <typename T> void addSample(T sample, unsigned int x, unsigned int y, unsigned int z) {
    pfdtd_safe(pfdtd_add_sample(solver_, x, y, z, (double)sample), "CudaMesh::addSample"); }
  template <typename T> void setSampleAt(T sample, unsigned int x, unsigned int y, unsigned int z, unsigned int partition) {
    pfdtd_safe(pfdtd_set_sample_at(solver_, x, y, z, partition, (double)sample), "CudaMesh::setSampleAt"); }
  template <typename T> T getSample(unsigned int x, unsigned int y, unsigned int z) {
    double v = 0; pfdtd_safe(pfdtd_get_sample(solver_, x, y, z, &v), "CudaMesh::getSample"); return (T)v; }
  template <typename T> T getSampleAt(unsigned int x, unsigned int y, unsigned int z, unsigned int partition) {
    double v = 0; pfdtd_safe(pfdtd_get_sample_at(solver_, x, y, z, partition, &v), "CudaMesh::getSampleAt"); return (T)v; }

  // position byte of a voxel, from the first partition containing z (reference cudaMesh.h:591-596)
  unsigned char getPositionSample(unsigned int x, unsigned int y, unsigned int z) {
    int part = -1, el = -1; getElementIdxAndDevice(x, y, z, &part, &el);
    if (part < 0) throw std::out_of_range("CudaMesh::getPositionSample: z outside every partition");
    unsigned char v = 0;
    pfdtd_safe(pfdtd_device_download((int)getDeviceAt(part), &v, getPositionIdxPtrAt((unsigned int)part) + el, 1), "CudaMesh::getPositionSample");
    return v;
  }
  // device of the LAST partition holding slice z (reference cudaMesh.h:268-278), -1 if none
  int getDeviceOfElement(unsigned int, unsigned int, unsigned int z) {
    int ret = -1;
    const unsigned int n = getNumberOfPartitions();
    for (unsigned int i = 0; i < n; i++) {
      unsigned int f, sz, d; part(i, &f, &sz, &d);
      if (z > f + sz - 1) continue;
      if (z < f) break;
      ret = (int)d;
    }
    return ret;
  }

  // ---- slices (reference cudaMesh.h:600-646 getSlice / getPositionSlice; malloc'ed, caller frees, orientation 0
  // only as in the reference -- the other orientations go through captureSlice)
  template <typename T> T* getSlice(unsigned int slice, unsigned int orientation) {
    if (orientation != 0 || (sizeof(T) == 8) != double_) return (T*)0;
    T* data = (T*)std::malloc((size_t)getDimXY() * sizeof(T));
    pfdtd_safe(pfdtd_capture_slice(solver_, slice, 0, data, 0), "CudaMesh::getSlice");
    return data;
  }
  unsigned char* getPositionSlice(unsigned int slice, unsigned int orientation) {
    if (orientation != 0) return (unsigned char*)0;
    const size_t n = getDimXY();
    std::vector<unsigned char> scratch(n * (double_ ? 8 : 4));
    unsigned char* data = (unsigned char*)std::malloc(n);
    pfdtd_safe(pfdtd_capture_slice(solver_, slice, 0, &scratch[0], data), "CudaMesh::getPositionSlice");
    return data;
  }
  // captureSliceFast / captureMesh of visualizationUtils.cu:111-254 as members: device-side gather, slice-sized D2H
  template <typename T> void captureSlice(unsigned int slice, unsigned int orientation, T* pressure, unsigned char* position) {
    pfdtd_safe(pfdtd_capture_slice(solver_, slice, orientation, pressure, position), "CudaMesh::captureSlice"); }
  template <typename T> void captureMesh(T* field) { pfdtd_safe(pfdtd_capture_mesh(solver_, field), "CudaMesh::captureMesh"); }

  // ---- step pieces (reference cudaMesh.h:755-791)
  void switchHalos() { pfdtd_safe(pfdtd_switch_halos(solver_), "CudaMesh::switchHalos"); }
  void flipPressurePointers() { pfdtd_safe(pfdtd_flip_pressure_pointers(solver_), "CudaMesh::flipPressurePointers"); }
  void resetPressures() { pfdtd_safe(pfdtd_reset_pressures(solver_), "CudaMesh::resetPressures"); }

 private:
  void part(int k, unsigned int* f, unsigned int* n, unsigned int* d) { pfdtd_safe(pfdtd_get_partition(solver_, (unsigned int)k, f, n, d), "CudaMesh::partition"); }
  void counts(unsigned long long* n, unsigned long long* a, unsigned long long* b) {
    uint64_t nn = 0, aa = 0, bb = 0; pfdtd_get_counts(solver_, &nn, &aa, &bb);
    if (n) *n = nn;
    if (a) *a = aa;
    if (b) *b = bb;
  }
  void ptrs(unsigned int k, void** p, void** pp, unsigned char** pos, unsigned char** mat) {
    pfdtd_safe(pfdtd_get_device_pointers(solver_, k, p, pp, pos, mat), "CudaMesh::devicePointers"); }
  pfdtd_solver* solver_;
  const void* bound_params_ = 0;
  unsigned long long bound_generation_ = 0;
  bool double_;
  uint3 block_;
};
