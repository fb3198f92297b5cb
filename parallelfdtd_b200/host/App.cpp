// App.cpp -- see App.h.  Behaviour follows reference src/App.cpp where the hot path is concerned:
// initializeMesh (:147-234) -> setupMesh + makePartition, runSimulation (:308-348) ->
// launchFDTD3d[Double], executeStep (:407-436) -> launchFDTD3dStep + captures.
#include "App.h"

#include <chrono>
#include <cmath>
#include <cstdio>

#include "kernels/kernels3d.h"
#include "voxelize.h"

namespace FDTD {

static bool defaultInterrupt(void) { return false; }
static void defaultProgress(int step, int max_step, float t_per_step) {
  std::printf("Step %d/%d, time per step %f\n", step, max_step, t_per_step);
}

App::App()
    : m_interrupt(0), m_progress(0), current_step_(0), step_direction_(1), number_of_devices_(0), best_device_(0),
      force_partition_to_(-1), capture_db_(60), time_per_step_(0.f), num_elements_(0) {
  vol_dim_[0] = vol_dim_[1] = vol_dim_[2] = 0;
  loggerInit();
  setupDefaultCallbacks();
}
App::~App() {}

void App::setupDefaultCallbacks() { m_interrupt = defaultInterrupt; m_progress = defaultProgress; }

void App::queryDevices() {
  int n = 0;
  pfdtd_device_count(&n);
  number_of_devices_ = n;
  device_mem_sizes_.clear();
  int best = 0, best_free = -1;
  for (int i = 0; i < n; i++) {
    int total = 0, free_mb = 0;
    pfdtd_safe(pfdtd_device_mem_mb(i, &total, &free_mb), "App::queryDevices");
    device_mem_sizes_.push_back(free_mb);
    if (free_mb > best_free) { best_free = free_mb; best = i; }
    log_msg<LOG_INFO>(L"App::queryDevices - memory size dev %d: %d MB") % i % free_mb;
  }
  best_device_ = best;
}
// the reference calls cudaDeviceReset on every device here; resetting would also tear down other users of the
// process' CUDA context (torch, MATLAB's GPU arrays), so this build only releases its own partitions
void App::resetDevices() { m_mesh.destroyPartitions(); }
void App::initializeDevices() {
  queryDevices();
  if (number_of_devices_ < 1) { c_log_msg(LOG_ERROR, "App::initializeDevices - no CUDA device"); throw(-1); }
}

void App::initializeGeometry(unsigned int* indices, float* vertices, unsigned int number_of_indices, unsigned int number_of_vertices) {
  m_geometry.initialize(indices, vertices, number_of_indices, number_of_vertices);
  nv::Vec3f bb = m_geometry.getBoundingBox();
  log_msg<LOG_INFO>(L"App::initializeGeometry - %d triangles, bounding box %f x %f x %f") % m_geometry.getNumberOfTriangles() % bb.x % bb.y % bb.z;
}

void App::setVoxelVolumes(const unsigned char* bid, const unsigned char* mat, unsigned int vx, unsigned int vy, unsigned int vz) {
  const size_t n = (size_t)vx * vy * vz;
  vol_bid_.assign(bid, bid + n);
  vol_mat_.assign(mat, mat + n);
  vol_dim_[0] = vx; vol_dim_[1] = vy; vol_dim_[2] = vz;
}

void App::initializeMesh(unsigned int number_of_partitions) {
  if (number_of_devices_ == 0) queryDevices();
  if (vol_bid_.empty()) {
    if (m_geometry.getNumberOfTriangles() == 0) { c_log_msg(LOG_ERROR, "App::initializeMesh - no geometry"); throw(-1); }
    pfdtd_host::VoxelVolumes v = pfdtd_host::voxelize(m_geometry, m_parameters.getDx(), m_materials.getMaterialIdxPtr());
    vol_bid_.swap(v.bid); vol_mat_.swap(v.mat);
    vol_dim_[0] = v.vx; vol_dim_[1] = v.vy; vol_dim_[2] = v.vz;
  }
  if (m_materials.getNumberOfUniqueMaterials() == 0) m_materials.setGlobalMaterial(uniform_surfaces_(), 0.f);
  const uint3 dim = make_uint3(vol_dim_[0], vol_dim_[1], vol_dim_[2]);
  const uint3 block = make_uint3(32, 4, 1);                     // reference App.cpp:193
  const unsigned int type = (unsigned int)m_parameters.getUpdateType();
  if (m_mesh.isDouble())
    m_mesh.setupMeshHost(&vol_bid_[0], &vol_mat_[0], m_materials.getNumberOfUniqueMaterials(), m_materials.getMaterialCoefficientPtrDouble(),
                         m_parameters.getParameterPtrDouble(), dim, block, type);
  else
    m_mesh.setupMeshHost(&vol_bid_[0], &vol_mat_[0], m_materials.getNumberOfUniqueMaterials(), m_materials.getMaterialCoefficientPtr(),
                         m_parameters.getParameterPtr(), dim, block, type);
  num_elements_ = m_mesh.getNumberOfElements();
  // Partition count.  The reference splits in two above 90e6 (45e6 double) voxels because of Kepler-era memory
  // (App.cpp:217-233); here one partition is used whenever the mesh fits the device, otherwise as many slabs
  // as needed (bounded by the device count).  forcePartitionTo keeps its meaning.
  unsigned int n = 1;
  const double bytes = (double)m_mesh.getNumberOfElements64() * (2.0 * (m_mesh.isDouble() ? 8 : 4) + 3.0);
  if (force_partition_to_ != -1 && force_partition_to_ <= number_of_devices_) n = (unsigned int)force_partition_to_;
  else {
    const double cap = device_mem_sizes_.empty() ? 150e9 : 0.9 * 1e6 * (double)device_mem_sizes_[0];
    while (n < (unsigned int)number_of_devices_ && n < number_of_partitions * 4 && bytes / n > cap) n++;
  }
  m_mesh.makePartition(n);
  current_step_ = 0;
}

void App::runSimulation() {
  const auto t0 = std::chrono::steady_clock::now();
  initializeMesh(2);
  const size_t nresp = (size_t)m_parameters.getNumSteps() * m_parameters.getNumReceivers();
  if (m_mesh.isDouble()) {
    responses_double_.assign(nresp ? nresp : 1, 0.0);
    time_per_step_ = launchFDTD3dDouble(&m_mesh, &m_parameters, &responses_double_[0], m_interrupt, m_progress);
    responses_double_.resize(nresp);
  } else {
    responses_.assign(nresp ? nresp : 1, 0.f);
    time_per_step_ = launchFDTD3d(&m_mesh, &m_parameters, &responses_[0], m_interrupt, m_progress);
    responses_.resize(nresp);
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  log_msg<LOG_INFO>(L"App::runSimulation - time: %f seconds, Mvox/sec: %f") % secs % getMvoxPerSec();
}

void App::runCapture() {
  const auto t0 = std::chrono::steady_clock::now();
  m_mesh.setDouble(false);                                        // reference App.cpp:354
  initializeMesh(2);
  const unsigned int steps = m_parameters.getNumSteps();
  responses_.assign((size_t)steps * m_parameters.getNumReceivers() + 1, 0.f);
  for (unsigned int i = 0; i < steps; i++) {
    executeStep();
    if (m_interrupt && m_interrupt()) break;
  }
  responses_.resize((size_t)steps * m_parameters.getNumReceivers());
  time_per_step_ = (float)(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / (steps ? steps : 1));
}

void App::executeStep() {
  const auto t0 = std::chrono::steady_clock::now();
  if (responses_.empty()) responses_.assign((size_t)m_parameters.getNumSteps() * m_parameters.getNumReceivers() + 1, 0.f);
  launchFDTD3dStep(&m_mesh, &m_parameters, &responses_[0], current_step_, step_direction_, m_progress);
  current_step_ += step_direction_;
  captureIfDue_();
  const float dt = (float)std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  time_per_step_ = (time_per_step_ + dt) / 2.f;                     // running mean as in the reference (:433-434)
}

// slice / mesh captures at the listed steps (reference visualizationUtils.cu:111-254 captureSliceFast /
// captureMesh): the current pressure field is gathered from all slabs (halo planes dropped)
void App::captureIfDue_() {
  bool due = false;
  for (size_t i = 0; i < step_to_capture_.size(); i++) due |= step_to_capture_[i] == current_step_;
  for (size_t i = 0; i < mesh_to_capture_.size(); i++) due |= mesh_to_capture_[i] == current_step_;
  if (!due) return;
  const unsigned int X = m_mesh.getDimX(), Y = m_mesh.getDimY(), Z = m_mesh.getDimZ();
  std::vector<float> field((size_t)X * Y * Z, 0.f);
  const unsigned int np = m_mesh.getNumberOfPartitions();
  for (unsigned int k = 0; k < np; k++) {
    const unsigned int first = m_mesh.getFirstSliceIdx((int)k), nz = m_mesh.getPartitionSize((int)k);
    std::vector<float> slab((size_t)nz * X * Y);
    pfdtd_safe(pfdtd_export_partition_pressure(m_mesh.handle(), k, 0, &slab[0]), "App::capture");
    const unsigned int lo = k == 0 ? 0 : 1, hi = k + 1 == np ? nz : nz - 1;   // own planes; halos belong to the neighbours
    for (unsigned int z = lo; z < hi; z++)
      std::copy(slab.begin() + (size_t)z * X * Y, slab.begin() + (size_t)(z + 1) * X * Y, field.begin() + (size_t)(first + z) * X * Y);
  }
  for (size_t i = 0; i < step_to_capture_.size(); i++) {
    if (step_to_capture_[i] != current_step_) continue;
    const unsigned int s = slice_to_capture_[i], o = slice_orientation_[i];
    std::vector<float> img;
    if (o == 0) { img.assign(field.begin() + (size_t)s * X * Y, field.begin() + (size_t)(s + 1) * X * Y); }                  // xy at z = s
    else if (o == 1) { img.resize((size_t)X * Z); for (unsigned int z = 0; z < Z; z++) for (unsigned int x = 0; x < X; x++) img[(size_t)z * X + x] = field[((size_t)z * Y + s) * X + x]; }   // xz at y = s
    else { img.resize((size_t)Y * Z); for (unsigned int z = 0; z < Z; z++) for (unsigned int y = 0; y < Y; y++) img[(size_t)z * Y + y] = field[((size_t)z * Y + y) * X + s]; }            // yz at x = s
    slice_captures_.push_back(img);
  }
  for (size_t i = 0; i < mesh_to_capture_.size(); i++)
    if (mesh_to_capture_[i] == current_step_) mesh_captures_.push_back(field);
}

void App::resetPressureMesh() { m_mesh.resetPressures(); current_step_ = 0; }
void App::close() { m_mesh.destroyPartitions(); }

std::vector<float> App::getResponse(unsigned int rec) {
  std::vector<float> r(m_parameters.getNumSteps(), 0.f);
  for (unsigned int i = 0; i < m_parameters.getNumSteps(); i++) r.at(i) = getResponseSampleAt(i, rec);
  return r;
}
std::vector<double> App::getResponseDouble(unsigned int rec) {
  std::vector<double> r(m_parameters.getNumSteps(), 0.0);
  for (unsigned int i = 0; i < m_parameters.getNumSteps(); i++) r.at(i) = getResponseDoubleSampleAt(i, rec);
  return r;
}

// reference App.cpp:445-500
float App::getVolume() {
  const float n = (float)m_mesh.getNumberOfAirElements() + (float)m_mesh.getNumberOfBoundaryElements();
  const float dx = m_parameters.getDx();
  return dx * dx * dx * n;
}
float App::getTotalAborptionArea(unsigned int octave) {
  float a = 0.f;
  for (unsigned int i = 0; i < m_geometry.getNumberOfTriangles(); i++) {
    const float r = admitance2Reflection(m_materials.getSurfaceCoefAt(i, octave));
    a += m_geometry.getSurfaceAreaAt(i) * (1 - r * r);
  }
  return a;
}
float App::getSabine(unsigned int octave) { return 0.1611f * getVolume() / getTotalAborptionArea(octave); }
float App::getEyring(unsigned int octave) {
  return 0.1611f * getVolume() / (-1.f * m_geometry.getTotalSurfaceArea() * logf(1 - m_materials.getMeanAbsorption(octave)));
}

}  // namespace FDTD
