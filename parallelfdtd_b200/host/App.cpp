// App.cpp -- see App.h.  Behaviour follows reference src/App.cpp where the hot path is concerned:
// initializeMesh (:147-234) -> setupMesh + makePartition, runSimulation (:308-348) ->
// launchFDTD3d[Double], executeStep (:407-436) -> launchFDTD3dStep + captures.
#include "App.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>

#include "kernels/kernels3d.h"

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
// process' CUDA context (torch, MATLAB's GPU arrays), so this build only releases its own partitions and the device
// blocks the library keeps for the next mesh
void App::resetDevices() {
  m_mesh.destroyPartitions();
  pfdtd_release_cached_memory(-1);
}
void App::initializeDevices() {
  queryDevices();
  if (number_of_devices_ < 1) { c_log_msg(LOG_ERROR, "App::initializeDevices - no CUDA device"); throw(-1); }
}

void App::initializeGeometryFromFile(std::string geometry_fp) {
  if (!m_file_reader.readVTK(&m_geometry, geometry_fp)) {
    log_msg<LOG_ERROR>(L"App::initializeGeometryFromFile - invalid file: %s") % geometry_fp;
    throw(-1);
  }
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
  if (m_materials.getNumberOfUniqueMaterials() == 0) m_materials.setGlobalMaterial(uniform_surfaces_(), 0.f);
  const uint3 block = make_uint3(32, 4, 1);                     // reference App.cpp:193
  const unsigned int type = (unsigned int)m_parameters.getUpdateType();
  // frequency-dependent boundaries: the material rows are digital impedance filters of this order (0 = the
  // reference's scalar admittance per octave)
  m_mesh.setOption(PFDTD_OPT_DIF_ORDER, (long long)m_materials.getFilterOrder());
  // capacity check before anything is allocated (reference App.cpp:148-174): estimated voxels x the bytes this build
  // keeps per voxel (two fields + three node bytes; the reference counts 8 / 18) against the memory of all devices
  {
    const float dx = m_parameters.getDx();
    const nv::Vec3f bb = m_geometry.getBoundingBox();
    const double est = vol_bid_.empty() ? (double)(bb.x / dx + 3) * (double)(bb.y / dx + 3) * (double)(bb.z / dx + 3)
                                        : (double)vol_dim_[0] * vol_dim_[1] * vol_dim_[2];
    const double need_mb = est * (2.0 * (m_mesh.isDouble() ? 8 : 4) + 3.0) / 1e6;
    double have_mb = 0;
    for (size_t i = 0; i < device_mem_sizes_.size(); i++) have_mb += device_mem_sizes_[i];
    log_msg<LOG_INFO>(L"App::initializeMesh - estimated size: %f voxels, dx %f, %f MB of %f MB") % est % dx % need_mb % have_mb;
    if (!device_mem_sizes_.empty() && need_mb > have_mb) {
      log_msg<LOG_ERROR>(L"App::initializeMesh - estimated size %f MB exceeds the devices' %f MB, exiting") % need_mb % have_mb;
      close();
      throw(-1);
    }
  }
  if (vol_bid_.empty()) {
    // triangle mesh -> node volumes on the device (reference App.cpp:181-190 voxelizeGeometry), adopted by setupMesh
    if (m_geometry.getNumberOfTriangles() == 0) { c_log_msg(LOG_ERROR, "App::initializeMesh - no geometry"); throw(-1); }
    unsigned char *d_bid = 0, *d_mat = 0;
    unsigned int vx = 0, vy = 0, vz = 0;
    pfdtd_safe(pfdtd_voxelize_device(-1, m_geometry.getVerticePtr(), m_geometry.getNumberOfVertices(), m_geometry.getIndexPtr(),
                                     m_geometry.getNumberOfTriangles(), m_materials.getMaterialIdxPtr(), m_parameters.getDx(), &d_bid, &d_mat,
                                     &vx, &vy, &vz), "App::initializeMesh - voxelizeGeometry");
    vol_dim_[0] = vx; vol_dim_[1] = vy; vol_dim_[2] = vz;
    const uint3 dim = make_uint3(vx, vy, vz);
    if (m_mesh.isDouble())
      m_mesh.setupMeshDouble(d_bid, d_mat, m_materials.getNumberOfUniqueMaterials(), m_materials.getMaterialCoefficientPtrDouble(),
                             m_parameters.getParameterPtrDouble(), dim, block, type);
    else
      m_mesh.setupMesh(d_bid, d_mat, m_materials.getNumberOfUniqueMaterials(), m_materials.getMaterialCoefficientPtr(),
                       m_parameters.getParameterPtr(), dim, block, type);
  } else {
    const uint3 dim = make_uint3(vol_dim_[0], vol_dim_[1], vol_dim_[2]);
    if (m_mesh.isDouble())
      m_mesh.setupMeshHost(&vol_bid_[0], &vol_mat_[0], m_materials.getNumberOfUniqueMaterials(), m_materials.getMaterialCoefficientPtrDouble(),
                           m_parameters.getParameterPtrDouble(), dim, block, type);
    else
      m_mesh.setupMeshHost(&vol_bid_[0], &vol_mat_[0], m_materials.getNumberOfUniqueMaterials(), m_materials.getMaterialCoefficientPtr(),
                           m_parameters.getParameterPtr(), dim, block, type);
  }
  num_elements_ = m_mesh.getNumberOfElements();
  // Partition count (reference App.cpp:217-233: one partition below 90e6 voxels -- 45e6 in double --, otherwise
  // `number_of_partitions`, sized for Kepler-era memory).  Same shape with B200 numbers: one partition while a
  // slab would be smaller than 2^28 voxels (a one-plane halo costs < 0.5 % of a slab that size and the interior
  // launch hides it), then one more device per 2^28 voxels up to the devices present, and never fewer than the
  // memory needs.  forcePartitionTo keeps its meaning; 0 (the reference would call makePartition(0)) means "choose".
  unsigned int n = 1;
  const double elements = (double)m_mesh.getNumberOfElements64();
  const double bytes = elements * (2.0 * (m_mesh.isDouble() ? 8 : 4) + 3.0);
  if (force_partition_to_ > 0 && force_partition_to_ <= number_of_devices_) n = (unsigned int)force_partition_to_;
  else {
    const double element_limit = m_mesh.isDouble() ? 134217728.0 : 268435456.0;
    const double cap = device_mem_sizes_.empty() ? 150e9 : 0.9 * 1e6 * (double)device_mem_sizes_[0];
    const unsigned int max_n = (unsigned int)std::max(1, number_of_devices_);
    while (n < max_n && (elements / n >= 2 * element_limit || bytes / n > cap)) n++;
    (void)number_of_partitions;
  }
  m_mesh.makePartition(n);
  current_step_ = 0;
}

void App::runVisualization() {
  m_mesh.setDouble(false);                                        // reference App.cpp:285-289
  m_parameters.setNumSteps(m_parameters.getSpatialFs() * 2);
  force_partition_to_ = 1;
  initializeMesh(1);
  const unsigned int steps = m_parameters.getNumSteps();
  responses_.assign((size_t)steps * m_parameters.getNumReceivers() + 1, 0.f);
  log_msg<LOG_INFO>(L"App::runVisualization - Volume: %f") % getVolume();
  log_msg<LOG_INFO>(L"App::runVisualization - TotalAbsorptionArea: %f, octave: %u") % getTotalAborptionArea(0) % m_parameters.getOctave();
  log_msg<LOG_INFO>(L"App::runVisualization - Sabine RT: %f") % getSabine(0);
  log_msg<LOG_INFO>(L"App::runVisualization - Eyrting RT: %f") % getEyring(0);
  log_msg<LOG_WARNING>(L"App::runVisualization - no OpenGL window in this build, stepping headless for %u steps") % steps;
  for (unsigned int i = 0; i < steps; i++) {
    executeStep();
    if (m_interrupt && m_interrupt()) break;
  }
  responses_.resize((size_t)steps * m_parameters.getNumReceivers());
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
// captureMesh, called from executeStep, App.cpp:421-431): the slice is gathered on the device(s) from the planes
// each slab owns and only the slice crosses PCIe.  A slice beyond the mesh is logged and skipped like the reference.
void App::captureIfDue_() {
  const unsigned int X = m_mesh.getDimX(), Y = m_mesh.getDimY(), Z = m_mesh.getDimZ();
  for (size_t i = 0; i < step_to_capture_.size(); i++) {
    if (step_to_capture_[i] != current_step_) continue;
    const unsigned int s = slice_to_capture_[i], o = slice_orientation_[i];
    const unsigned int lim = o == 0 ? Z : (o == 1 ? Y : X);
    if (o > 2 || s >= lim) {
      log_msg<LOG_INFO>(L"App::capture - slice %u out of bounds %u, no capture made") % s % lim;
      continue;
    }
    CaptureShape sh;
    sh.cols = o == 2 ? Y : X; sh.rows = o == 0 ? Y : Z; sh.slice = s; sh.orientation = o; sh.step = current_step_;
    std::vector<float> img((size_t)sh.rows * sh.cols);
    std::vector<unsigned char> pos((size_t)sh.rows * sh.cols);
    m_mesh.captureSlice<float>(s, o, &img[0], &pos[0]);
    slice_captures_.push_back(img);
    slice_positions_.push_back(pos);
    slice_shapes_.push_back(sh);
  }
  for (size_t i = 0; i < mesh_to_capture_.size(); i++) {
    if (mesh_to_capture_[i] != current_step_) continue;
    mesh_captures_.push_back(std::vector<float>((size_t)X * Y * Z));
    m_mesh.captureMesh<float>(&mesh_captures_.back()[0]);
  }
}

// the pixel mapping of the reference's capture callback (App::saveBitmap, App.cpp:492-567) without the TGA writer:
// solid nodes white, air/boundary nodes green for positive and blue for negative pressure on a log scale of
// capture_db_ / 10 decades.  RGBA, row-major like the capture.
std::vector<unsigned char> App::getSliceCaptureRGBA(unsigned int i) {
  const std::vector<float>& d = slice_captures_.at(i);
  const std::vector<unsigned char>& pos = slice_positions_.at(i);
  std::vector<unsigned char> px(d.size() * 4, 0);
  const float dB = capture_db_ / 10.f;
  for (size_t e = 0; e < d.size(); e++) {
    unsigned char* q = &px[4 * e];
    q[3] = 255;
    const unsigned char c_pos = pos[e];
    const bool centred = m_parameters.getUpdateType() == SRL;
    // forward schemes: switch bit clear = solid; the reference paints "switch set and K != 6" white, which marks the
    // boundary shell; the centred scheme's position byte has no neighbour count, so there solid (switch clear) is white
    const bool white = centred ? (c_pos >> 7) == 0 : ((c_pos >> 7) == 1 && (c_pos & 0x7F) != 6);
    if (white) { q[0] = q[1] = q[2] = 255; continue; }
    const float c = d[e];
    float v = (std::log10(c * c) + dB) / dB;
    v = v > 0.f ? (v < 1.f ? v : 1.f) : 0.f;
    if (c >= 0.f) q[1] = (unsigned char)(v * 255); else q[2] = (unsigned char)(v * 255);
  }
  return px;
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
