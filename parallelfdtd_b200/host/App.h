// App.h -- `FDTD::App`, the facade the MEX gateway, the Python module and main() drive (reference
// src/App.h:50-396).  Same public methods for everything on or next to the time-stepping path; the
// OpenGL window / Voxelizer members are not part of this build.
// Geometry reaches the solver either as voxelizer-style node volumes (setVoxelVolumes, what the
// reference obtains from the third-party Voxelizer) or, for axis-aligned rooms, through the
// built-in box voxelizer of initializeMesh.
#pragma once
#include <string>
#include <vector>
#include "base/GeometryHandler.h"
#include "base/MaterialHandler.h"
#include "base/SimulationParameters.h"
#include "global_includes.h"
#include "io/FileReader.h"
#include "kernels/cudaMesh.h"

typedef bool (*InterruptCallback)(void);
typedef void (*ProgressCallback)(int, int, float);

namespace FDTD {

class App {
 public:
  App();
  ~App();

  SimulationParameters m_parameters;
  GeometryHandler m_geometry;
  MaterialHandler m_materials;
  CudaMesh m_mesh;
  FileReader m_file_reader;
  InterruptCallback m_interrupt;
  ProgressCallback m_progress;

  void queryDevices();
  void resetDevices();
  void initializeDevices();
  // VTK POLYDATA in inches (reference App.cpp:133-139, FileReader.cpp:41-101); throws -1 on an unreadable file
  void initializeGeometryFromFile(std::string geometry_fp);
  void initializeGeometry(unsigned int* indices, float* vertices, unsigned int number_of_indices, unsigned int number_of_vertices);
  void setupDefaultCallbacks();
  void initializeMesh(unsigned int number_of_partitions);

  // voxelizer-style volumes [vz][vy][vx]: `bid` 0..27 and material index per voxel
  void setVoxelVolumes(const unsigned char* bid, const unsigned char* mat, unsigned int vx, unsigned int vy, unsigned int vz);

  // The reference opens its OpenGL viewer here and steps the solver from the window's idle callback (App.cpp:278-306).
  // This build has no window: same preparation (single precision, one partition, 2 x fs steps), then the steps run
  // headless through executeStep so the captures and responses a viewer session would produce are still there.
  void runVisualization();
  void runSimulation();
  void runCapture();
  void close();
  void executeStep();
  void resetPressureMesh();
  void invertTime() { step_direction_ *= -1; }

  float getTimePerStep() { return time_per_step_; }
  unsigned int getNumElements() { return num_elements_; }
  unsigned int getResponseSize() { return (unsigned int)(m_mesh.isDouble() ? responses_double_.size() : responses_.size()); }
  float getMvoxPerSec() { return (float)((1.f / time_per_step_ * m_mesh.getNumberOfElements64()) / 1e6); }
  float* getResponsePointer() { return &responses_[0]; }
  float getResponseSampleAt(unsigned int step, unsigned int rec) { return responses_.at((size_t)m_parameters.getNumSteps() * rec + step); }
  double getResponseDoubleSampleAt(unsigned int step, unsigned int rec) { return responses_double_.at((size_t)m_parameters.getNumSteps() * rec + step); }
  float* getMeshCaptureAt(unsigned int i) { return &mesh_captures_.at(i)[0]; }
  unsigned int getNumberOfMeshCaptures() { return (unsigned int)mesh_captures_.size(); }
  void addSliceToCapture(unsigned int slice, unsigned int step, unsigned int orientation) {
    step_to_capture_.push_back(step); slice_to_capture_.push_back(slice); slice_orientation_.push_back(orientation); }
  void addMeshToCapture(unsigned int step) { mesh_to_capture_.push_back(step); }
  const std::vector<float>& getSliceCaptureAt(unsigned int i) { return slice_captures_.at(i); }
  const std::vector<unsigned char>& getSlicePositionCaptureAt(unsigned int i) { return slice_positions_.at(i); }
  struct CaptureShape { unsigned int rows, cols, slice, orientation, step; };
  CaptureShape getSliceCaptureShapeAt(unsigned int i) { return slice_shapes_.at(i); }
  std::vector<unsigned char> getSliceCaptureRGBA(unsigned int i);
  unsigned int getNumberOfSliceCaptures() { return (unsigned int)slice_captures_.size(); }

  float getVolume();
  float getTotalAborptionArea(unsigned int octave);
  float getSabine(unsigned int octave);
  float getEyring(unsigned int octave);

  unsigned int current_step_;
  int step_direction_;

  // ---- the methods the Python binding exposes (reference App.h:344-395, AppPy.cpp:106-133)
  void addSource(float x, float y, float z, int type, int signal, int input_signal_idx) {
    m_parameters.addSource(Source(x, y, z, (enum SrcType)type, (enum InputType)signal, input_signal_idx)); }
  void addReceiver(float x, float y, float z) { m_parameters.addReceiver(x, y, z); }
  void setSpatialFs(unsigned int fs) { m_parameters.setSpatialFs(fs); }
  void setNumSteps(unsigned int num) { m_parameters.setNumSteps(num); }
  void setUpdateType(int i) { m_parameters.setUpdateType((enum UpdateType)i); }
  void setUniformMaterial(float R) { m_materials.setGlobalMaterial(uniform_surfaces_(), reflection2Admitance(R)); }
  // frequency-dependent boundaries (additions; MaterialHandler::addFilterMaterials / setGlobalFilter)
  void addSurfaceFilters(const float* coefs, unsigned int n_surfaces, unsigned int order) { m_materials.addFilterMaterials(coefs, n_surfaces, order); }
  void setUniformFilter(const std::vector<float>& b, const std::vector<float>& a) { m_materials.setGlobalFilter(uniform_surfaces_(), b, a); }
  std::vector<float> getResponse(unsigned int rec);
  std::vector<double> getResponseDouble(unsigned int rec);
  void setDouble(bool set_to) { m_mesh.setDouble(set_to); }
  void setForcePartitionTo(int num_partitions) { force_partition_to_ = num_partitions; }
  void setCapturedB(float db) { capture_db_ = db; }

 private:
  unsigned int uniform_surfaces_() { return m_geometry.getNumberOfTriangles() ? m_geometry.getNumberOfTriangles() : 1u; }
  void captureIfDue_();
  std::vector<unsigned int> step_to_capture_, slice_to_capture_, slice_orientation_, mesh_to_capture_;
  int number_of_devices_;
  int best_device_;
  std::vector<int> device_mem_sizes_;
  int force_partition_to_;
  float capture_db_;
  std::vector<float> responses_;
  std::vector<double> responses_double_;
  std::vector<std::vector<float> > mesh_captures_;
  std::vector<std::vector<float> > slice_captures_;
  std::vector<std::vector<unsigned char> > slice_positions_;
  std::vector<CaptureShape> slice_shapes_;
  float time_per_step_;
  unsigned int num_elements_;
  std::vector<unsigned char> vol_bid_, vol_mat_;
  unsigned int vol_dim_[3];
};

}  // namespace FDTD
