// SimulationParameters.h -- host-side run configuration with the reference's public API
// (reference src/base/SimulationParameters.h:39-177).  Re-implemented for the B200 build: no CUDA or
// Boost dependency; per-step source values can also be produced as whole tables, which is what
// launchFDTD3d[Double] uploads once instead of one H2D copy per source per step.
#pragma once
#include <string>
#include <vector>
#include "../math/geomMath.h"
#include "SrcRec.h"

// 0..2 as in the reference; the interpolated compact schemes are appended so existing numeric values
// used by MEX / Python callers keep their meaning
enum UpdateType { SRL_FORWARD, SHARED, SRL, IISO, IWB };

class SimulationParameters {
 public:
  SimulationParameters();
  ~SimulationParameters() {}

  void readGridIr(std::string ir_fp);
  void setGridIr(const std::vector<float>& ir) { grid_ir_ = ir; ++generation_; }
  void setUpdateType(enum UpdateType update_type);
  void setC(float c) { c_ = c; ++generation_; }
  void setLambda(double lambda) { lambda_ = lambda; ++generation_; }
  void setOctave(unsigned int octave) { octave_ = octave; ++generation_; }
  void setNumSteps(unsigned int num_steps) { num_steps_ = num_steps; ++generation_; }
  void setSpatialFs(unsigned int spatial_fs) { spatial_fs_ = spatial_fs; ++generation_; }
  void setBoundingBox(nv::Vec3f bb_min, nv::Vec3f bb_max) { bounding_box_min_ = bb_min; bounding_box_max_ = bb_max; ++generation_; }
  void setAddPaddingToElementIdx(bool v) { add_padding_to_element_idx_ = v; ++generation_; }

  enum UpdateType getUpdateType() const { return update_type_; }
  float getC() const { return c_; }
  double getLambda() const { return lambda_; }
  float getDx() const;
  unsigned int getOctave() const { return octave_; }
  unsigned int getNumSteps() const { return num_steps_; }
  unsigned int getSpatialFs() const { return spatial_fs_; }
  unsigned int getStepAtTime(float t) { return (unsigned int)(spatial_fs_ * t); }

  void addSource(float x, float y, float z);
  void addSource(Source src);
  void addReceiver(float x, float y, float z) { receivers_.push_back(Receiver(x, y, z)); ++generation_; }
  void addReceiver(Receiver rec) { receivers_.push_back(rec); ++generation_; }
  void addSourceDData(float* d_vector);
  void removeSource(unsigned int i);
  void removeReceiver(unsigned int i);
  void updateSourceAt(unsigned int i, Source src);
  void updateReceiverAt(unsigned int i, Receiver rec);
  void resetSourcesAndReceivers();

  void addInputData(float* data, unsigned int number_of_samples);
  void addInputData(std::vector<float> data) { source_input_data_.push_back(data); ++generation_; }
  void addInputDataDouble(std::vector<double> data) { source_input_data_double_.push_back(data); ++generation_; }

  float getSourceSample(unsigned int source_idx, unsigned int step);
  double getSourceSampleDouble(unsigned int source_idx, unsigned int step);
  float* getSourceDData(unsigned int source_idx) { return d_source_output_data_.at(source_idx); }
  float getInputDataSample(unsigned int idx, unsigned int sample);
  double getInputDataSampleDouble(unsigned int idx, unsigned int sample);
  float getGridIrDataSample(unsigned int sample);

  nv::Vec3i getSourceElementCoordinates(unsigned int source_idx);
  nv::Vec3i getReceiverElementCoordinates(unsigned int receiver_idx);
  unsigned int getSourceElementIdx(unsigned int source_idx, unsigned int dim_x, unsigned int dim_y);
  unsigned int getReceiverElementIdx(unsigned int receiver_idx, unsigned int dim_x, unsigned int dim_y);
  unsigned int getNumSources() const { return (unsigned int)sources_.size(); }
  unsigned int getNumReceivers() const { return (unsigned int)receivers_.size(); }
  float* getSourceVectorAt(unsigned int source_idx);
  Source getSource(unsigned int i) const { return sources_.at(i); }
  Receiver getReceiver(unsigned int i) const { return receivers_.at(i); }

  float* getParameterPtr();
  double* getParameterPtrDouble();

  // ---- additions of this build --------------------------------------------------------------
  // [num_sources][num_steps] table of getSourceSample / getSourceSampleDouble; the transparent
  // source is evaluated as one running convolution per source, O(steps * ir_length) instead of the
  // reference's O(steps^2) per-step recomputation, with the same summation order per sample.
  void fillSourceTable(std::vector<float>& out, unsigned int num_steps);
  void fillSourceTableDouble(std::vector<double>& out, unsigned int num_steps);
  // Bumped by every call that can change a source / receiver position, a source sample or the step count: what the
  // step-by-step launcher (launchFDTD3dStep) compares to know that the tables it uploaded are still the right ones.
  unsigned long long generation() const { return generation_; }

 private:
  float getRegularSourceSample(unsigned int source_idx, unsigned int step);
  double getRegularSourceSampleDouble(unsigned int source_idx, unsigned int step);
  float getTransparentSourceSample(unsigned int source_idx, unsigned int step);
  double getTransparentSourceSampleDouble(unsigned int source_idx, unsigned int step);

  UpdateType update_type_;
  float c_;
  double lambda_;
  unsigned int octave_;
  unsigned int num_steps_;
  unsigned int spatial_fs_;
  nv::Vec3f bounding_box_min_, bounding_box_max_;
  bool add_padding_to_element_idx_;
  std::vector<Source> sources_;
  std::vector<Receiver> receivers_;
  std::vector<std::vector<float> > source_input_data_;
  std::vector<std::vector<double> > source_input_data_double_;
  std::vector<std::vector<float> > source_output_data_;
  std::vector<float*> d_source_output_data_;
  std::vector<float> parameter_vec_;
  std::vector<double> parameter_vec_double_;
  std::vector<float> grid_ir_;
  unsigned long long generation_ = 0;
};
