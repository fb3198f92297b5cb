// SimulationParameters.cpp -- see the header.  Numerical behaviour follows the reference
// (src/base/SimulationParameters.cpp); line references are given per function.
#include "SimulationParameters.h"

#include <cmath>
#include <fstream>
#include <stdexcept>

#include "../logger.h"

SimulationParameters::SimulationParameters()
    : update_type_(SRL_FORWARD),
      c_(344.f),
      lambda_((double)1 / std::sqrt((double)3)),   // reference SimulationParameters.h:47
      octave_(0),
      num_steps_(1),
      spatial_fs_(7000),
      add_padding_to_element_idx_(true) {}

// whitespace-separated float list (reference src/io/FileReader.cpp readFloat)
void SimulationParameters::readGridIr(std::string ir_fp) {
  std::ifstream in(ir_fp.c_str());
  grid_ir_.clear();
  ++generation_;
  if (!in) {
    log_msg<LOG_ERROR>(L"SimulationParameters::readGridIr - cannot open %s") % ir_fp;
    return;
  }
  float v;
  while (in >> v) grid_ir_.push_back(v);
}

// reference :123-137 -- every 7-point type resets the Courant number to sqrt(1/3) (NOT the
// constructor's 1/sqrt(3): one ulp apart in double, and its square is exactly the double nearest 1/3).
// Interpolated schemes run at their own stability limits.
void SimulationParameters::setUpdateType(enum UpdateType update_type) {
  update_type_ = update_type;
  ++generation_;
  switch (update_type) {
    case SRL_FORWARD:
    case SHARED:
    case SRL: lambda_ = std::sqrt((double)1 / 3); break;
    case IISO: lambda_ = std::sqrt((double)3) / 2; break;
    case IWB: lambda_ = 1.0; break;
  }
}

// reference :362-364
float SimulationParameters::getDx() const { return (float)((double)c_ / ((double)spatial_fs_ * lambda_)); }

void SimulationParameters::addSource(float x, float y, float z) { sources_.push_back(Source(x, y, z)); ++generation_; }
void SimulationParameters::addSource(Source src) { sources_.push_back(src); ++generation_; }
void SimulationParameters::addSourceDData(float* d_vector) {
  if (d_source_output_data_.size() < getNumSources()) d_source_output_data_.push_back(d_vector);
}

void SimulationParameters::removeSource(unsigned int i) {
  if (i >= sources_.size()) throw std::out_of_range("SimulationParameters::removeSource: idx out of range");
  sources_.erase(sources_.begin() + i);
  ++generation_;
}
void SimulationParameters::removeReceiver(unsigned int i) {
  if (i >= receivers_.size()) throw std::out_of_range("SimulationParameters::removeReceiver: idx out of range");
  receivers_.erase(receivers_.begin() + i);
  ++generation_;
}
void SimulationParameters::updateSourceAt(unsigned int i, Source src) { sources_.at(i) = src; ++generation_; }
// the reference bounds-checks against the SOURCE list here (:62-66); .at() on the receiver list is what throws
void SimulationParameters::updateReceiverAt(unsigned int i, Receiver rec) { receivers_.at(i) = rec; ++generation_; }
void SimulationParameters::resetSourcesAndReceivers() { receivers_.clear(); sources_.clear(); ++generation_; }

void SimulationParameters::addInputData(float* data, unsigned int number_of_samples) {
  std::vector<float> v(number_of_samples, 0.f);
  if (data) for (unsigned int i = 0; i < number_of_samples; i++) v[i] = data[i];
  else log_msg<LOG_WARNING>(L"SimulationParameters::addInputData : invalid input data (NULL)");
  source_input_data_.push_back(v);
  ++generation_;
}

// reference :160-186: out-of-range data index throws (vector::at), sample index past the end reads 0
float SimulationParameters::getInputDataSample(unsigned int idx, unsigned int sample) {
  const std::vector<float>& v = source_input_data_.at(idx);
  return sample < v.size() ? v[sample] : 0.f;
}
double SimulationParameters::getInputDataSampleDouble(unsigned int idx, unsigned int sample) {
  const std::vector<double>& v = source_input_data_double_.at(idx);
  return sample < v.size() ? v[sample] : 0.0;
}
float SimulationParameters::getGridIrDataSample(unsigned int sample) { return sample < grid_ir_.size() ? grid_ir_[sample] : 0.f; }

// reference :200-224: ROUND(p/dx) with dx in float from (float)lambda, +1 when padding is on
nv::Vec3i SimulationParameters::getSourceElementCoordinates(unsigned int source_idx) {
  nv::Vec3i r = getSource(source_idx).getElementIdx(getSpatialFs(), getC(), (float)getLambda());
  if (add_padding_to_element_idx_) { r.x += 1; r.y += 1; r.z += 1; }
  return r;
}
nv::Vec3i SimulationParameters::getReceiverElementCoordinates(unsigned int receiver_idx) {
  nv::Vec3i r = getReceiver(receiver_idx).getElementIdx(getSpatialFs(), getC(), (float)getLambda());
  if (add_padding_to_element_idx_) { r.x += 1; r.y += 1; r.z += 1; }
  return r;
}
// reference :226-260: the source variant adds the padding a second time (as written), the receiver one once
unsigned int SimulationParameters::getSourceElementIdx(unsigned int source_idx, unsigned int dim_x, unsigned int dim_y) {
  nv::Vec3i p = getSourceElementCoordinates(source_idx);
  const int inc = add_padding_to_element_idx_ ? 1 : 0;
  return (unsigned int)((p.z + inc) * (int)(dim_x * dim_y) + (p.y + inc) * (int)dim_x + (p.x + inc));
}
unsigned int SimulationParameters::getReceiverElementIdx(unsigned int receiver_idx, unsigned int dim_x, unsigned int dim_y) {
  nv::Vec3i p = getReceiver(receiver_idx).getElementIdx(getSpatialFs(), getC(), (float)getLambda());
  const int inc = add_padding_to_element_idx_ ? 1 : 0;
  return (unsigned int)((p.z + inc) * (int)(dim_x * dim_y) + (p.y + inc) * (int)dim_x + (p.x + inc));
}

// ---- waveforms (reference :262-360) -----------------------------------------------------------------
float SimulationParameters::getRegularSourceSample(unsigned int source_idx, unsigned int step) {
  const Source& s = sources_.at(source_idx);
  switch (s.getInputType()) {
    case IMPULSE: return step == 1 ? 1.f : 0.f;                       // fires at step 1 (:266-272)
    case GAUSSIAN: { const float t0 = 40, width = 4; const float e = (float)(step - t0) / width; return expf(-0.5f * (e * e)); }
    case SINE: { const float freq = 120; const float t = (float)step / (float)getSpatialFs(); return sinf(2.f * (float)PI * freq * t); }
    case DATA: return getInputDataSample(s.getInputDataIdx(), step);
  }
  return 0.f;
}
double SimulationParameters::getRegularSourceSampleDouble(unsigned int source_idx, unsigned int step) {
  const Source& s = sources_.at(source_idx);
  switch (s.getInputType()) {
    case IMPULSE: return step == 1 ? 1.0 : 0.0;
    case GAUSSIAN: { const double t0 = 40, width = 4; const double e = (float)(step - t0) / width; return exp(-0.5f * (e * e)); }
    case SINE: { const double freq = 120; const double t = (double)step / (double)getSpatialFs(); return sin(2.f * (double)PI * freq * t); }
    case DATA: return getInputDataSampleDouble(s.getInputDataIdx(), step);
  }
  return 0.0;
}
// regular(step) - sum_{i<step} ir[step-i] * regular(i), summed in increasing i (:300-310, :350-360)
float SimulationParameters::getTransparentSourceSample(unsigned int source_idx, unsigned int step) {
  float acc = 0.f;
  for (unsigned int i = 0; i < step; i++) acc += getGridIrDataSample(step - i) * getRegularSourceSample(source_idx, i);
  return getRegularSourceSample(source_idx, step) - acc;
}
double SimulationParameters::getTransparentSourceSampleDouble(unsigned int source_idx, unsigned int step) {
  double acc = 0.f;
  for (unsigned int i = 0; i < step; i++) acc += getGridIrDataSample(step - i) * getRegularSourceSampleDouble(source_idx, i);
  return getRegularSourceSampleDouble(source_idx, step) - acc;
}
float SimulationParameters::getSourceSample(unsigned int source_idx, unsigned int step) {
  float sample = 0.f;
  if (getSource(source_idx).getSourceType() == SRC_TRANSPARENT) sample += getTransparentSourceSample(source_idx, step);
  else sample += getRegularSourceSample(source_idx, step);
  return sample;
}
double SimulationParameters::getSourceSampleDouble(unsigned int source_idx, unsigned int step) {
  double sample = 0.f;
  if (getSource(source_idx).getSourceType() == SRC_TRANSPARENT) sample += getTransparentSourceSampleDouble(source_idx, step);
  else sample += getRegularSourceSampleDouble(source_idx, step);
  return sample;
}

float* SimulationParameters::getSourceVectorAt(unsigned int source_idx) {
  std::vector<float> v(getNumSteps(), 0.f);
  for (unsigned int i = 0; i < getNumSteps(); i++) v[i] = getSourceSample(source_idx, i);
  if (source_output_data_.size() < getNumSources()) source_output_data_.resize(getNumSources());
  source_output_data_.at(source_idx) = v;
  return &(source_output_data_.at(source_idx)[0]);
}

template <typename T, typename RegularFn>
static void fill_table(std::vector<T>& out, unsigned int n_src, unsigned int num_steps, const std::vector<bool>& transparent,
                       const std::vector<float>& ir, RegularFn regular) {
  out.assign((size_t)n_src * num_steps, (T)0);
  std::vector<T> reg(num_steps);
  for (unsigned int s = 0; s < n_src; s++) {
    for (unsigned int n = 0; n < num_steps; n++) reg[n] = regular(s, n);
    T* row = &out[(size_t)s * num_steps];
    if (!transparent[s]) {
      for (unsigned int n = 0; n < num_steps; n++) { T v = (T)0; v += reg[n]; row[n] = v; }
      continue;
    }
    for (unsigned int n = 0; n < num_steps; n++) {
      // same order as the reference: i = 0 .. n-1; terms whose IR index is past the table are exactly 0
      T acc = (T)0.f;
      const unsigned int i0 = (ir.size() > 0 && n >= ir.size()) ? n - (unsigned int)ir.size() + 1 : 0;
      for (unsigned int i = i0; i < n; i++) acc += ir[n - i] * reg[i];
      T v = (T)0; v += reg[n] - acc; row[n] = v;
    }
  }
}

void SimulationParameters::fillSourceTable(std::vector<float>& out, unsigned int num_steps) {
  std::vector<bool> tr(getNumSources());
  for (unsigned int s = 0; s < getNumSources(); s++) tr[s] = sources_[s].getSourceType() == SRC_TRANSPARENT;
  fill_table<float>(out, getNumSources(), num_steps, tr, grid_ir_, [this](unsigned int s, unsigned int n) { return getRegularSourceSample(s, n); });
}
void SimulationParameters::fillSourceTableDouble(std::vector<double>& out, unsigned int num_steps) {
  std::vector<bool> tr(getNumSources());
  for (unsigned int s = 0; s < getNumSources(); s++) tr[s] = sources_[s].getSourceType() == SRC_TRANSPARENT;
  fill_table<double>(out, getNumSources(), num_steps, tr, grid_ir_, [this](unsigned int s, unsigned int n) { return getRegularSourceSampleDouble(s, n); });
}

// reference :379-396
float* SimulationParameters::getParameterPtr() {
  parameter_vec_.assign(4, 0.f);
  parameter_vec_[0] = (float)getLambda();
  parameter_vec_[1] = (float)(getLambda() * getLambda());
  parameter_vec_[2] = 1.f / 3.f;
  parameter_vec_[3] = (float)getOctave();
  return &parameter_vec_[0];
}
double* SimulationParameters::getParameterPtrDouble() {
  parameter_vec_double_.assign(4, 0.0);
  parameter_vec_double_[0] = getLambda();
  parameter_vec_double_[1] = getLambda() * getLambda();
  parameter_vec_double_[2] = (double)1 / (double)3;
  parameter_vec_double_[3] = (double)getOctave();
  return &parameter_vec_double_[0];
}
