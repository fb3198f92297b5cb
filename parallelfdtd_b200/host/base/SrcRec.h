// SrcRec.h -- Position / Source / Receiver with the reference's API (reference src/base/SrcRec.h:33-168,
// src/base/SrcRec.cpp:26-34).  Header-only.
#pragma once
#include <string>
#include "../math/geomMath.h"

enum SrcType { SRC_HARD, SRC_SOFT, SRC_TRANSPARENT };
enum InputType { IMPULSE, GAUSSIAN, SINE, DATA };

class Position {
 public:
  Position() : p_() {}
  Position(float x, float y, float z) : p_(x, y, z) {}
  Position(float x, float y) : p_(x, y, 0.f) {}
  virtual ~Position() {}
  nv::Vec3f getP() { return p_; }
  // voxel index of the position: ROUND(p / dx) per axis with dx = c / (fs * lambda) in float
  nv::Vec3i getElementIdx(unsigned spatial_fs, float c, float lambda) {
    const float dx = c / ((float)spatial_fs * lambda);
    return nv::Vec3i((int)nv::ROUND(p_.x / dx), (int)nv::ROUND(p_.y / dx), (int)nv::ROUND(p_.z / dx));
  }
 protected:
  nv::Vec3f p_;
};

class Source : public Position {
 public:
  Source() : Position() { init_(SRC_HARD, IMPULSE, 0, 0); }
  Source(float x, float y, float z) : Position(x, y, z) { init_(SRC_HARD, IMPULSE, 0, 0); }
  Source(float x, float y, float z, enum SrcType t) : Position(x, y, z) { init_(t, IMPULSE, 0, 0); }
  Source(float x, float y, float z, enum SrcType t, unsigned int group) : Position(x, y, z) { init_(t, IMPULSE, group, 0); }
  Source(float x, float y, float z, enum SrcType t, enum InputType in, unsigned int data_idx) : Position(x, y, z) { init_(t, in, 0, data_idx); }
  Source(float x, float y) : Position(x, y) { init_(SRC_HARD, IMPULSE, 0, 0); }
  void setSourceType(enum SrcType t) { source_type_ = t; }
  void setInputType(enum InputType t) { input_type_ = t; }
  void setGroup(unsigned int g) { group_ = g; }
  void setInputDataIdx(unsigned int i) { input_data_idx_ = i; }
  unsigned int getInputDataIdx() const { return input_data_idx_; }
  enum SrcType getSourceType() const { return source_type_; }
  enum InputType getInputType() const { return input_type_; }
  unsigned int getGroup() const { return group_; }
 private:
  void init_(SrcType t, InputType in, unsigned int g, unsigned int d) { source_type_ = t; input_type_ = in; group_ = g; input_data_idx_ = d; }
  SrcType source_type_;
  InputType input_type_;
  unsigned int group_;
  unsigned int input_data_idx_;
};

class Receiver : public Position {
 public:
  Receiver() : Position() {}
  Receiver(float x, float y, float z) : Position(x, y, z) {}
  Receiver(float x, float y) : Position(x, y) {}
  void setOutputFp(std::string fp) { output_fp_ = fp; }
 private:
  std::string output_fp_;
};
