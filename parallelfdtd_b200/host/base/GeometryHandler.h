// GeometryHandler.h -- triangle soup container with the accessors App uses (reference
// src/base/GeometryHandler.h: vertices/indices, bounding box, per-triangle surface area).  Header-only.
#pragma once
#include <cmath>
#include <iterator>
#include <map>
#include <string>
#include <stdexcept>
#include <vector>
#include "../math/geomMath.h"

class GeometryHandler {
 public:
  GeometryHandler() {}
  // number_of_vertices counts the float COORDINATES (3 per vertex), as in the reference, whose callers pass
  // mxGetNumberOfElements(vertices) (reference GeometryHandler.cpp:87-108, matlab/mex_FDTD.cpp:89,198)
  void initialize(unsigned int* indices, float* vertices, unsigned int number_of_indices, unsigned int number_of_vertices) {
    indices_.assign(indices, indices + number_of_indices);
    vertices_.assign(vertices, vertices + number_of_vertices);
    for (size_t i = 0; i < indices_.size(); i++)
      if (indices_[i] >= number_of_vertices / 3) throw std::out_of_range("GeometryHandler::initialize: vertex index out of range");
    update_();
  }
  void initialize(std::vector<unsigned int> indices, std::vector<float> vertices) {
    initialize(indices.empty() ? 0 : &indices[0], vertices.empty() ? 0 : &vertices[0], (unsigned int)indices.size(),
               (unsigned int)vertices.size());
  }
  unsigned int getNumberOfTriangles() const { return (unsigned int)(indices_.size() / 3); }
  unsigned int getNumberOfVertices() const { return (unsigned int)(vertices_.size() / 3); }
  unsigned int getNumberOfIndices() const { return (unsigned int)indices_.size(); }
  unsigned int* getIndexPtr() { return indices_.empty() ? 0 : &indices_[0]; }
  float* getVerticePtr() { return vertices_.empty() ? 0 : &vertices_[0]; }
  nv::Vec3f getVertexAt(unsigned int i) const { return nv::Vec3f(vertices_.at(3 * i), vertices_.at(3 * i + 1), vertices_.at(3 * i + 2)); }
  nv::Vec3ui getTriangleAt(unsigned int t) const { return nv::Vec3ui(indices_.at(3 * t), indices_.at(3 * t + 1), indices_.at(3 * t + 2)); }
  nv::Vec3f getBoundingBox() const { return bb_max_ - bb_min_; }
  nv::Vec3f getBoundingBoxMin() const { return bb_min_; }
  nv::Vec3f getBoundingBoxMax() const { return bb_max_; }
  float getSurfaceAreaAt(unsigned int t) const { return areas_.at(t); }
  // named triangle groups (reference GeometryHandler.h:95-133, GeometryHandler.cpp:177-201): a list holding an index
  // below zero or beyond the triangle count is ignored
  void setLayerIndices(std::vector<int> indices, std::string name) {
    for (size_t i = 0; i < indices.size(); i++)
      if (indices[i] < 0 || indices[i] >= (int)getNumberOfTriangles()) return;
    layers_[name] = indices;
  }
  unsigned int getNumberOfLayers() const { return (unsigned int)layers_.size(); }
  std::string getLayerNameAt(int idx) const {
    if (idx < 0 || idx >= (int)layers_.size()) return std::string();
    std::map<std::string, std::vector<int> >::const_iterator it = layers_.begin();
    std::advance(it, idx);
    return it->first;
  }
  std::vector<int> getLayerIndices(const std::string& name) const {
    std::map<std::string, std::vector<int> >::const_iterator it = layers_.find(name);
    return it == layers_.end() ? std::vector<int>() : it->second;
  }
  float getTotalSurfaceArea() const { float s = 0.f; for (size_t i = 0; i < areas_.size(); i++) s += areas_[i]; return s; }

 private:
  void update_() {
    bb_min_ = nv::Vec3f(0, 0, 0); bb_max_ = nv::Vec3f(0, 0, 0);
    for (unsigned int i = 0; i < getNumberOfVertices(); i++) {
      nv::Vec3f v = getVertexAt(i);
      if (i == 0) { bb_min_ = v; bb_max_ = v; }
      bb_min_.set(std::fmin(bb_min_.x, v.x), std::fmin(bb_min_.y, v.y), std::fmin(bb_min_.z, v.z));
      bb_max_.set(std::fmax(bb_max_.x, v.x), std::fmax(bb_max_.y, v.y), std::fmax(bb_max_.z, v.z));
    }
    areas_.resize(getNumberOfTriangles());
    for (unsigned int t = 0; t < getNumberOfTriangles(); t++) {
      nv::Vec3ui tri = getTriangleAt(t);
      nv::Vec3f a = getVertexAt(tri.x), b = getVertexAt(tri.y), c = getVertexAt(tri.z);
      nv::Vec3f u = b - a, v = c - a;
      nv::Vec3f n(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
      areas_[t] = 0.5f * std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    }
  }
  std::vector<unsigned int> indices_;
  std::vector<float> vertices_;
  std::vector<float> areas_;
  std::map<std::string, std::vector<int> > layers_;
  nv::Vec3f bb_min_, bb_max_;
};
