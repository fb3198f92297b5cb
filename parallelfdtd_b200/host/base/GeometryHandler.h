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
  GeometryHandler() : longest_edge_(0.f) {}
  // The vertices are shifted so that the bounding box starts at the origin, and the shift is kept as the geometry
  // offset (reference GeometryHandler.cpp:45-85); the voxel grid starts at that corner.  Index values are not
  // checked here -- the reference's own tests hand in index lists that do not refer to the vertices
  // (tests/GeometryHandlerTest.cpp:10-56) -- but where they are used (surface areas: std::out_of_range; the
  // voxeliser: PFDTD_ERR_RANGE).
  void initialize(std::vector<unsigned int> indices, std::vector<float> vertices) {
    indices_ = indices;
    vertices_ = vertices;
    nv::Vec3f lo(0, 0, 0), hi(0, 0, 0);
    for (unsigned int i = 0; i < getNumberOfVertices(); i++) {
      const float* v = getVertexAt(i);
      if (i == 0) { lo.set(v[0], v[1], v[2]); hi = lo; }
      lo.set(std::fmin(lo.x, v[0]), std::fmin(lo.y, v[1]), std::fmin(lo.z, v[2]));
      hi.set(std::fmax(hi.x, v[0]), std::fmax(hi.y, v[1]), std::fmax(hi.z, v[2]));
    }
    offset_ = lo;
    for (unsigned int i = 0; i < getNumberOfVertices(); i++) {
      float* v = getVertexAt(i);
      v[0] -= lo.x; v[1] -= lo.y; v[2] -= lo.z;
    }
    bb_min_ = nv::Vec3f(0, 0, 0);
    bb_max_ = hi - lo;
    longest_edge_ = std::fmax(std::fmax(bb_max_.x, bb_max_.y), bb_max_.z);
  }
  // number_of_vertices counts the float COORDINATES (3 per vertex), as in the reference, whose callers pass
  // mxGetNumberOfElements(vertices) (reference GeometryHandler.cpp:87-108, matlab/mex_FDTD.cpp:89,198)
  void initialize(unsigned int* indices, float* vertices, unsigned int number_of_indices, unsigned int number_of_vertices) {
    initialize(std::vector<unsigned int>(indices, indices + number_of_indices), std::vector<float>(vertices, vertices + number_of_vertices));
  }
  unsigned int getNumberOfTriangles() const { return (unsigned int)(indices_.size() / 3); }
  unsigned int getNumberOfVertices() const { return (unsigned int)(vertices_.size() / 3); }
  unsigned int getNumberOfIndices() const { return (unsigned int)indices_.size(); }
  unsigned int* getIndexPtr() { return indices_.empty() ? 0 : &indices_[0]; }
  float* getVerticePtr() { return vertices_.empty() ? 0 : &vertices_[0]; }
  // pointers into the containers, like the reference (GeometryHandler.h:79-86)
  unsigned int* getTriangleAt(unsigned int idx) { return &indices_[(size_t)idx * 3]; }
  float* getVertexAt(unsigned int idx) { return &vertices_[(size_t)idx * 3]; }
  void setVertexAt(unsigned int i, float x, float y, float z) { float* v = getVertexAt(i); v[0] = x; v[1] = y; v[2] = z; }
  // value forms for code that wants bounds checks
  nv::Vec3f vertex(unsigned int i) const { return nv::Vec3f(vertices_.at((size_t)3 * i), vertices_.at((size_t)3 * i + 1), vertices_.at((size_t)3 * i + 2)); }
  nv::Vec3ui triangle(unsigned int t) const { return nv::Vec3ui(indices_.at((size_t)3 * t), indices_.at((size_t)3 * t + 1), indices_.at((size_t)3 * t + 2)); }
  nv::Vec3f getBoundingBox() const { return bb_max_ - bb_min_; }
  nv::Vec3f getBoundingBoxMin() const { return bb_min_; }
  nv::Vec3f getBoundingBoxMax() const { return bb_max_; }
  nv::Vec3f getGeometryOffset() const { return offset_; }
  unsigned int getNumberOfLongEdgeNodes(float dx) const { return (unsigned int)(longest_edge_ / dx + 0.5f); }
  // rotations about the origin, as written in the reference (GeometryHandler.cpp:123-146; the elevation form writes
  // x and y from z and x); the bounding box is not recomputed there either
  void rotateGeometryAzimuth(float angle) {
    const float rad = angle / 180.f * nv::PI_F;
    for (unsigned int i = 0; i < getNumberOfVertices(); i++) {
      float* v = getVertexAt(i);
      const float x = v[0], y = v[1];
      v[0] = x * std::cos(rad) - y * std::sin(rad);
      v[1] = x * std::sin(rad) + y * std::cos(rad);
    }
  }
  void rotateGeometryElevation(float angle) {
    const float rad = angle / 180.f * nv::PI_F;
    for (unsigned int i = 0; i < getNumberOfVertices(); i++) {
      float* v = getVertexAt(i);
      const float x = v[0], z = v[2];
      v[0] = z * std::cos(rad) - x * std::sin(rad);
      v[1] = z * std::sin(rad) + x * std::cos(rad);
    }
  }
  // half the length of the cross product of two edges (reference GeometryHandler.cpp:148-166)
  float getSurfaceAreaAt(unsigned int t) const {
    const nv::Vec3ui tri = triangle(t);
    const nv::Vec3f a = vertex(tri.x), u = a - vertex(tri.y), v = a - vertex(tri.z);
    const nv::Vec3f n(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
    return 0.5f * std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
  }
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
  float getTotalSurfaceArea() const { float s = 0.f; for (unsigned int i = 0; i < getNumberOfTriangles(); i++) s += getSurfaceAreaAt(i); return s; }

 private:
  std::vector<unsigned int> indices_;
  std::vector<float> vertices_;
  std::map<std::string, std::vector<int> > layers_;
  nv::Vec3f bb_min_, bb_max_, offset_;
  float longest_edge_;
};
