// geomMath.h -- the few nv:: types the hot-path API uses (reference src/math/geomMath.h:114-141,
// src/math/Vector.h): 3-component vectors and ROUND.
#pragma once
#include <cmath>
namespace nv {
template <typename T> struct Vec3 {
  T x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  void set(T x_, T y_, T z_) { x = x_; y = y_; z = z_; }
  Vec3 operator+(const Vec3& o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
  Vec3 operator-(const Vec3& o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
  Vec3 operator*(T s) const { return Vec3(x * s, y * s, z * s); }
  bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }
};
typedef Vec3<float> Vec3f;
typedef Vec3<int> Vec3i;
typedef Vec3<unsigned int> Vec3ui;
const float PI_F = 3.14159265358979323846f;
// reference geomMath.h:126-128: round half away from zero for positive, floor(x + 0.5)
inline float ROUND(float x) { return std::floor(x + 0.5f); }
}  // namespace nv
#ifndef PI
#define PI 3.14159265358979323846
#endif
