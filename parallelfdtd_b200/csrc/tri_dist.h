// tri_dist.h -- squared distance from a point to a triangle (closest-point regions: the three vertices, the three
// edges, the face), written over an `Ops` policy so that the device kernel (round-to-nearest intrinsics, no FMA
// contraction) and the host restatement of the voxeliser (tests/cpp/voxelize_ref.h, plain float operators) evaluate
// the same operations in the same order and agree bit for bit.  Used to give a boundary voxel the material of the
// NEAREST TRIANGLE (not of the nearest centroid: on a coarse room mesh a floor voxel next to a wall is closer to
// the wall triangle's centroid than to the floor triangle's).
#pragma once

#ifdef __CUDACC__
#define PFDTD_HD __host__ __device__ __forceinline__
#else
#define PFDTD_HD inline
#endif

namespace pfdtd_geom {

template <class O>
PFDTD_HD float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return O::add(O::add(O::mul(ax, bx), O::mul(ay, by)), O::mul(az, bz));
}

template <class O>
PFDTD_HD float point_triangle_dist2(float px, float py, float pz, float ax, float ay, float az, float bx, float by, float bz, float cx,
                                    float cy, float cz) {
  const float abx = O::sub(bx, ax), aby = O::sub(by, ay), abz = O::sub(bz, az);
  const float acx = O::sub(cx, ax), acy = O::sub(cy, ay), acz = O::sub(cz, az);
  const float apx = O::sub(px, ax), apy = O::sub(py, ay), apz = O::sub(pz, az);
  float qx, qy, qz;                                   // closest point
  const float d1 = dot3<O>(abx, aby, abz, apx, apy, apz), d2 = dot3<O>(acx, acy, acz, apx, apy, apz);
  const float bpx = O::sub(px, bx), bpy = O::sub(py, by), bpz = O::sub(pz, bz);
  const float d3 = dot3<O>(abx, aby, abz, bpx, bpy, bpz), d4 = dot3<O>(acx, acy, acz, bpx, bpy, bpz);
  const float cpx = O::sub(px, cx), cpy = O::sub(py, cy), cpz = O::sub(pz, cz);
  const float d5 = dot3<O>(abx, aby, abz, cpx, cpy, cpz), d6 = dot3<O>(acx, acy, acz, cpx, cpy, cpz);
  const float vc = O::sub(O::mul(d1, d4), O::mul(d3, d2));
  const float vb = O::sub(O::mul(d5, d2), O::mul(d1, d6));
  const float va = O::sub(O::mul(d3, d6), O::mul(d5, d4));
  if (d1 <= 0.f && d2 <= 0.f) { qx = ax; qy = ay; qz = az; }                                   // vertex a
  else if (d3 >= 0.f && d4 <= d3) { qx = bx; qy = by; qz = bz; }                               // vertex b
  else if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {                                              // edge ab
    const float v = O::div(d1, O::sub(d1, d3));
    qx = O::add(ax, O::mul(v, abx)); qy = O::add(ay, O::mul(v, aby)); qz = O::add(az, O::mul(v, abz));
  } else if (d6 >= 0.f && d5 <= d6) { qx = cx; qy = cy; qz = cz; }                             // vertex c
  else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {                                              // edge ac
    const float w = O::div(d2, O::sub(d2, d6));
    qx = O::add(ax, O::mul(w, acx)); qy = O::add(ay, O::mul(w, acy)); qz = O::add(az, O::mul(w, acz));
  } else if (va <= 0.f && O::sub(d4, d3) >= 0.f && O::sub(d5, d6) >= 0.f) {                    // edge bc
    const float w = O::div(O::sub(d4, d3), O::add(O::sub(d4, d3), O::sub(d5, d6)));
    qx = O::add(bx, O::mul(w, O::sub(cx, bx))); qy = O::add(by, O::mul(w, O::sub(cy, by))); qz = O::add(bz, O::mul(w, O::sub(cz, bz)));
  } else {                                                                                      // face
    const float den = O::add(O::add(va, vb), vc);
    const float v = O::div(vb, den), w = O::div(vc, den);
    qx = O::add(O::add(ax, O::mul(abx, v)), O::mul(acx, w));
    qy = O::add(O::add(ay, O::mul(aby, v)), O::mul(acy, w));
    qz = O::add(O::add(az, O::mul(abz, v)), O::mul(acz, w));
  }
  const float ex = O::sub(px, qx), ey = O::sub(py, qy), ez = O::sub(pz, qz);
  return dot3<O>(ex, ey, ez, ex, ey, ez);
}

struct PlainOps {   // host: one rounding per operator (build without FMA contraction)
  static float add(float a, float b) { return a + b; }
  static float sub(float a, float b) { return a - b; }
  static float mul(float a, float b) { return a * b; }
  static float div(float a, float b) { return a / b; }
};

}  // namespace pfdtd_geom
