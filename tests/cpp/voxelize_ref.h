// voxelize_ref.h -- TEST INFRASTRUCTURE: host restatement of the device voxeliser
// (parallelfdtd_b200/csrc/voxelize_kernels.cu, pfdtd_voxelize*), operation by operation, used by tests/cpp/host_tests.cpp
// to check it bit for bit.  Solid voxelisation of a closed triangle mesh into the voxelizer-style node
// volumes the solver consumes: `bid` (0 solid, 27 air, 1..26 boundary codes, SURVEY Appendix B /
// reference src/kernels/cudaMesh.cu:372-476) and a material index per boundary voxel.
//
// The reference delegates this to the third-party hakarlss/Voxelizer (not vendored, no pinned
// version), so there is nothing to be bit-compatible with: PARITY UNPINNED.  Convention chosen to
// agree with the reference's source/receiver indexing (voxel = ROUND(p/dx) + 1,
// SimulationParameters.cpp:200-208): voxel (i,j,k) samples the point ((i-1)dx, (j-1)dx, (k-1)dx).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "base/GeometryHandler.h"
#include "../../parallelfdtd_b200/csrc/tri_dist.h"

namespace pfdtd_host {

struct VoxelVolumes { std::vector<unsigned char> bid, mat; unsigned int vx, vy, vz; };

// air-neighbour bit set -> bid.  bits: L=1 (x-1), R=2 (x+1), IN=4 (y-1), OUT=8 (y+1), D=16 (z-1), U=32 (z+1)
inline unsigned char bid_from_mask(unsigned int m) {
  static unsigned char lut[64]; static bool init = false;
  if (!init) {
    for (int i = 0; i < 64; i++) lut[i] = 0;
    const int L = 1, R = 2, I = 4, O = 8, D = 16, U = 32;
    const int sets[28] = {0, D|L|I, D|R|I, D|L|O, D|R|O, U|L|I, U|R|I, U|L|O, U|R|O,
                          D|L|R|I, D|L|R|O, D|L|I|O, D|R|I|O, U|L|R|I, U|L|R|O, U|L|I|O, U|R|I|O,
                          U|D|L|I, U|D|R|I, U|D|L|O, U|D|R|O,
                          L|R|I|O|D, L|R|O|D|U, L|R|I|D|U, R|I|O|D|U, L|I|O|D|U, L|R|I|O|U, L|R|I|O|D|U};
    for (int b = 1; b < 28; b++) lut[sets[b]] = (unsigned char)b;
    init = true;
  }
  return lut[m & 63];
}

// Ray-parity fill along x for every (y,z) grid line, then 6-neighbour classification.  Air voxels whose
// neighbour set has no bid code (thin features) are turned solid until the volume is consistent.
inline VoxelVolumes voxelize(const GeometryHandler& g, float dx, const unsigned char* tri_material /*may be null*/) {
  VoxelVolumes v;
  nv::Vec3f mx = g.getBoundingBoxMax();
  v.vx = (unsigned int)std::ceil(mx.x / dx) + 3; v.vy = (unsigned int)std::ceil(mx.y / dx) + 3; v.vz = (unsigned int)std::ceil(mx.z / dx) + 3;
  const size_t n = (size_t)v.vx * v.vy * v.vz;
  // A lattice point ON the surface counts as solid: the point is tested with its (y,z) nudged by +-eps in
  // all four combinations (rays never graze an edge or lie in a wall plane) and must be inside, by more
  // than eps along x as well, for every one of them.
  std::vector<unsigned char> in(n, 1);
  const unsigned int nt = g.getNumberOfTriangles();
  const float eps = 1e-4f * dx;
  std::vector<float> hits;
  std::vector<unsigned char> row(v.vx);
  for (int sy = -1; sy <= 1; sy += 2)
    for (int sz = -1; sz <= 1; sz += 2)
      for (unsigned int k = 0; k < v.vz; k++)
        for (unsigned int j = 0; j < v.vy; j++) {
          const float py = ((float)j - 1.f) * dx + sy * eps, pz = ((float)k - 1.f) * dx + sz * 2 * eps;
          hits.clear();
          for (unsigned int t = 0; t < nt; t++) {
            nv::Vec3ui tr = g.triangle(t);
            nv::Vec3f a = g.vertex(tr.x), b = g.vertex(tr.y), c = g.vertex(tr.z);
            // barycentric test of (py,pz) in the triangle projected on the yz plane
            const float d = (b.y - a.y) * (c.z - a.z) - (c.y - a.y) * (b.z - a.z);
            if (std::fabs(d) < 1e-20f) continue;
            const float u = ((py - a.y) * (c.z - a.z) - (c.y - a.y) * (pz - a.z)) / d;
            const float w = ((b.y - a.y) * (pz - a.z) - (py - a.y) * (b.z - a.z)) / d;
            if (u < 0 || w < 0 || u + w > 1) continue;
            hits.push_back(a.x + u * (b.x - a.x) + w * (c.x - a.x));
          }
          std::sort(hits.begin(), hits.end());
          std::fill(row.begin(), row.end(), 0);
          for (size_t h = 0; h + 1 < hits.size(); h += 2)
            for (unsigned int i = 0; i < v.vx; i++) {
              const float px = ((float)i - 1.f) * dx;
              if (px - 3 * eps > hits[h] && px + 3 * eps < hits[h + 1]) row[i] = 1;
            }
          for (unsigned int i = 0; i < v.vx; i++) in[((size_t)k * v.vy + j) * v.vx + i] &= row[i];
        }
  v.bid.assign(n, 0);
  auto at = [&](int i, int j, int k) -> int {
    if (i < 0 || j < 0 || k < 0 || i >= (int)v.vx || j >= (int)v.vy || k >= (int)v.vz) return 0;
    return in[((size_t)k * v.vy + j) * v.vx + i]; };
  for (bool changed = true; changed;) {
    changed = false;
    for (int k = 0; k < (int)v.vz; k++) for (int j = 0; j < (int)v.vy; j++) for (int i = 0; i < (int)v.vx; i++) {
      const size_t e = ((size_t)k * v.vy + j) * v.vx + i;
      if (!in[e]) { v.bid[e] = 0; continue; }
      const unsigned int m = at(i - 1, j, k) | at(i + 1, j, k) << 1 | at(i, j - 1, k) << 2 | at(i, j + 1, k) << 3 | at(i, j, k - 1) << 4 | at(i, j, k + 1) << 5;
      const unsigned char b = bid_from_mask(m);
      if (b == 0) { in[e] = 0; v.bid[e] = 0; changed = true; } else v.bid[e] = b;
    }
  }
  // material of a boundary voxel = material of the nearest triangle (point-to-triangle distance, csrc/tri_dist.h with
  // plain float operators; first of equally near triangles)
  v.mat.assign(n, 0);
  if (tri_material && nt) {
    for (unsigned int k = 0; k < v.vz; k++) for (unsigned int j = 0; j < v.vy; j++) for (unsigned int i = 0; i < v.vx; i++) {
      const size_t e = ((size_t)k * v.vy + j) * v.vx + i;
      if (v.bid[e] == 0 || v.bid[e] == 27) continue;
      const nv::Vec3f p(((float)i - 1.f) * dx, ((float)j - 1.f) * dx, ((float)k - 1.f) * dx);
      float best = 1e30f; unsigned int bt = 0;
      for (unsigned int t = 0; t < nt; t++) {
        nv::Vec3ui tr = g.triangle(t);
        const nv::Vec3f a = g.vertex(tr.x), b = g.vertex(tr.y), c = g.vertex(tr.z);
        const float dd = pfdtd_geom::point_triangle_dist2<pfdtd_geom::PlainOps>(p.x, p.y, p.z, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z);
        if (dd < best) { best = dd; bt = t; }
      }
      v.mat[e] = tri_material[bt];
    }
  }
  return v;
}

}  // namespace pfdtd_host
