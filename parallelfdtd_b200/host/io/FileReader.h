// FileReader.h -- the two text readers of the reference's front ends (reference src/io/FileReader.{h,cpp}):
//   readVTK    legacy ASCII VTK POLYDATA with triangles: "POINTS n <type>" followed by 3n coordinates in INCHES
//              (converted with 0.0254 and snapped to a grid of `truncate_vertices_to` metres, FileReader.cpp:74-79),
//              "POLYGONS n m" followed by n records "3 i j k"; the result initialises a GeometryHandler.
//   readFloat  whitespace-separated floats (grid impulse responses, input signals).
// Token-based: keywords are found anywhere in the file (the reference scans the first 20 lines only), a polygon that
// is not a triangle or an index beyond the point list makes readVTK return false instead of producing a broken mesh.
// Header-only.
#pragma once
#include <fstream>
#include <string>
#include <vector>

#include "../base/GeometryHandler.h"
#include "../logger.h"
#include "../math/geomMath.h"

class FileReader {
 public:
  FileReader() : counter(0) {}
  int counter;                                     // values read by the last call (reference FileReader.h:38)

  void printCount(std::string type, bool reset) {
    log_msg<LOG_INFO>(L"FileReader %d %s values read") % counter % type;
    if (reset) counter = 0;
  }

  bool readVTK(GeometryHandler* gh, std::string fp, float truncate_vertices_to = 0.01f) {
    std::ifstream in(fp.c_str());
    if (!in.good()) {
      log_msg<LOG_ERROR>(L"FileReader::readVTK - invalid geometry file %s") % fp;
      return false;
    }
    std::vector<float> vertices;
    std::vector<unsigned int> indices;
    bool have_points = false, have_polygons = false;
    std::string tok;
    counter = 0;
    while (in >> tok) {
      if (tok == "POINTS") {
        unsigned long n = 0;
        std::string type;
        if (!(in >> n >> type)) return fail_(fp, "POINTS header");
        vertices.assign((size_t)n * 3, 0.f);
        for (size_t i = 0; i < vertices.size(); i++) {
          float v;
          if (!(in >> v)) return fail_(fp, "point coordinates");
          vertices[i] = nv::ROUND((v * 0.0254f) / truncate_vertices_to) * truncate_vertices_to;
        }
        counter += (int)vertices.size();
        have_points = true;
      } else if (tok == "POLYGONS") {
        unsigned long n = 0, total = 0;
        if (!(in >> n >> total)) return fail_(fp, "POLYGONS header");
        indices.assign((size_t)n * 3, 0u);
        for (size_t t = 0; t < n; t++) {
          unsigned int k = 0, a = 0, b = 0, c = 0;
          if (!(in >> k >> a >> b >> c)) return fail_(fp, "polygon record");
          if (k != 3) return fail_(fp, "a polygon that is not a triangle");
          indices[3 * t] = a; indices[3 * t + 1] = b; indices[3 * t + 2] = c;
        }
        counter += (int)indices.size();
        have_polygons = true;
      }
    }
    if (!have_points || !have_polygons) return fail_(fp, "no POINTS / POLYGONS section");
    for (size_t i = 0; i < indices.size(); i++)
      if (indices[i] >= vertices.size() / 3) return fail_(fp, "a polygon index beyond the point list");
    gh->initialize(indices, vertices);
    return true;
  }

  std::vector<float> readFloat(std::string fp) {
    std::vector<float> ret;
    std::ifstream in(fp.c_str());
    if (!in.good()) {
      log_msg<LOG_DEBUG>(L"FileReader::readFloat - invalid file %s") % fp;
      return ret;
    }
    float v;
    while (in >> v) ret.push_back(v);
    counter = (int)ret.size();
    return ret;
  }

 private:
  bool fail_(const std::string& fp, const char* what) {
    log_msg<LOG_ERROR>(L"FileReader::readVTK - %s: %s") % fp % what;
    return false;
  }
};
