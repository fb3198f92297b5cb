// host_tests.cpp -- the reference's own unit tests for the host classes on the hot path, re-expressed
// without Boost.Test against parallelfdtd_b200/host/ (reference tests/SimulationParametersTest.cpp,
// tests/MaterialHandlerTest.cpp, tests/SrcRecTest.cpp, tests/CudaMeshTest.cpp).
//   host_tests cpu   -> everything that needs no device
//   host_tests gpu   -> CudaMesh / launchFDTD3d / App tests (needs a CUDA device)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "App.h"
#include "kernels/kernels3d.h"
#include "voxelize_ref.h"
#include <fstream>

static int g_fail = 0, g_checks = 0;
#define CHECK(c) do { g_checks++; if (!(c)) { g_fail++; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)
#define CHECK_EQ(a, b) do { g_checks++; if (!((a) == (b))) { g_fail++; std::printf("FAIL %s:%d  %s == %s\n", __FILE__, __LINE__, #a, #b); } } while (0)
#define CHECK_THROW(expr, ex) do { g_checks++; bool t_ = false; try { expr; } catch (const ex&) { t_ = true; } catch (...) {} if (!t_) { g_fail++; std::printf("FAIL %s:%d  %s should throw %s\n", __FILE__, __LINE__, #expr, #ex); } } while (0)

// ---- SimulationParametersTest.cpp -------------------------------------------------------------------------
static void test_simulation_parameters() {
  { SimulationParameters sp;                                            // :8-14
    CHECK_EQ(sp.getC(), 344.f);
    CHECK_EQ(sp.getLambda(), (double)1 / std::sqrt((double)3)); }
  { SimulationParameters sp; sp.setSpatialFs(2000);                     // :16-22
    float reference = sp.getC() / sp.getSpatialFs() / (1 / sqrtf(3.f));
    CHECK(std::fabs(sp.getDx() - reference) <= 1e-7f * reference); }
  { SimulationParameters sp;                                            // :24-43
    sp.addSource(Source(3.f, 4.f, 5.f)); sp.addSource(7.f, 8.f, 9.f);
    CHECK_EQ(sp.getSource(0).getP().x, 3.f); CHECK_EQ(sp.getSource(1).getP().x, 7.f);
    CHECK_THROW(sp.getSource(2), std::out_of_range); CHECK_EQ(sp.getNumSources(), 2u);
    sp.addReceiver(Receiver(0.f, 4.f, 5.f)); sp.addReceiver(1.f, 5.f, 3.f);
    CHECK_EQ(sp.getReceiver(0).getP().x, 0.f); CHECK_EQ(sp.getReceiver(1).getP().x, 1.f);
    CHECK_THROW(sp.getReceiver(2), std::out_of_range); CHECK_EQ(sp.getNumReceivers(), 2u); }
  { SimulationParameters sp;                                            // :45-61
    sp.addSource(3.f, 4.f, 5.f); sp.addSource(7.f, 4.f, 3.f); sp.addReceiver(1.f, 2.f, 3.f); sp.addReceiver(6.f, 7.f, 8.f);
    CHECK_THROW(sp.updateSourceAt(2, Source()), std::out_of_range);
    sp.updateSourceAt(1, Source(1.f, 1.f, 1.f, SRC_SOFT));
    CHECK_EQ(sp.getSource(1).getP().x, 1.f); CHECK_EQ(sp.getSource(1).getSourceType(), SRC_SOFT);
    sp.updateReceiverAt(0, Receiver(0.f, 0.f, 0.f)); CHECK_EQ(sp.getReceiver(0).getP().x, 0.f); }
  { SimulationParameters sp;                                            // :63-78
    sp.addSource(3.f, 4.f, 5.f); sp.addSource(7.f, 4.f, 3.f); sp.addReceiver(1.f, 2.f, 3.f); sp.addReceiver(6.f, 7.f, 8.f);
    CHECK_THROW(sp.removeSource(2), std::out_of_range); sp.removeSource(1); CHECK_EQ(sp.getNumSources(), 1u);
    CHECK_THROW(sp.removeReceiver(2), std::out_of_range); sp.removeReceiver(1); CHECK_EQ(sp.getNumReceivers(), 1u); }
  { SimulationParameters sp;                                            // :80-97
    float* p = sp.getParameterPtr(); double* pd = sp.getParameterPtrDouble();
    double lambda = 1 / std::sqrt(double(3));
    CHECK_EQ(p[0], (float)lambda); CHECK_EQ(p[1], (float)(lambda * lambda)); CHECK_EQ(p[2], 1.f / 3.f); CHECK_EQ(p[3], 0.f);
    CHECK_EQ(pd[0], lambda); CHECK_EQ(pd[1], lambda * lambda); CHECK_EQ(pd[2], (double)1 / (double)3); CHECK_EQ(pd[3], 0.0); }
  { SimulationParameters sp; sp.setNumSteps(1000);                      // :99-124
    std::vector<float> d1(100), d2(200);
    for (unsigned i = 0; i < 200; i++) { if (i < 100) d1[i] = (float)i; d2[i] = -1.f * (float)i; }
    sp.addInputData(&d1[0], 100); sp.addInputData(&d2[0], 200);
    CHECK_EQ(sp.getInputDataSample(0, 0), 0.f); CHECK_EQ(sp.getInputDataSample(1, 20), -20.f);
    CHECK_EQ(sp.getInputDataSample(0, 99), 99.f); CHECK_EQ(sp.getInputDataSample(0, 1000), 0.f);
    CHECK_THROW(sp.getInputDataSample(2, 1000), std::out_of_range); }
  { SimulationParameters sp;                                            // :126-152
    sp.addSource(Source(1.f, 1.f, 1.f, SRC_HARD, IMPULSE, 0)); sp.addSource(Source(1.f, 1.f, 1.f, SRC_HARD, DATA, 0));
    std::vector<float> in; for (unsigned i = 0; i < 200; i++) in.push_back((float)i);
    sp.addInputData(&in[0], (unsigned)in.size() - 1);
    CHECK_EQ(sp.getSourceSample(0, 0), 0.f); CHECK_EQ(sp.getSourceSample(0, 1), 1.f); CHECK_EQ(sp.getSourceSample(0, 2), 0.f);
    CHECK_EQ(sp.getSourceSample(1, 0), 0.f); CHECK_EQ(sp.getSourceSample(1, 1), 1.f); CHECK_EQ(sp.getSourceSample(1, 100), 100.f);
    CHECK_EQ(sp.getSourceSample(1, 300), 0.f); CHECK_THROW(sp.getSourceSample(2, 1000), std::out_of_range); }
  { SimulationParameters sp; sp.setNumSteps(200);                       // :154-185
    sp.addSource(Source(1.f, 1.f, 1.f, SRC_HARD, IMPULSE, 0)); sp.addSource(Source(1.f, 1.f, 1.f, SRC_HARD, DATA, 0));
    std::vector<float> in; for (unsigned i = 0; i < 200; i++) in.push_back((float)i);
    sp.addInputData(&in[0], (unsigned)in.size() - 1);
    float* s1 = sp.getSourceVectorAt(0); float* s2 = sp.getSourceVectorAt(1);
    for (unsigned i = 0; i < sp.getNumSteps(); i++) { CHECK_EQ(s1[i], sp.getSourceSample(0, i)); CHECK_EQ(s2[i], sp.getSourceSample(1, i)); }
    CHECK_EQ(s1[1], 1.f); CHECK_EQ(s2[100], 100.f); }
  { SimulationParameters sp;                                            // :187-207, grid IR values quoted at :191-192
    std::vector<float> ir(4, 0.f); ir[2] = -0.333333343f; sp.setGridIr(ir);
    CHECK_EQ(sp.getGridIrDataSample(0), 0.f); CHECK_EQ(sp.getGridIrDataSample(2), -0.333333343f); CHECK_EQ(sp.getGridIrDataSample(99), 0.f);
    sp.addSource(Source(1.f, 1.f, 1.f, SRC_TRANSPARENT, IMPULSE, 0)); sp.setNumSteps(100);
    CHECK_EQ(sp.getSourceSample(0, 0), 0.f); CHECK_EQ(sp.getSourceSample(0, 1), 1.f); CHECK_EQ(sp.getSourceSample(0, 3), 1.f / 3.f);
    // the table form equals the per-step form, float and double, every source type
    sp.addSource(Source(1.f, 1.f, 1.f, SRC_TRANSPARENT, GAUSSIAN, 0)); sp.addSource(Source(1.f, 1.f, 1.f, SRC_SOFT, SINE, 0));
    std::vector<float> tf; std::vector<double> td; sp.fillSourceTable(tf, 100); sp.fillSourceTableDouble(td, 100);
    bool same = true;
    for (unsigned s = 0; s < 3; s++) for (unsigned n = 0; n < 100; n++) {
      same &= tf[s * 100 + n] == sp.getSourceSample(s, n); same &= td[s * 100 + n] == sp.getSourceSampleDouble(s, n); }
    CHECK(same); }
  { SimulationParameters sp; sp.setUpdateType(SRL);                     // setUpdateType resets lambda (SimulationParameters.cpp:123-137)
    CHECK_EQ(sp.getLambda(), std::sqrt((double)1 / 3)); CHECK_EQ(sp.getParameterPtrDouble()[1], (double)1 / 3); }
  { SimulationParameters sp; sp.setSpatialFs(7000);                     // element coordinates: ROUND(p/dx) + 1
    float dx = sp.getC() / ((float)7000 * (float)sp.getLambda());
    sp.addSource(Source(10 * dx, 3.4f * dx, 3.6f * dx)); sp.addReceiver(Receiver(2 * dx, 0.f, dx));
    nv::Vec3i s = sp.getSourceElementCoordinates(0), r = sp.getReceiverElementCoordinates(0);
    CHECK(s == nv::Vec3i(11, 4, 5)); CHECK(r == nv::Vec3i(3, 1, 2));
    sp.setAddPaddingToElementIdx(false); CHECK(sp.getSourceElementCoordinates(0) == nv::Vec3i(10, 3, 4)); }
}

// ---- SrcRecTest.cpp ---------------------------------------------------------------------------------------------
static void test_parameter_generation() {
  SimulationParameters sp;
  unsigned long long g = sp.generation();
  sp.addSource(Source(1.f, 1.f, 1.f)); CHECK(sp.generation() > g); g = sp.generation();
  sp.updateSourceAt(0, Source(2.f, 1.f, 1.f)); CHECK(sp.generation() > g); g = sp.generation();
  sp.addReceiver(1.f, 2.f, 3.f); CHECK(sp.generation() > g); g = sp.generation();
  sp.setNumSteps(7); CHECK(sp.generation() > g); g = sp.generation();
  sp.setSpatialFs(8000); CHECK(sp.generation() > g); g = sp.generation();
  sp.addInputData(std::vector<float>(3, 1.f)); CHECK(sp.generation() > g); g = sp.generation();
  sp.removeSource(0); CHECK(sp.generation() > g); g = sp.generation();
  (void)sp.getNumSources(); (void)sp.getDx(); CHECK_EQ(sp.generation(), g);
}

static void test_srcrec() {
  Source s; CHECK_EQ(s.getSourceType(), SRC_HARD); CHECK_EQ(s.getInputType(), IMPULSE); CHECK_EQ(s.getP().x, 0.f);
  Source s2(1.f, 2.f, 3.f, SRC_SOFT, DATA, 4); CHECK_EQ(s2.getSourceType(), SRC_SOFT); CHECK_EQ(s2.getInputType(), DATA);
  CHECK_EQ(s2.getInputDataIdx(), 4u); CHECK_EQ(s2.getP().z, 3.f);
  Source s3(1.f, 2.f); CHECK_EQ(s3.getP().z, 0.f);
  s3.setSourceType(SRC_TRANSPARENT); s3.setInputType(SINE); s3.setGroup(3); CHECK_EQ(s3.getGroup(), 3u); CHECK_EQ(s3.getInputType(), SINE);
  Receiver r(4.f, 5.f, 6.f); CHECK_EQ(r.getP().y, 5.f);
}

// ---- MaterialHandlerTest.cpp ------------------------------------------------------------------------------------
static void test_material_handler() {
  { MaterialHandler mh; CHECK_EQ(mh.getNumberOfSurfaces(), 0u); }
  { MaterialHandler mh; std::vector<float> t;                            // :17-44
    for (unsigned i = 0; i < mh.getNumberOfCoefficients(); i++) t.push_back((float)i);
    mh.addSurfaceMaterial(t); mh.addSurfaceMaterial(t);
    CHECK_EQ(mh.getNumberOfUniqueMaterials(), 1u); CHECK_EQ(mh.getNumberOfSurfaces(), 2u);
    t.at(0) = 3.f; mh.addSurfaceMaterial(t); CHECK_EQ(mh.getNumberOfUniqueMaterials(), 2u); CHECK_EQ(mh.getNumberOfSurfaces(), 3u);
    t.clear(); for (unsigned i = 0; i < 10; i++) t.push_back((float)i);
    mh.addSurfaceMaterial(t); CHECK_EQ(mh.getNumberOfUniqueMaterials(), 3u); CHECK_EQ(mh.getNumberOfSurfaces(), 4u);
    CHECK_EQ(mh.getMaterialIdxAt(0), 0u); CHECK_EQ(mh.getMaterialIdxAt(1), 0u); CHECK_EQ(mh.getMaterialIdxAt(2), 1u); CHECK_EQ(mh.getMaterialIdxAt(3), 2u);   // :46-87
    CHECK_EQ(mh.getUniqueCoefAt(0, 0), 0.f); CHECK_EQ(mh.getUniqueCoefAt(0, 19), 19.f); CHECK_EQ(mh.getUniqueCoefAt(1, 0), 3.f); CHECK_EQ(mh.getUniqueCoefAt(2, 1), 1.f);
    CHECK_EQ(mh.getSurfaceCoefAt(0, 5), 5.f); CHECK_EQ(mh.getSurfaceCoefAt(2, 0), 3.f); CHECK_EQ(mh.getSurfaceCoefAt(3, 11), 0.f);
    CHECK_THROW(mh.getMaterialIdxAt(4), std::out_of_range); CHECK_THROW(mh.getUniqueCoefAt(4, 0), std::out_of_range);
    CHECK_THROW(mh.getUniqueCoefAt(1, 22), std::out_of_range); CHECK_THROW(mh.getSurfaceCoefAt(1, 22), std::out_of_range); }
  { MaterialHandler mh; std::vector<float> m(50);                        // :89-107
    for (unsigned i = 0; i < 50; i++) m[i] = (float)i;
    mh.addMaterials(&m[0], 5, 10);
    CHECK_EQ(mh.getNumberOfUniqueMaterials(), 5u); CHECK_EQ(mh.getUniqueCoefAt(1, 0), 10.f); CHECK_EQ(mh.getUniqueCoefAt(1, 14), 0.f); CHECK_EQ(mh.getUniqueCoefAt(4, 9), 49.f); }
  { MaterialHandler mh; std::vector<float> t;                            // :109-143 flat table layout m_ptr[mat*20 + k]
    for (unsigned i = 0; i < 10; i++) t.push_back(((float)i) / 10.f);
    mh.addSurfaceMaterial(t); t.at(0) = 0.7f; mh.addSurfaceMaterial(t);
    float* p = mh.getMaterialCoefficientPtr(); double* pd = mh.getMaterialCoefficientPtrDouble();
    CHECK_EQ(p[0], 0.f); CHECK_EQ(p[5], 0.5f); CHECK_EQ(p[15], 0.f); CHECK_EQ(p[20], 0.7f); CHECK_EQ(p[25], 0.5f);
    CHECK_EQ(pd[20], (double)0.7f); CHECK_EQ(pd[9], (double)0.9f);
    mh.coefsAreReflectances(); p = mh.getMaterialCoefficientPtr(); CHECK_EQ(p[20], reflection2Admitance(0.7f)); }
  { MaterialHandler mh; mh.setGlobalMaterial(4, 0.5f);                   // :145-159
    CHECK_EQ(mh.getNumberOfSurfaces(), 4u); CHECK_EQ(mh.getNumberOfUniqueMaterials(), 1u);
    mh.setMaterialIndexAt(2, 7); CHECK_EQ(mh.getMaterialIdxAt(2), 7u); mh.setMaterialIndexAt(9, 1); CHECK_EQ(mh.getMaterialIdxAt(3), 0u); }
  { MaterialHandler mh;                                                  // filter materials (addition): rows [b0..bN, a1..aN]
    CHECK_EQ(mh.getFilterOrder(), 0u);
    const float rows[15] = {0.10f, 0.02f, 0.01f, -0.5f, 0.1f,   0.10f, 0.02f, 0.01f, -0.5f, 0.1f,   0.30f, 0.f, 0.f, -0.2f, 0.f};
    mh.coefsAreReflectances();                                           // must not convert filter rows
    mh.addFilterMaterials(rows, 3, 2);
    CHECK_EQ(mh.getFilterOrder(), 2u); CHECK_EQ(mh.getNumberOfSurfaces(), 3u); CHECK_EQ(mh.getNumberOfUniqueMaterials(), 2u);
    CHECK_EQ(mh.getMaterialIdxAt(1), 0u); CHECK_EQ(mh.getMaterialIdxAt(2), 1u);
    float* p = mh.getMaterialCoefficientPtr(); double* pd = mh.getMaterialCoefficientPtrDouble();
    CHECK_EQ(p[0], 0.10f); CHECK_EQ(p[3], -0.5f); CHECK_EQ(p[4], 0.1f); CHECK_EQ(p[5], 0.f); CHECK_EQ(p[20], 0.30f); CHECK_EQ(p[23], -0.2f);
    CHECK_EQ(pd[3], (double)-0.5f);
    CHECK_THROW(mh.setFilterOrder(3), std::logic_error); CHECK_THROW(mh.setFilterOrder(5), std::out_of_range);
    CHECK_THROW(mh.addFilterMaterials(rows, 1, 0), std::out_of_range); }
  { MaterialHandler mh; std::vector<float> b(3, 0.1f), a(2, -0.3f);
    mh.setGlobalFilter(4, b, a);
    CHECK_EQ(mh.getFilterOrder(), 2u); CHECK_EQ(mh.getNumberOfSurfaces(), 4u); CHECK_EQ(mh.getNumberOfUniqueMaterials(), 1u);
    CHECK_EQ(mh.getUniqueCoefAt(0, 2), 0.1f); CHECK_EQ(mh.getUniqueCoefAt(0, 4), -0.3f); CHECK_EQ(mh.getUniqueCoefAt(0, 5), 0.f);
    a.push_back(0.f); CHECK_THROW(mh.setGlobalFilter(1, b, a), std::out_of_range); }
}

// ---- CudaMeshTest.cpp:182-218 (host only) ------------------------------------------------------------------------
// without a device the allocation helpers fail like every other device call of the reference: log + throw(-1)
static void test_device_helpers_without_a_device() {
  int ndev = 0; pfdtd_device_count(&ndev);
  if (ndev > 0) return;
  CHECK_THROW(valueToDevice<unsigned char>(8, (unsigned char)1, 0), int);
  CHECK_THROW(toDevice<float>(8, 0), int);
  CHECK_THROW(getCurrentDevice(), int);
}

static void box_mesh(float lx, float ly, float lz, std::vector<unsigned>& idx, std::vector<float>& v);
// ---- FileReaderTest.cpp / GeometryHandlerTest.cpp:80-92 (the VTK fixture is not in the reference tree: written here) ------------
static void test_file_reader() {
  std::vector<unsigned> idx; std::vector<float> v; box_mesh(1.f, 1.f, 1.f, idx, v);
  { std::ofstream f("box1m.vtk");
    f << "# vtk DataFile Version 3.0\nvtk output\nASCII\nDATASET POLYDATA\nPOINTS 8 float\n";
    for (size_t i = 0; i < v.size(); i += 3) f << v[i] / 0.0254f << " " << v[i + 1] / 0.0254f << " " << v[i + 2] / 0.0254f << "\n";   // inches
    f << "POLYGONS 12 48\n";
    for (size_t i = 0; i < idx.size(); i += 3) f << "3 " << idx[i] << " " << idx[i + 1] << " " << idx[i + 2] << "\n"; }
  FileReader fr; GeometryHandler gh;
  CHECK(fr.readVTK(&gh, "box1m.vtk", 0.1f));
  CHECK_EQ(gh.getNumberOfTriangles(), 12u); CHECK_EQ(gh.getNumberOfVertices(), 8u); CHECK_EQ(fr.counter, 24 + 36);
  CHECK_EQ(gh.getTotalSurfaceArea(), 6.f); CHECK_EQ(gh.getSurfaceAreaAt(0), 0.5f);          // 1 m box: 6 m^2 (GeometryHandlerTest.cpp:88-91)
  CHECK(gh.getBoundingBox() == nv::Vec3f(1.f, 1.f, 1.f));
  GeometryHandler none; CHECK(!fr.readVTK(&none, "no_such_file.vtk")); CHECK_EQ(none.getNumberOfTriangles(), 0u);
  { std::ofstream f("quad.vtk"); f << "DATASET POLYDATA\nPOINTS 4 float\n0 0 0 1 0 0 1 1 0 0 1 0\nPOLYGONS 1 5\n4 0 1 2 3\n"; }
  CHECK(!fr.readVTK(&none, "quad.vtk"));
  { std::ofstream f("badidx.vtk"); f << "DATASET POLYDATA\nPOINTS 3 float\n0 0 0 1 0 0 1 1 0\nPOLYGONS 1 4\n3 0 1 7\n"; }
  CHECK(!fr.readVTK(&none, "badidx.vtk"));
  { std::ofstream f("ir.txt"); f << "0\n0 -0.333333343\n 0.5\n"; }
  std::vector<float> ir = fr.readFloat("ir.txt");
  CHECK_EQ(ir.size(), (size_t)4); CHECK_EQ(ir[2], -0.333333343f); CHECK(fr.readFloat("no_such_file.txt").empty());
  SimulationParameters sp; sp.readGridIr("ir.txt"); CHECK_EQ(sp.getGridIrDataSample(2), -0.333333343f);
  FDTD::App app; CHECK_THROW(app.initializeGeometryFromFile("no_such_file.vtk"), int);
  app.initializeGeometryFromFile("box1m.vtk"); CHECK_EQ(app.m_geometry.getNumberOfTriangles(), 12u);
}

static void test_partition_indexing() {
  CudaMesh mesh;
  const int dim_z = 100, num_p = 13, ps = dim_z / num_p;
  std::vector<std::vector<unsigned int> > part = mesh.getPartitionIndexing(num_p, dim_z);
  CHECK_EQ(part.size(), (size_t)num_p);
  for (int i = 0; i < num_p; i++) {
    int s_inc = i == 0 ? 0 : 1, e_inc = i == num_p - 1 ? 0 : 1;
    int expect = ps + s_inc + e_inc + (i == num_p - 1 ? dim_z - ps * num_p : 0);
    CHECK_EQ((int)part[i].size(), expect);
    for (size_t j = 0; j < part[i].size(); j++) CHECK_EQ((int)part[i][j], i * ps - s_inc + (int)j);
  }
  part = mesh.getPartitionIndexing(1, dim_z);
  CHECK_EQ(part.size(), (size_t)1); CHECK_EQ(part[0].size(), (size_t)dim_z);
  for (int j = 0; j < dim_z; j++) CHECK_EQ((int)part[0][j], j);
}

// ---- geometry helpers ---------------------------------------------------------------------------------------------
static void box_mesh(float lx, float ly, float lz, std::vector<unsigned>& idx, std::vector<float>& v) {
  const float c[8][3] = {{0,0,0},{lx,0,0},{lx,ly,0},{0,ly,0},{0,0,lz},{lx,0,lz},{lx,ly,lz},{0,ly,lz}};
  v.clear(); for (int i = 0; i < 8; i++) for (int k = 0; k < 3; k++) v.push_back(c[i][k]);
  const unsigned t[12][3] = {{0,1,2},{0,2,3},{4,6,5},{4,7,6},{0,5,1},{0,4,5},{3,2,6},{3,6,7},{0,3,7},{0,7,4},{1,5,6},{1,6,2}};
  idx.clear(); for (int i = 0; i < 12; i++) for (int k = 0; k < 3; k++) idx.push_back(t[i][k]);
}

static void test_geometry_and_voxelizer() {
  std::vector<unsigned> idx; std::vector<float> v; box_mesh(1.f, 1.f, 1.f, idx, v);
  GeometryHandler g; g.initialize(idx, v);                               // GeometryHandlerTest.cpp:85-100: 1 m box, area 6
  CHECK_EQ(g.getNumberOfTriangles(), 12u); CHECK(std::fabs(g.getTotalSurfaceArea() - 6.f) < 1e-5f);
  CHECK(g.getBoundingBox() == nv::Vec3f(1.f, 1.f, 1.f));
  { GeometryHandler gp; gp.initialize(&idx[0], &v[0], (unsigned)idx.size(), (unsigned)v.size());   // pointer form: the last argument counts
    CHECK_EQ(gp.getNumberOfVertices(), 8u); CHECK_EQ(gp.getNumberOfTriangles(), 12u);                // floats (reference GeometryHandler.cpp:87-108)
    CHECK(gp.getBoundingBox() == nv::Vec3f(1.f, 1.f, 1.f));
    std::vector<unsigned> bad(idx); bad[0] = 8; GeometryHandler gb;                                  // index values are checked where used
    gb.initialize(&bad[0], &v[0], (unsigned)bad.size(), (unsigned)v.size());
    CHECK_THROW(gb.getSurfaceAreaAt(0), std::out_of_range); CHECK(std::fabs(gb.getSurfaceAreaAt(1) - 0.5f) < 1e-6f); }
  { // GeometryHandlerTest.cpp:10-84 (constructor / pointer_initialize / getters): 33 floats, 33 indices
    std::vector<float> vv; std::vector<unsigned> ii;
    for (unsigned i = 0; i < 33; i++) { vv.push_back((float)(i * 10)); ii.push_back(i); }
    GeometryHandler a; a.initialize(ii, vv);
    GeometryHandler b; b.initialize(&ii[0], &vv[0], 33, 33);
    for (GeometryHandler* gh : {&a, &b}) {
      CHECK_EQ(gh->getNumberOfIndices(), 33u); CHECK_EQ(gh->getNumberOfTriangles(), 11u); CHECK_EQ(gh->getNumberOfVertices(), 11u);
      unsigned int* t = gh->getTriangleAt(3); CHECK_EQ(*t, 9u); CHECK_EQ(*(t + 1), 10u);
      float* f = gh->getVertexAt(3); CHECK_EQ(*f, 90.f);
      // the bounding-box corner (0, 10, 20) is removed and kept as the geometry offset (GeometryHandler.cpp:63-75)
      CHECK(gh->getGeometryOffset() == nv::Vec3f(0.f, 10.f, 20.f)); CHECK_EQ(f[1], 90.f); CHECK_EQ(f[2], 90.f);
      CHECK(gh->getBoundingBox() == nv::Vec3f(300.f, 300.f, 300.f)); CHECK_EQ(gh->getNumberOfLongEdgeNodes(10.f), 30u);
    }
    a.setVertexAt(0, 1.f, 2.f, 3.f); CHECK_EQ(a.getVertexAt(0)[2], 3.f);
    a.setVertexAt(1, 1.f, 0.f, 5.f); a.rotateGeometryAzimuth(90.f);
    CHECK(std::fabs(a.getVertexAt(1)[0]) < 1e-6f && std::fabs(a.getVertexAt(1)[1] - 1.f) < 1e-6f && a.getVertexAt(1)[2] == 5.f); }
  { // a box away from the origin voxelises like the same box at the origin
    std::vector<float> vs(v); for (size_t i = 0; i < vs.size(); i += 3) { vs[i] += 4.f; vs[i + 1] -= 2.5f; vs[i + 2] += 0.25f; }
    GeometryHandler gs; gs.initialize(idx, vs);
    CHECK(gs.getGeometryOffset() == nv::Vec3f(4.f, -2.5f, 0.25f)); CHECK(std::fabs(gs.getTotalSurfaceArea() - 6.f) < 1e-5f);
    pfdtd_host::VoxelVolumes v0 = pfdtd_host::voxelize(g, 0.1f, 0), v1 = pfdtd_host::voxelize(gs, 0.1f, 0);
    CHECK_EQ(v0.vx, v1.vx); CHECK_EQ(v0.vz, v1.vz); CHECK(v0.bid == v1.bid); }
  { // GeometryHandler_insertLauyers (:103-147)
    GeometryHandler gl; gl.setLayerIndices(std::vector<int>(10, 0), "Layer0"); CHECK_EQ(gl.getNumberOfLayers(), 0u);
    gl.initialize(idx, v);
    std::vector<int> li(10, 0); li[0] = -10; li[2] = 10; gl.setLayerIndices(li, "Layer0"); CHECK_EQ(gl.getNumberOfLayers(), 0u);
    std::vector<int> h0, h1; for (int i = 0; i < 6; i++) { h0.push_back(i); h1.push_back(i + 6); }
    gl.setLayerIndices(h0, "Layer0"); gl.setLayerIndices(h1, "Layer1"); CHECK_EQ(gl.getNumberOfLayers(), 2u);
    CHECK(gl.getLayerNameAt(3) == ""); CHECK(gl.getLayerNameAt(0) == "Layer0"); CHECK(gl.getLayerNameAt(1) == "Layer1"); }
  const float dx = 0.1f;
  pfdtd_host::VoxelVolumes vol = pfdtd_host::voxelize(g, dx, 0);
  CHECK_EQ(vol.vx, 13u);
  // interior points (i-1)*dx strictly inside (0,1): i = 2..10 -> 9 per axis
  size_t air = 0, bnd = 0;
  for (size_t e = 0; e < vol.bid.size(); e++) { air += vol.bid[e] == 27; bnd += vol.bid[e] > 0 && vol.bid[e] < 27; }
  CHECK_EQ(air, (size_t)7 * 7 * 7); CHECK_EQ(air + bnd, (size_t)9 * 9 * 9);
  CHECK_EQ((int)vol.bid[((size_t)2 * vol.vy + 2) * vol.vx + 2], 8);     // corner voxel: Up, Right, Out are air
}

// ---- GPU: CudaMeshTest.cpp:220-469 re-expressed ------------------------------------------------------------------
// Coarse, flat room (10 x 10 x 3 m, two triangles per face, dx = 0.25): every face voxel must carry the material of the
// face it lies on -- the nearest TRIANGLE, not the nearest centroid (which gives half the floor to the walls).
static void test_voxelizer_materials_follow_the_nearest_triangle() {
  std::vector<unsigned> idx; std::vector<float> v; box_mesh(10.f, 10.f, 3.f, idx, v);
  GeometryHandler g; g.initialize(&idx[0], &v[0], (unsigned)idx.size(), (unsigned)v.size());
  const unsigned nt = g.getNumberOfTriangles();
  std::vector<unsigned char> tm(nt);
  for (unsigned t = 0; t < nt; t++) {         // material = 1 + index of the axis-aligned face the triangle lies in
    nv::Vec3ui tr = g.triangle(t);
    const nv::Vec3f a = g.vertex(tr.x), b = g.vertex(tr.y), c = g.vertex(tr.z);
    unsigned char m = 0;
    if (a.z == b.z && b.z == c.z) m = a.z == 0.f ? 1 : 2;
    else if (a.y == b.y && b.y == c.y) m = a.y == 0.f ? 3 : 4;
    else if (a.x == b.x && b.x == c.x) m = a.x == 0.f ? 6 : 5;
    tm[t] = m;
  }
  pfdtd_host::VoxelVolumes vol = pfdtd_host::voxelize(g, 0.25f, &tm[0]);
  int face[28] = {0}; face[26] = 1; face[21] = 2; face[22] = 3; face[23] = 4; face[25] = 5; face[24] = 6;
  long total = 0, wrong = 0;
  for (size_t e = 0; e < vol.bid.size(); e++) {
    const int b = vol.bid[e];
    if (b >= 21 && b <= 26) { total++; if (vol.mat[e] != face[b]) wrong++; }
  }
  CHECK(total > 3000); CHECK_EQ(wrong, 0l);
}

static void shoebox_bid(unsigned vx, unsigned vy, unsigned vz, std::vector<unsigned char>& bid, std::vector<unsigned char>& mat) {
  bid.assign((size_t)vx * vy * vz, 0); mat.assign(bid.size(), 0);
  auto in = [&](int x, int y, int z) { return x >= 1 && y >= 1 && z >= 1 && x < (int)vx - 1 && y < (int)vy - 1 && z < (int)vz - 1; };
  for (int z = 0; z < (int)vz; z++) for (int y = 0; y < (int)vy; y++) for (int x = 0; x < (int)vx; x++) {
    if (!in(x, y, z)) continue;
    unsigned m = in(x - 1, y, z) | in(x + 1, y, z) << 1 | in(x, y - 1, z) << 2 | in(x, y + 1, z) << 3 | in(x, y, z - 1) << 4 | in(x, y, z + 1) << 5;
    bid[((size_t)z * vy + y) * vx + x] = pfdtd_host::bid_from_mask(m);
  }
}

// device voxeliser (pfdtd_voxelize) against the host restatement (voxelize_ref.h), bit for bit: box, off-grid box with
// per-face materials, and a room with a pillar (rows that cross the surface four times)
static void test_voxelizer_gpu() {
  struct Case { float lx, ly, lz, dx; bool pillar; };
  const Case cases[] = {{1.f, 1.f, 1.f, 0.1f, false}, {1.37f, 0.93f, 1.11f, 0.043f, false}, {2.f, 1.5f, 1.f, 0.05f, true}};
  for (const Case& c : cases) {
    std::vector<unsigned> idx; std::vector<float> v; box_mesh(c.lx, c.ly, c.lz, idx, v);
    if (c.pillar) {   // a box-shaped pillar from floor to ceiling inside the room: a second closed surface (parity fill makes it solid)
      std::vector<unsigned> pi; std::vector<float> pv; box_mesh(0.3f, 0.3f, c.lz, pi, pv);
      const unsigned base = (unsigned)(v.size() / 3);
      for (size_t i = 0; i < pv.size(); i += 3) { v.push_back(pv[i] + 0.8f); v.push_back(pv[i + 1] + 0.6f); v.push_back(pv[i + 2]); }
      for (size_t i = 0; i < pi.size(); i++) idx.push_back(pi[i] + base);
    }
    GeometryHandler g; g.initialize(idx, v);
    std::vector<unsigned char> tm(g.getNumberOfTriangles());
    for (size_t t = 0; t < tm.size(); t++) tm[t] = (unsigned char)(t / 2 % 7);
    pfdtd_host::VoxelVolumes ref = pfdtd_host::voxelize(g, c.dx, &tm[0]);
    unsigned vx = 0, vy = 0, vz = 0;
    CHECK_EQ(pfdtd_voxelize_dims(&v[0], (unsigned)(v.size() / 3), c.dx, &vx, &vy, &vz), 0);
    CHECK_EQ(vx, ref.vx); CHECK_EQ(vy, ref.vy); CHECK_EQ(vz, ref.vz);
    std::vector<unsigned char> bid((size_t)vx * vy * vz, 255), mat(bid.size(), 255);
    CHECK_EQ(pfdtd_voxelize(&v[0], (unsigned)(v.size() / 3), &idx[0], g.getNumberOfTriangles(), &tm[0], c.dx, &bid[0], &mat[0]), 0);
    CHECK(bid == ref.bid);
    CHECK(mat == ref.mat);
    size_t air = 0; for (size_t e = 0; e < bid.size(); e++) air += bid[e] == 27;
    CHECK(air > 0);
  }
}

static bool never(void) { return false; }
static void quiet(int, int, float) {}

static void test_cuda_mesh_gpu() {
  int ndev = 0; pfdtd_device_count(&ndev);
  if (ndev < 1) { std::printf("no CUDA device\n"); g_fail++; return; }
  SimulationParameters sp; MaterialHandler mh; mh.setGlobalMaterial(1, 0.5f);
  { // CudaMesh_set_get_utils (:220-258): raw 20^3 zero volume, block (32,4,2): element index and out-of-range
    std::vector<unsigned char> z((size_t)20 * 20 * 20, 0);
    CudaMesh mesh;
    mesh.setupMeshHost(&z[0], &z[0], mh.getNumberOfUniqueMaterials(), mh.getMaterialCoefficientPtr(), sp.getParameterPtr(),
                       make_uint3(20, 20, 20), make_uint3(32, 4, 2), 0);
    mesh.makePartition(1);
    CHECK_EQ(mesh.getDimX(), 32u); CHECK_EQ(mesh.getDimY(), 20u); CHECK_EQ(mesh.getDimZ(), 20u);
    int dev = 0, el = 0; mesh.getElementIdxAndDevice(10, 10, 10, &dev, &el);
    CHECK_EQ(dev, 0); CHECK_EQ(el, 10 * 20 * 32 + 10 * 32 + 10);
    mesh.getElementIdxAndDevice(10, 10, 30, &dev, &el); CHECK_EQ(dev, -1); CHECK_EQ(el, -1);
  }
  { // the reference's own hand-over (CudaMeshTest.cpp:226-233): device volumes made with valueToDevice, adopted by setupMesh
    const unsigned n = 20 * 20 * 20;
    unsigned char* d_pos = valueToDevice<unsigned char>(n, (unsigned char)0, 0);
    unsigned char* d_mat = valueToDevice<unsigned char>(n, (unsigned char)0, 0);
    CudaMesh mesh;
    mesh.setupMesh(d_pos, d_mat, mh.getNumberOfUniqueMaterials(), mh.getMaterialCoefficientPtr(), sp.getParameterPtr(), make_uint3(20, 20, 20),
                   make_uint3(32, 4, 2), 0);
    mesh.makePartition(1);
    CHECK_EQ(mesh.getDimX(), 32u); CHECK_EQ(mesh.getDimY(), 20u); CHECK_EQ(mesh.getDimZ(), 20u);
    CHECK_EQ(mesh.getNumberOfAirElements(), 0u);
  }
  { // cudaUtils.h helpers (reference cudaUtils.h:59-171)
    CHECK_EQ(getCurrentDevice(), 0);
    std::vector<float> h(1000); for (size_t i = 0; i < h.size(); i++) h[i] = (float)i * 0.5f;
    float* d = toDevice<float>(1000, &h[0], 0);
    float* back = fromDevice<float>(1000, d, 0);
    CHECK(std::memcmp(back, &h[0], 4000) == 0); std::free(back);
    CHECK_EQ(getSample<float>(7, d), 3.5f);
    resetData<float>(1000, d, 0); CHECK_EQ(getSample<float>(7, d), 0.f);
    h[3] = -2.f; copyHostToDevice<float>(1000, d, &h[0], 0);
    std::vector<float> h2(1000, 0.f); copyDeviceToHost<float>(1000, &h2[0], d, 0); CHECK(h2 == h);
    destroyMem(d, 0);
    double* dd = valueToDevice<double>(333, 1.25, 0); double* hd = fromDevice<double>(333, dd, 0);
    bool all = true; for (int i = 0; i < 333; i++) all &= hd[i] == 1.25; CHECK(all); std::free(hd); destroyMem(dd);
    float* df = valueToDevice<float>(777, -3.f, 0); float* hf = fromDevice<float>(777, df, 0);
    all = true; for (int i = 0; i < 777; i++) all &= hf[i] == -3.f; CHECK(all); std::free(hf); destroyMem(df);
    unsigned short* ds = valueToDevice<unsigned short>(5, (unsigned short)0xBEEF, 0); unsigned short* hs = fromDevice<unsigned short>(5, ds, 0);
    CHECK_EQ(hs[4], (unsigned short)0xBEEF); std::free(hs); destroyMem(ds);
    unsigned char* dz = toDevice<unsigned char>(64, 0); unsigned char* hz = fromDevice<unsigned char>(64, dz, 0);
    all = true; for (int i = 0; i < 64; i++) all &= hz[i] == 0; CHECK(all); std::free(hz); destroyMem(dz);
    CHECK_THROW(valueToDevice<float>(4, 1.f, 99), std::out_of_range);     // no such device
  }
  { // CudaMesh_get_set_multi (:260-303): 50^3, 5 partitions on one device
    std::vector<unsigned char> z((size_t)50 * 50 * 50, 0);
    CudaMesh mesh;
    mesh.setupMeshHost(&z[0], &z[0], 1, mh.getMaterialCoefficientPtr(), sp.getParameterPtr(), make_uint3(50, 50, 50), make_uint3(1, 1, 1), 0);
    mesh.makePartition(5, std::vector<unsigned>(5, 0));
    int dev = 0, el = 0; mesh.getElementIdxAndDevice(10, 10, 11, &dev, &el);
    CHECK_EQ(dev, 1); CHECK_EQ(el, 2 * 50 * 50 + 10 * 50 + 10);
  }
  for (int dbl = 0; dbl < 2; dbl++) {   // :306-469: set/add/get across slabs, halo duplication, switchHalos (float and double)
    std::vector<unsigned char> z((size_t)128 * 96 * 49, 0);
    CudaMesh mesh; mesh.setDouble(dbl != 0);
    mesh.setupMeshHost(&z[0], &z[0], 1, dbl ? (void*)mh.getMaterialCoefficientPtrDouble() : (void*)mh.getMaterialCoefficientPtr(),
                       dbl ? (void*)sp.getParameterPtrDouble() : (void*)sp.getParameterPtr(), make_uint3(128, 96, 49), make_uint3(32, 4, 1), 0);
    mesh.makePartition(2, std::vector<unsigned>(2, 0));
    CHECK_EQ(mesh.getPartitionSize(0), 25u); CHECK_EQ(mesh.getPartitionSize(1), 26u); CHECK_EQ(mesh.getFirstSliceIdx(1), 23u);
    mesh.setSample<float>(3.f, 10, 10, 10); CHECK_EQ(mesh.getSample<float>(10, 10, 10), 3.f);
    mesh.setSample<float>(2.f, 10, 10, 23);        // z = 23 is slab0[23] and slab1[0]
    CHECK_EQ(mesh.getSampleAt<float>(10, 10, 23, 0), 2.f); CHECK_EQ(mesh.getSampleAt<float>(10, 10, 0, 1), 2.f);
    mesh.setSample<float>(4.f, 10, 10, 24);        // z = 24 is slab0[24] and slab1[1]
    CHECK_EQ(mesh.getSampleAt<float>(10, 10, 24, 0), 4.f); CHECK_EQ(mesh.getSampleAt<float>(10, 10, 1, 1), 4.f);
    mesh.addSample<float>(5.f, 10, 10, 10);        // as written: addSample overwrites (cudaMesh.h:362-366)
    CHECK_EQ(mesh.getSample<float>(10, 10, 10), 5.f);
    mesh.setOption(PFDTD_OPT_SOFT_ACCUMULATE, 1);  // what the reference's own test expects (CudaMeshTest.cpp:311-314)
    mesh.addSample<float>(3.f, 10, 10, 10); CHECK_EQ(mesh.getSample<float>(10, 10, 10), 8.f);
    mesh.setSampleAt<float>(7.f, 5, 5, 23, 0); mesh.setSampleAt<float>(9.f, 5, 5, 1, 1);
    mesh.setSampleAt<float>(0.f, 5, 5, 0, 1); mesh.setSampleAt<float>(0.f, 5, 5, 24, 0);
    CHECK_EQ(mesh.getDeviceOfElement(1, 1, 10), 0); CHECK_EQ(mesh.getDeviceOfElement(1, 1, 48), 0); CHECK_EQ(mesh.getDeviceOfElement(1, 1, 60), -1);
    CHECK_EQ((int)mesh.getPositionSample(10, 10, 10), 0);   // an all-solid volume: every position byte is 0
    CHECK_THROW(mesh.getPositionSample(10, 10, 200), std::out_of_range);
    mesh.switchHalos();                             // slab0[23] -> slab1[0], slab1[1] -> slab0[24]
    CHECK_EQ(mesh.getSampleAt<float>(5, 5, 0, 1), 7.f); CHECK_EQ(mesh.getSampleAt<float>(5, 5, 24, 0), 9.f);
  }
  { // CudaMesh_Run_single (:472-575): 1 / 2 / 5 partitions give bitwise equal responses; double too
    std::vector<unsigned char> bid, mat; shoebox_bid(60, 44, 49, bid, mat);
    for (int dbl = 0; dbl < 2; dbl++) {
      std::vector<std::vector<double> > runs;
      for (unsigned np : {1u, 2u, 5u}) {
        SimulationParameters p; p.setSpatialFs(10000); p.setNumSteps(300); p.setUpdateType(SRL_FORWARD); p.setAddPaddingToElementIdx(false);
        const float dx = p.getC() / ((float)p.getSpatialFs() * (float)p.getLambda());
        p.addSource(Source(10 * dx, 10 * dx, 3 * dx)); p.addSource(Source(20 * dx, 12 * dx, 23 * dx)); p.addSource(Source(30 * dx, 30 * dx, 40 * dx));
        p.addReceiver(Receiver(25 * dx, 20 * dx, 5 * dx)); p.addReceiver(Receiver(25 * dx, 20 * dx, 24 * dx)); p.addReceiver(Receiver(25 * dx, 20 * dx, 44 * dx));
        CudaMesh mesh; mesh.setDouble(dbl != 0);
        mesh.setupMeshHost(&bid[0], &mat[0], 1, dbl ? (void*)mh.getMaterialCoefficientPtrDouble() : (void*)mh.getMaterialCoefficientPtr(),
                           dbl ? (void*)p.getParameterPtrDouble() : (void*)p.getParameterPtr(), make_uint3(60, 44, 49), make_uint3(32, 4, 1), 0);
        mesh.makePartition(np, std::vector<unsigned>(np, 0));
        std::vector<double> r(3 * 300, 0.0);
        if (dbl) launchFDTD3dDouble(&mesh, &p, &r[0], never, quiet);
        else { std::vector<float> rf(3 * 300, 0.f); launchFDTD3d(&mesh, &p, &rf[0], never, quiet); for (size_t i = 0; i < rf.size(); i++) r[i] = rf[i]; }
        runs.push_back(r);
      }
      double mx = 0; bool same = true;
      for (size_t i = 0; i < runs[0].size(); i++) { mx = std::fmax(mx, std::fabs(runs[0][i])); same &= runs[0][i] == runs[1][i] && runs[0][i] == runs[2][i]; }
      CHECK(mx > 0); CHECK(same);
    }
  }
  { // FDTD::App end to end: box geometry -> built-in voxelizer -> runSimulation; executeStep path gives the same response
    std::vector<unsigned> idx; std::vector<float> v; box_mesh(2.f, 1.5f, 1.2f, idx, v);
    std::vector<float> resp[2];
    for (int mode = 0; mode < 2; mode++) {
      FDTD::App app; app.m_progress = quiet; app.initializeDevices();
      app.initializeGeometry(&idx[0], &v[0], (unsigned)idx.size(), (unsigned)v.size());   // float count, like the reference's callers
      app.setUniformMaterial(0.9f); app.setSpatialFs(7000); app.setNumSteps(120); app.setUpdateType(0); app.setForcePartitionTo(1);
      app.addSource(0.5f, 0.5f, 0.5f, 0, 0, 0); app.addReceiver(1.5f, 1.0f, 0.7f);
      if (mode == 0) { app.runSimulation(); CHECK(app.getMvoxPerSec() > 0); CHECK(app.getVolume() > 2.5f && app.getVolume() < 4.5f); CHECK(app.getSabine(0) > 0); }
      else { app.addMeshToCapture(50); app.addSliceToCapture(8, 60, 0); app.runCapture(); CHECK_EQ(app.getNumberOfMeshCaptures(), 1u); CHECK_EQ(app.getNumberOfSliceCaptures(), 1u); }
      resp[mode] = app.getResponse(0);
      app.close();
    }
    double mx = 0; bool same = resp[0].size() == resp[1].size();
    for (size_t i = 0; same && i < resp[0].size(); i++) { mx = std::fmax(mx, std::fabs(resp[0][i])); same &= resp[0][i] == resp[1][i]; }
    CHECK(mx > 0); CHECK(same);
  }
  { // launchFDTD3dStep re-uploads its source / receiver tables when the parameters change with the COUNTS unchanged
    // (a moved source, another App / parameters object on the same thread): no stale tables
    std::vector<unsigned char> bid, mat; shoebox_bid(60, 44, 49, bid, mat);
    MaterialHandler mh; mh.setGlobalMaterial(1, reflection2Admitance(0.9f));
    auto run_steps = [&](CudaMesh& mesh, SimulationParameters& p, std::vector<float>& out, unsigned first, unsigned n) {
      for (unsigned i = first; i < first + n; i++) launchFDTD3dStep(&mesh, &p, &out[0], i, 1, quiet);
    };
    auto fresh = [&](float sx, std::vector<float>& out) {
      SimulationParameters p; p.setSpatialFs(7000); p.setUpdateType(SRL_FORWARD); p.setNumSteps(40);
      const float dx = p.getDx();
      p.addSource(Source(sx * dx, 20 * dx, 24 * dx, SRC_HARD, IMPULSE, 0)); p.addReceiver(Receiver(25 * dx, 20 * dx, 24 * dx));
      CudaMesh mesh;
      mesh.setupMeshHost(&bid[0], &mat[0], 1, mh.getMaterialCoefficientPtr(), p.getParameterPtr(), make_uint3(60, 44, 49), make_uint3(32, 4, 1), 0);
      mesh.makePartition(1);
      out.assign(40, 0.f); run_steps(mesh, p, out, 0, 40);
    };
    std::vector<float> near_src, far_src, moved(40, 0.f);
    fresh(22.f, near_src); fresh(12.f, far_src);
    CHECK(near_src != far_src);
    SimulationParameters p; p.setSpatialFs(7000); p.setUpdateType(SRL_FORWARD); p.setNumSteps(40);
    const float dx = p.getDx();
    p.addSource(Source(22 * dx, 20 * dx, 24 * dx, SRC_HARD, IMPULSE, 0)); p.addReceiver(Receiver(25 * dx, 20 * dx, 24 * dx));
    CudaMesh mesh;
    mesh.setupMeshHost(&bid[0], &mat[0], 1, mh.getMaterialCoefficientPtr(), p.getParameterPtr(), make_uint3(60, 44, 49), make_uint3(32, 4, 1), 0);
    mesh.makePartition(1);
    run_steps(mesh, p, moved, 0, 5);
    p.updateSourceAt(0, Source(12 * dx, 20 * dx, 24 * dx, SRC_HARD, IMPULSE, 0));   // same counts, other position
    mesh.resetPressures();
    run_steps(mesh, p, moved, 0, 40);
    CHECK(moved == far_src);
  }
}

int main(int argc, char** argv) {
  const std::string what = argc > 1 ? argv[1] : "cpu";
  if (what == "cpu") { test_simulation_parameters(); test_parameter_generation(); test_srcrec(); test_material_handler(); test_device_helpers_without_a_device(); test_file_reader(); test_partition_indexing(); test_geometry_and_voxelizer(); test_voxelizer_materials_follow_the_nearest_triangle(); }
  else if (what == "gpu") { test_cuda_mesh_gpu(); test_voxelizer_gpu(); }
  std::printf("%s: %d checks, %d failures\n", what.c_str(), g_checks, g_fail);
  return g_fail ? 1 : 0;
}
