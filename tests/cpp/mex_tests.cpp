// mex_tests.cpp -- drives the MATLAB gateway (parallelfdtd_b200/host/mex_FDTD.cpp) through the stand-in MEX API of
// tests/cpp/mex_stub/mex.h with the argument list matlab/runFDTD.m builds (reference matlab/runFDTD.m:60-84,
// matlab/testBench.m), and compares what comes back with the same run made directly on FDTD::App.
//   mex_tests cpu   argument checking, behaviour without a device
//   mex_tests gpu   1 / 3 / 8 output forms, double precision, captures, filter materials
#include "mex.h"

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "App.h"

extern "C" bool utIsInterruptPending() { return false; }

static int g_checks = 0, g_fail = 0;
#define CHECK(c) do { g_checks++; if (!(c)) { g_fail++; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)

template <typename T> static mxArray* matrix(size_t m, size_t n, mxClassID cls, const std::vector<T>& v) {
  mxArray* a = mxCreateNumericMatrix(m, n, cls, mxREAL);
  std::memcpy(mxGetData(a), v.data(), v.size() * sizeof(T));
  return a;
}
static mxArray* u32(unsigned int v) { return matrix<unsigned int>(1, 1, mxUINT32_CLASS, std::vector<unsigned int>(1, v)); }

struct Job {
  std::vector<float> vertices, materials, sources, receivers;
  std::vector<unsigned int> indices, captures, mesh_captures;
  std::vector<double> input;
  size_t n_coef = 20, n_input_samples = 0;
  unsigned int fs = 7000, steps = 120, update_type = 0, visualization = 0, dbl = 0, force = 1, octave = 0;
  int filter_order = -1;                                      // -1: 15 arguments
};

static Job box_job() {
  Job j;
  const float L = 1.f;
  const float v[24] = {0, 0, 0, L, 0, 0, L, L, 0, 0, L, 0, 0, 0, L, L, 0, L, L, L, L, 0, L, L};
  j.vertices.assign(v, v + 24);
  const unsigned int q[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {1, 2, 6, 5}, {3, 0, 4, 7}};
  for (auto& f : q) { const unsigned int t[6] = {f[0], f[1], f[2], f[0], f[2], f[3]}; j.indices.insert(j.indices.end(), t, t + 6); }
  const size_t n_tri = j.indices.size() / 3;
  for (size_t s = 0; s < n_tri; s++)
    for (size_t k = 0; k < 20; k++) j.materials.push_back(reflection2Admitance(0.9f - 0.05f * (float)(s / 2)));
  const float src[6] = {0.5f, 0.5f, 0.5f, 0.f, 0.f, 0.f};    // hard impulse
  j.sources.assign(src, src + 6);
  const float rec[6] = {0.6f, 0.6f, 0.6f, 0.3f, 0.4f, 0.7f};
  j.receivers.assign(rec, rec + 6);
  return j;
}

struct Call {
  std::vector<mxArray*> in, out;
  std::string error;
  ~Call() { for (auto a : in) mxDestroyArray(a); for (auto a : out) mxDestroyArray(a); }
};

static void call(const Job& j, int nlhs, Call& c) {
  c.in = {matrix<float>(3, j.vertices.size() / 3, mxSINGLE_CLASS, j.vertices),
          matrix<unsigned int>(3, j.indices.size() / 3, mxUINT32_CLASS, j.indices),
          matrix<float>(j.n_coef, j.materials.size() / j.n_coef, mxSINGLE_CLASS, j.materials),
          matrix<float>(6, j.sources.size() / 6, mxSINGLE_CLASS, j.sources),
          matrix<float>(3, j.receivers.size() / 3, mxSINGLE_CLASS, j.receivers),
          matrix<double>(j.n_input_samples, j.n_input_samples ? j.input.size() / j.n_input_samples : 0, mxDOUBLE_CLASS, j.input),
          u32(j.fs), u32(j.steps), u32(j.update_type), u32(j.visualization),
          matrix<unsigned int>(3, j.captures.size() / 3, mxUINT32_CLASS, j.captures),
          matrix<unsigned int>(1, j.mesh_captures.size(), mxUINT32_CLASS, j.mesh_captures),
          u32(j.dbl), u32(j.force), u32(j.octave)};
  if (j.filter_order >= 0) c.in.push_back(u32((unsigned int)j.filter_order));
  c.out.assign(8, (mxArray*)0);
  try {
    mexFunction(nlhs, c.out.data(), (int)c.in.size(), (const mxArray**)c.in.data());
  } catch (const MexError& e) {
    c.error = e.what();
  }
}

static bool has_device() { int n = 0; pfdtd_device_count(&n); return n > 0; }

static void test_cpu() {
  Job j = box_job();
  { Call c; c.in = {u32(1), u32(2), u32(3)}; c.out.assign(8, (mxArray*)0);
    try { mexFunction(1, c.out.data(), 3, (const mxArray**)c.in.data()); } catch (const MexError& e) { c.error = e.what(); }
    CHECK(c.error.find("15 input arguments") != std::string::npos); }
  { Call c; call(j, 2, c); CHECK(c.error.find("output arguments") != std::string::npos); }
  { Call c; Job k = j; k.visualization = 1; call(k, 1, c); CHECK(c.error.find("OpenGL") != std::string::npos); }
  { Call c; Job k = j; k.update_type = 9; call(k, 1, c); CHECK(c.error.find("update type") != std::string::npos); }
  { Call c; Job k = j; k.filter_order = 7; call(k, 1, c); CHECK(c.error.find("filter order") != std::string::npos); }
  { Call c; Job k = j; k.n_coef = 4; k.materials.resize(4 * 12); k.filter_order = 2; call(k, 1, c); CHECK(c.error.find("2N+1") != std::string::npos); }
  { Call c; Job k = j; k.vertices.clear(); call(k, 1, c); CHECK(c.error.empty()); CHECK(c.out[0] == 0);      // no geometry: returns quietly
    CHECK(mexStubLog().find("No geometry assigned") != std::string::npos); }
  { Call c; call(j, 1, c);                                                // a wrong class is an error, not a wild read
    Call d; d.in = {matrix<double>(3, 8, mxDOUBLE_CLASS, std::vector<double>(24, 0.0))};
    for (size_t i = 1; i < c.in.size(); i++) { d.in.push_back(c.in[i]); }
    d.out.assign(8, (mxArray*)0);
    try { mexFunction(1, d.out.data(), 15, (const mxArray**)d.in.data()); } catch (const MexError& e) { d.error = e.what(); }
    CHECK(d.error.find("vertices must be single") != std::string::npos);
    d.in.resize(1);                                                       // the rest belongs to c
    if (!has_device()) { CHECK(c.error.find("solver error") != std::string::npos); CHECK(c.out[0] == 0); }   // no CPU fallback
    else { CHECK(c.error.empty()); CHECK(c.out[0] != 0); } }
}

static void quiet(int, int, float) {}

// the same job made directly on FDTD::App
static void direct(const Job& j, std::vector<float>& resp, std::vector<double>& resp_d, FDTD::App& app) {
  app.m_progress = quiet;
  app.initializeDevices();
  app.initializeGeometry(const_cast<unsigned int*>(j.indices.data()), const_cast<float*>(j.vertices.data()), (unsigned int)j.indices.size(),
                         (unsigned int)j.vertices.size());
  const unsigned int n_surf = (unsigned int)(j.materials.size() / j.n_coef);
  if (j.filter_order > 0) {
    std::vector<float> rows;
    for (unsigned int s = 0; s < n_surf; s++) for (int k = 0; k < 2 * j.filter_order + 1; k++) rows.push_back(j.materials[s * j.n_coef + k]);
    app.m_materials.addFilterMaterials(rows.data(), n_surf, (unsigned int)j.filter_order);
  } else {
    app.m_materials.addMaterials(const_cast<float*>(j.materials.data()), n_surf, (unsigned int)j.n_coef);
  }
  app.m_parameters.setSpatialFs(j.fs);
  app.m_parameters.setNumSteps(j.steps);
  app.m_parameters.setUpdateType((enum UpdateType)j.update_type);
  app.m_parameters.setOctave(j.octave);
  app.setForcePartitionTo((int)j.force);
  for (size_t i = 0; i < j.sources.size() / 6; i++) {
    const float* s = &j.sources[6 * i];
    app.m_parameters.addSource(Source(s[0], s[1], s[2], (enum SrcType)(unsigned)s[3], (enum InputType)(unsigned)s[4], (unsigned)s[5]));
  }
  for (size_t i = 0; i < j.receivers.size() / 3; i++) app.m_parameters.addReceiver(j.receivers[3 * i], j.receivers[3 * i + 1], j.receivers[3 * i + 2]);
  if (j.dbl) app.m_mesh.setDouble(true);
  app.runSimulation();
  const unsigned int nr = app.m_parameters.getNumReceivers();
  for (unsigned int s = 0; s < j.steps; s++)
    for (unsigned int r = 0; r < nr; r++) {
      if (j.dbl) resp_d.push_back(app.getResponseDoubleSampleAt(s, r)); else resp.push_back(app.getResponseSampleAt(s, r));
    }
}

static void test_gpu() {
  Job j = box_job();
  std::vector<float> ref; std::vector<double> ref_d;
  unsigned int dims[3]; float dx; unsigned int n_elem;
  { FDTD::App app; direct(j, ref, ref_d, app);
    dims[0] = app.m_mesh.getDimX(); dims[1] = app.m_mesh.getDimY(); dims[2] = app.m_mesh.getDimZ(); dx = app.m_parameters.getDx();
    n_elem = app.getNumElements(); app.close(); }
  float mx = 0; for (float v : ref) mx = std::fmax(mx, std::fabs(v));
  CHECK(mx > 0); CHECK(ref.size() == 2u * j.steps);

  { Call c; call(j, 1, c);                                                // [p] = mex_FDTD(...)
    CHECK(c.error.empty()); CHECK(c.out[0] && mxGetM(c.out[0]) == 2 && mxGetN(c.out[0]) == j.steps && mxIsSingle(c.out[0]));
    if (c.out[0]) CHECK(std::memcmp(mxGetData(c.out[0]), ref.data(), ref.size() * 4) == 0); }
  { Call c; call(j, 3, c);                                                // [p, n_elements, t_step]
    CHECK(c.error.empty()); CHECK(c.out[2] != 0);
    if (c.out[2]) { CHECK(std::memcmp(mxGetData(c.out[0]), ref.data(), ref.size() * 4) == 0);
      CHECK(*(float*)mxGetData(c.out[1]) == (float)n_elem); CHECK(*(float*)mxGetData(c.out[2]) > 0.f); } }
  { Job k = j; k.dbl = 1;                                                 // 8 outputs, double precision
    std::vector<float> r; std::vector<double> rd;
    { FDTD::App app; direct(k, r, rd, app); app.close(); }
    Call c; call(k, 8, c);
    CHECK(c.error.empty()); CHECK(c.out[7] != 0);
    if (c.out[7]) {
      CHECK(mxIsDouble(c.out[0]) && mxGetM(c.out[0]) == 2 && mxGetN(c.out[0]) == k.steps);
      CHECK(std::memcmp(mxGetData(c.out[0]), rd.data(), rd.size() * 8) == 0);
      CHECK(*(float*)mxGetData(c.out[3]) == (float)dims[0]); CHECK(*(float*)mxGetData(c.out[4]) == (float)dims[1]);
      CHECK(*(float*)mxGetData(c.out[5]) == (float)dims[2]); CHECK(*(float*)mxGetData(c.out[6]) == dx);
      CHECK(mxGetNumberOfElements(c.out[7]) == 0); }
    Call c1; call(k, 1, c1);                                              // 1-output form narrows to single (reference :258-270)
    CHECK(c1.error.empty()); CHECK(c1.out[0] && mxIsSingle(c1.out[0]));
    if (c1.out[0]) { const float* p = (const float*)mxGetData(c1.out[0]); bool same = true;
      for (size_t i = 0; i < rd.size(); i++) same &= p[i] == (float)rd[i];
      CHECK(same); } }
  { Job k = j; k.captures = {10, 50, 1}; k.mesh_captures = {60};          // capture run: step by step, single precision
    k.dbl = 1;                                                            // ignored by runCapture (reference App.cpp:354)
    Call c; call(k, 8, c);
    CHECK(c.error.empty()); CHECK(c.out[7] != 0);
    if (c.out[7]) {
      CHECK(mxIsSingle(c.out[0])); CHECK(std::memcmp(mxGetData(c.out[0]), ref.data(), ref.size() * 4) == 0);
      CHECK(mxGetM(c.out[7]) == 1 && mxGetN(c.out[7]) == (size_t)dims[0] * dims[1] * dims[2]);
      const float* f = (const float*)mxGetData(c.out[7]); float m = 0;
      for (size_t i = 0; i < mxGetNumberOfElements(c.out[7]); i++) m = std::fmax(m, std::fabs(f[i]));
      CHECK(m > 0); } }
  { Job k = j; k.filter_order = 2; k.n_coef = 5; k.materials.clear();     // filter materials through the 16th argument
    for (size_t s = 0; s < 12; s++) { const float y0 = reflection2Admitance(0.9f - 0.05f * (float)(s / 2));
      const float row[5] = {y0, -0.71f * y0, 0.1258f * y0, -0.98f, 0.2365f}; /* zeros 0.37, 0.34; poles 0.55, 0.43 */ k.materials.insert(k.materials.end(), row, row + 5); }
    std::vector<float> r; std::vector<double> rd;
    { FDTD::App app; direct(k, r, rd, app); app.close(); }
    Call c; call(k, 1, c);
    CHECK(c.error.empty()); CHECK(c.out[0] != 0);
    if (c.out[0]) { CHECK(std::memcmp(mxGetData(c.out[0]), r.data(), r.size() * 4) == 0); CHECK(std::memcmp(r.data(), ref.data(), r.size() * 4) != 0); } }
  { Job k = j; k.sources[4] = 3.f; k.sources[5] = 0.f; k.n_input_samples = 6; k.input = {0.0, 0.25, 1.0, 0.25, -0.5, 0.0};   // DATA source
    std::vector<float> r; std::vector<double> rd;
    { FDTD::App app; std::vector<float> in(k.input.begin(), k.input.end()); app.m_parameters.addInputData(in); direct(k, r, rd, app); app.close(); }
    Call c; call(k, 1, c);
    CHECK(c.error.empty()); CHECK(c.out[0] != 0);
    if (c.out[0]) { CHECK(std::memcmp(mxGetData(c.out[0]), r.data(), r.size() * 4) == 0); CHECK(std::memcmp(r.data(), ref.data(), r.size() * 4) != 0); } }
}

int main(int argc, char** argv) {
  const std::string what = argc > 1 ? argv[1] : "cpu";
  if (what == "cpu") test_cpu(); else if (what == "gpu") test_gpu();
  std::printf("%s: %d checks, %d failures\n", what.c_str(), g_checks, g_fail);
  return g_fail ? 1 : 0;
}
