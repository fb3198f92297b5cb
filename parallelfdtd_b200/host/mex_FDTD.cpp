// mex_FDTD.cpp -- the MATLAB gateway `mex_FDTD` over FDTD::App (this build's App: hand-written sm_100a CUDA behind
// the C ABI of include/pfdtd.h).  It keeps the calling convention of the reference's gateway
// (reference matlab/mex_FDTD.cpp:30-368, called from matlab/runFDTD.m and matlab/testBench.m), so the MATLAB side
// does not change:
//
//   [p]                                     = mex_FDTD(a1 .. a15)
//   [p, n_elements, t_step]                 = mex_FDTD(a1 .. a15)
//   [p, n_elements, t_step, dimX, dimY, dimZ, dx, mesh_captures] = mex_FDTD(a1 .. a15)
//
//    # | argument                         | class  | shape
//   ---+----------------------------------+--------+---------------------------------------------
//    1 | vertices                         | single | 3 x nVertices
//    2 | triangle indices (0-based)       | uint32 | 3 x nTriangles
//    3 | material coefficients            | single | nCoefficients x nSurfaces
//    4 | sources [x y z type input group] | single | 6 x nSources
//    5 | receivers [x y z]                | single | 3 x nReceivers
//    6 | source input data                | double | nSamples x nVectors
//    7 | spatial fs                       | uint32 | scalar
//    8 | number of steps                  | uint32 | scalar
//    9 | update type                      | uint32 | scalar (0 SRL_FORWARD, 1 SHARED, 2 SRL, 3 IISO, 4 IWB)
//   10 | visualization                    | uint32 | scalar (the OpenGL viewer is not part of this build: error if 1)
//   11 | slice captures [slice step dim]  | uint32 | 3 x nCaptures
//   12 | mesh capture steps               | uint32 | 1 x nMeshCaptures
//   13 | double precision                 | uint32 | scalar
//   14 | force partition to               | uint32 | scalar
//   15 | octave                           | uint32 | scalar
//   16 | (optional, addition) filter order| uint32 | scalar 0..4: rows of argument 3 are digital impedance filters
//      |                                  |        | [b0 .. bN, a1 .. aN] per surface (PFDTD_OPT_DIF_ORDER)
//
// p is nReceivers x nSteps (single, or double when argument 13 is set and the 8-output form is used -- the 1- and
// 3-output forms always return single, as in the reference, :258-286).  mesh_captures is nMeshCaptures x nElements.
//
// Differences from the reference gateway, on purpose: argument classes and sizes are checked (a wrong class is an
// error message, not a wild read); every call starts from a fresh App (the reference's file-scope App keeps the
// sources and receivers of earlier calls until `clear mex`); an exception from the solver becomes a MATLAB error
// with the library's message instead of a silent return without outputs.
//
// Build inside MATLAB:  mex -I<repo>/parallelfdtd_b200/host -I<repo>/include mex_FDTD.cpp
//                                            -L<repo>/parallelfdtd_b200 -lpfdtd_host -lpfdtd_b200
#include "mex.h"

#include <csignal>
#include <exception>
#include <memory>
#include <string>
#include <vector>

#include "App.h"

// MATLAB's interrupt query (libut); the same undocumented hook the reference polls (mex_FDTD.cpp:26-28)
extern "C" bool utIsInterruptPending();

namespace {

FDTD::App* g_app = 0;                                       // for the SIGINT handler only

void on_sigint(int signum) {
  if (signum == SIGINT && g_app) g_app->close();
}

void progress_to_matlab(int step, int max_step, float t_per_step) {
  mexPrintf("Step %d/%d, time per step %f, estimated time left %f s \n", step, max_step, t_per_step, t_per_step * (float)(max_step - step));
  mexEvalString("drawnow;");
}

bool interrupt_from_matlab(void) { return utIsInterruptPending(); }

void fail(const std::string& msg) { mexErrMsgTxt(msg.c_str()); }

const float* single_matrix(const mxArray* a, size_t rows, const char* what, size_t* cols) {
  if (!mxIsSingle(a)) fail(std::string(what) + " must be single");
  const size_t n = mxGetNumberOfElements(a);
  if (n != 0 && mxGetM(a) != rows) fail(std::string(what) + " must have " + std::to_string(rows) + " rows");
  *cols = n == 0 ? 0 : mxGetN(a);
  return (const float*)mxGetData(a);
}

const unsigned int* uint32_matrix(const mxArray* a, size_t rows, const char* what, size_t* cols) {
  const size_t n = mxGetNumberOfElements(a);
  if (n != 0 && !mxIsUint32(a)) fail(std::string(what) + " must be uint32");
  if (n != 0 && rows != 0 && mxGetM(a) != rows) fail(std::string(what) + " must have " + std::to_string(rows) + " rows");
  *cols = n == 0 ? 0 : (rows == 0 ? n : mxGetN(a));
  return (const unsigned int*)mxGetData(a);
}

unsigned int uint32_scalar(const mxArray* a, const char* what) {
  if (!mxIsUint32(a) || mxGetNumberOfElements(a) < 1) fail(std::string(what) + " must be a uint32 scalar");
  return *(const unsigned int*)mxGetData(a);
}

mxArray* single_scalar(float v) {
  mxArray* a = mxCreateNumericMatrix(1, 1, mxSINGLE_CLASS, mxREAL);
  *(float*)mxGetData(a) = v;
  return a;
}

// responses as an nReceivers x nSteps column-major matrix of T
template <typename T, typename Get>
mxArray* response_matrix(size_t n_rec, size_t n_steps, mxClassID cls, bool have, Get get) {
  mxArray* a = mxCreateNumericMatrix(n_rec, n_steps, cls, mxREAL);
  if (!have) return a;
  T* p = (T*)mxGetData(a);
  for (size_t step = 0; step < n_steps; step++)
    for (size_t rec = 0; rec < n_rec; rec++) p[step * n_rec + rec] = get((unsigned int)step, (unsigned int)rec);
  return a;
}

}  // namespace

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 15 && nrhs != 16) fail("mex_FDTD: 15 input arguments expected (16 with a filter order)");
  if (nlhs != 0 && nlhs != 1 && nlhs != 3 && nlhs != 8) fail("mex_FDTD: 1, 3 or 8 output arguments expected");

  // ---- inputs
  size_t n_vertices = 0, n_triangles = 0, n_surfaces = 0, n_sources = 0, n_receivers = 0, n_captures = 0, n_mesh_captures = 0;
  const float* vertices = single_matrix(prhs[0], 3, "vertices", &n_vertices);
  const unsigned int* indices = uint32_matrix(prhs[1], 3, "indices", &n_triangles);
  if (n_vertices == 0 || n_triangles == 0) {
    mexPrintf("No geometry assigned, check the geometry file, returning\n");   // reference :99-102
    return;
  }
  if (!mxIsSingle(prhs[2])) fail("materials must be single");
  const size_t n_coefficients = mxGetM(prhs[2]);
  n_surfaces = mxGetN(prhs[2]);
  const float* materials = (const float*)mxGetData(prhs[2]);
  const float* src_list = single_matrix(prhs[3], 6, "sources", &n_sources);
  const float* rec_list = single_matrix(prhs[4], 3, "receivers", &n_receivers);
  if (mxGetNumberOfElements(prhs[5]) != 0 && !mxIsDouble(prhs[5])) fail("source input data must be double");
  const size_t n_input_samples = mxGetM(prhs[5]);
  const size_t n_input_vectors = mxGetNumberOfElements(prhs[5]) == 0 ? 0 : mxGetN(prhs[5]);
  const double* input_data = (const double*)mxGetData(prhs[5]);
  const unsigned int spatial_fs = uint32_scalar(prhs[6], "spatial fs");
  const unsigned int n_steps = uint32_scalar(prhs[7], "number of steps");
  const unsigned int update_type = uint32_scalar(prhs[8], "update type");
  const unsigned int visualization = uint32_scalar(prhs[9], "visualization");
  const unsigned int* captures = uint32_matrix(prhs[10], 3, "slice captures", &n_captures);
  const unsigned int* mesh_captures = uint32_matrix(prhs[11], 0, "mesh captures", &n_mesh_captures);
  const bool double_precision = uint32_scalar(prhs[12], "double precision") == 1 && visualization != 1;
  const unsigned int force_partition_to = uint32_scalar(prhs[13], "force partition to");
  const unsigned int octave = uint32_scalar(prhs[14], "octave");
  const unsigned int filter_order = nrhs == 16 ? uint32_scalar(prhs[15], "filter order") : 0;
  if (update_type > 4) fail("update type must be 0..4");
  if (filter_order > 4) fail("filter order must be 0..4");
  if (filter_order && n_coefficients < 2 * filter_order + 1) fail("materials: a filter of order N needs 2N+1 coefficients per surface");
  if (!filter_order && n_coefficients > MATERIAL_COEF_NUM) fail("materials: at most 20 coefficients per surface");
  if (visualization == 1) fail("mex_FDTD: the OpenGL viewer is not part of this build; request slice captures instead");

  mexPrintf("Number of vertices : %u \n", (unsigned)n_vertices);
  mexPrintf("Number of triangles : %u \n", (unsigned)n_triangles);
  mexPrintf("Number of surfaces in materials : %u \n", (unsigned)n_surfaces);
  mexPrintf("Number of sources %u, receivers %u, input data samples %u, captures %u \n", (unsigned)n_sources, (unsigned)n_receivers,
            (unsigned)n_input_samples, (unsigned)n_captures);
  mexEvalString("drawnow;");

  // ---- a fresh App per call
  std::unique_ptr<FDTD::App> app(new FDTD::App());
  g_app = app.get();
  void (*prev_handler)(int) = std::signal(SIGINT, on_sigint);
  std::string error;
  try {
    app->initializeDevices();
    app->m_interrupt = interrupt_from_matlab;
    app->m_progress = progress_to_matlab;
    app->initializeGeometry(const_cast<unsigned int*>(indices), const_cast<float*>(vertices), (unsigned int)(3 * n_triangles),
                            (unsigned int)(3 * n_vertices));
    if (filter_order) {
      std::vector<float> rows((size_t)n_surfaces * (2 * filter_order + 1));
      for (size_t s = 0; s < n_surfaces; s++)
        for (size_t k = 0; k < 2 * filter_order + 1; k++) rows[s * (2 * filter_order + 1) + k] = materials[s * n_coefficients + k];
      app->m_materials.addFilterMaterials(rows.data(), (unsigned int)n_surfaces, filter_order);
    } else {
      app->m_materials.addMaterials(const_cast<float*>(materials), (unsigned int)n_surfaces, (unsigned int)n_coefficients);
    }
    app->m_parameters.setSpatialFs(spatial_fs);
    app->m_parameters.setNumSteps(n_steps);
    app->m_parameters.readGridIr("./Data/grid_ir.txt");         // transparent sources only; a missing file is logged
    app->m_parameters.setUpdateType((enum UpdateType)update_type);
    app->m_parameters.setOctave(octave);
    app->setForcePartitionTo((int)force_partition_to);

    for (size_t v = 0; v < n_input_vectors; v++) {
      const double* col = input_data + v * n_input_samples;
      if (double_precision) app->m_parameters.addInputDataDouble(std::vector<double>(col, col + n_input_samples));
      else app->m_parameters.addInputData(std::vector<float>(col, col + n_input_samples));   // narrowed element-wise
    }
    for (size_t i = 0; i < n_sources; i++) {
      const float* s = src_list + 6 * i;
      app->m_parameters.addSource(Source(s[0], s[1], s[2], (enum SrcType)(unsigned int)s[3], (enum InputType)(unsigned int)s[4], (int)s[5]));
    }
    for (size_t i = 0; i < n_receivers; i++) app->m_parameters.addReceiver(Receiver(rec_list[3 * i], rec_list[3 * i + 1], rec_list[3 * i + 2]));
    for (size_t i = 0; i < n_captures; i++) app->addSliceToCapture(captures[3 * i], captures[3 * i + 1], captures[3 * i + 2]);
    for (size_t i = 0; i < n_mesh_captures; i++) app->addMeshToCapture(mesh_captures[i]);

    if (n_captures == 0 && n_mesh_captures == 0) {               // reference :239-246
      if (double_precision) app->m_mesh.setDouble(true);
      app->runSimulation();
    } else {
      app->runCapture();                                         // step by step, single precision
    }
  } catch (int code) {                                           // the solver's error convention (cudaUtils.h:47-51)
    const char* msg = pfdtd_last_error();
    error = "mex_FDTD: solver error " + std::to_string(code) + (msg && *msg ? std::string(": ") + msg : std::string());
  } catch (const std::exception& e) {
    error = std::string("mex_FDTD: ") + e.what();
  }
  std::signal(SIGINT, prev_handler);
  if (!error.empty()) {
    app->close();
    g_app = 0;
    app.reset();
    fail(error);                                                 // does not return inside MATLAB
    return;
  }

  // ---- outputs
  const bool have = app->getResponseSize() != 0;
  const bool dbl = app->m_mesh.isDouble();
  FDTD::App* a = app.get();
  auto get_f = [a, dbl](unsigned int step, unsigned int rec) {
    return dbl ? (float)a->getResponseDoubleSampleAt(step, rec) : a->getResponseSampleAt(step, rec);
  };
  auto get_d = [a](unsigned int step, unsigned int rec) { return a->getResponseDoubleSampleAt(step, rec); };
  if (nlhs <= 3) {
    plhs[0] = response_matrix<float>(n_receivers, n_steps, mxSINGLE_CLASS, have, get_f);
    if (nlhs == 3) {
      plhs[1] = single_scalar((float)app->getNumElements());
      plhs[2] = single_scalar(app->getTimePerStep());
    }
  } else {                                                       // 8 outputs, reference :289-365
    plhs[0] = dbl ? response_matrix<double>(n_receivers, n_steps, mxDOUBLE_CLASS, have, get_d)
                  : response_matrix<float>(n_receivers, n_steps, mxSINGLE_CLASS, have, get_f);
    plhs[1] = single_scalar((float)app->getNumElements());
    plhs[2] = single_scalar(app->getTimePerStep());
    plhs[3] = single_scalar((float)app->m_mesh.getDimX());
    plhs[4] = single_scalar((float)app->m_mesh.getDimY());
    plhs[5] = single_scalar((float)app->m_mesh.getDimZ());
    plhs[6] = single_scalar(app->m_parameters.getDx());
    const size_t n_caps = app->getNumberOfMeshCaptures();
    const size_t n_elements = (size_t)app->m_mesh.getNumberOfElements64();
    plhs[7] = mxCreateNumericMatrix(n_caps, n_caps ? n_elements : 0, mxSINGLE_CLASS, mxREAL);
    float* out = (float*)mxGetData(plhs[7]);
    for (size_t c = 0; c < n_caps; c++) {
      const float* field = app->getMeshCaptureAt((unsigned int)c);
      for (size_t e = 0; e < n_elements; e++) out[e * n_caps + c] = field[e];
    }
  }
  app->close();
  g_app = 0;
}
