// AppPy.cpp -- the Python module `libPyFDTD`: class `App` with the method names of the reference's
// boost::python module (reference src/AppPy.cpp:96-133), so scripts written against it
// (reference python/testBench.py:110-149) keep running.  pybind11 instead of Boost.Python (Boost is
// not part of this build); list arguments keep their flattened-list meaning
// (reference src/AppPy.cpp:29-94).  The method without a counterpart here (the OpenGL viewer) runs its session headless.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "App.h"

namespace py = pybind11;

namespace {

// the reference throws a bare int (-1) on device errors (cudaUtils.h:47-51); surface it as RuntimeError
template <class F>
auto guarded(F&& f) -> decltype(f()) {
  try {
    return f();
  } catch (int code) {
    const char* msg = pfdtd_last_error();
    throw std::runtime_error(std::string("ParallelFDTD error ") + std::to_string(code) + (msg && *msg ? std::string(": ") + msg : std::string()));
  }
}

bool py_interrupt(void) {
  py::gil_scoped_acquire gil;
  return PyErr_CheckSignals() != 0;
}

// A run with the GIL released.  An interrupted run returns normally from App (pfdtd_run's INTERRUPTED is not an
// error there), but the KeyboardInterrupt that PyErr_CheckSignals raised is still pending: hand it to the caller
// instead of returning with an exception set.
template <class F>
void run_released(F&& f) {
  {
    py::gil_scoped_release nogil;
    guarded(f);
  }
  if (PyErr_Occurred()) throw py::error_already_set();
}

void initializeGeometryPy(FDTD::App& a, const std::vector<unsigned int>& indices, const std::vector<float>& vertices) {
  a.m_geometry.initialize(indices, vertices);                      // reference AppPy.cpp:29-44
}

void addSurfaceMaterials(FDTD::App& a, std::vector<float> coefs, unsigned int n_surfaces, unsigned int n_coefs) {
  if (coefs.size() < (size_t)n_surfaces * n_coefs) throw std::out_of_range("addSurfaceMaterials: list shorter than surfaces*coefficients");
  a.m_materials.addMaterials(coefs.data(), n_surfaces, n_coefs);   // reference AppPy.cpp:57-69
}

// additions: per-surface digital impedance filters, rows [b0 .. bN, a1 .. aN] flattened like addSurfaceMaterials
void addSurfaceFilters(FDTD::App& a, std::vector<float> coefs, unsigned int n_surfaces, unsigned int order) {
  if (order < 1 || order > 4) throw std::out_of_range("addSurfaceFilters: filter order must be 1..4");
  if (coefs.size() < (size_t)n_surfaces * (2 * order + 1)) throw std::out_of_range("addSurfaceFilters: list shorter than surfaces*(2*order+1)");
  a.addSurfaceFilters(coefs.data(), n_surfaces, order);
}

void addSourceDataFloat(FDTD::App& a, const std::vector<float>& data, int num_steps, int num_sources) {
  if ((long long)data.size() < (long long)num_steps * num_sources) throw std::out_of_range("addSourceDataFloat: list shorter than steps*sources");
  a.m_parameters.addInputData(std::vector<float>(data.begin(), data.begin() + (size_t)num_steps * num_sources));   // :71-82
}

void addSourceDataDouble(FDTD::App& a, const std::vector<double>& data, int num_steps, int num_sources) {
  if ((long long)data.size() < (long long)num_steps * num_sources) throw std::out_of_range("addSourceDataDouble: list shorter than steps*sources");
  a.m_parameters.addInputDataDouble(std::vector<double>(data.begin(), data.begin() + (size_t)num_steps * num_sources));   // :84-94
}

void setVoxelVolumes(FDTD::App& a, py::array_t<unsigned char, py::array::c_style | py::array::forcecast> bid,
                     py::array_t<unsigned char, py::array::c_style | py::array::forcecast> mat) {
  if (bid.ndim() != 3 || mat.ndim() != 3) throw std::invalid_argument("setVoxelVolumes: expected [z][y][x] uint8 volumes");
  for (int i = 0; i < 3; i++)
    if (bid.shape(i) != mat.shape(i)) throw std::invalid_argument("setVoxelVolumes: bid and mat shapes differ");
  a.setVoxelVolumes(bid.data(), mat.data(), (unsigned int)bid.shape(2), (unsigned int)bid.shape(1), (unsigned int)bid.shape(0));
}

py::array_t<float> getSliceCapture(FDTD::App& a, unsigned int i) {
  const std::vector<float>& s = a.getSliceCaptureAt(i);
  const FDTD::App::CaptureShape sh = a.getSliceCaptureShapeAt(i);
  py::array_t<float> out({(py::ssize_t)sh.rows, (py::ssize_t)sh.cols});
  std::copy(s.begin(), s.end(), out.mutable_data());
  return out;
}

py::array_t<float> getMeshCapture(FDTD::App& a, unsigned int i) {
  float* p = a.getMeshCaptureAt(i);
  const py::ssize_t X = a.m_mesh.getDimX(), Y = a.m_mesh.getDimY(), Z = a.m_mesh.getDimZ();
  py::array_t<float> out({Z, Y, X});
  std::copy(p, p + (size_t)X * Y * Z, out.mutable_data());
  return out;
}

}  // namespace

PYBIND11_MODULE(libPyFDTD, m) {
  m.doc() = "ParallelFDTD Python module (API of the reference's boost::python libPyFDTD) over libpfdtd_b200 (sm_100a CUDA)";

  py::class_<FDTD::App>(m, "App")
      .def(py::init([]() {
        FDTD::App* a = new FDTD::App();
        a->m_interrupt = py_interrupt;                               // Ctrl-C stops the step loop between blocks
        return a;
      }))
      .def("initializeDevices", [](FDTD::App& a) { guarded([&] { a.initializeDevices(); }); })
      .def("initializeGeometryFromFile", [](FDTD::App& a, const std::string& fp) { guarded([&] { a.initializeGeometryFromFile(fp); }); })
      .def("initializeGeometryPy", &initializeGeometryPy)
      .def("setLayerIndices", [](FDTD::App& a, const std::vector<int>& idx, const std::string& name) { a.m_geometry.setLayerIndices(idx, name); })
      .def("addSource", &FDTD::App::addSource)
      .def("addSourceDataFloat", &addSourceDataFloat)
      .def("addSourceDataDouble", &addSourceDataDouble)
      .def("addReceiver", &FDTD::App::addReceiver)
      .def("addSurfaceMaterials", &addSurfaceMaterials)
      .def("addSurfaceFilters", &addSurfaceFilters, "per-surface digital impedance filters [b0..bN, a1..aN], flattened; order 1..4")
      .def("setUniformFilter", &FDTD::App::setUniformFilter, "the same digital impedance filter (b: order+1 taps, a: a1..aN) on every surface")
      .def("setSpatialFs", &FDTD::App::setSpatialFs)
      .def("setNumSteps", &FDTD::App::setNumSteps)
      .def("setUpdateType", &FDTD::App::setUpdateType)
      .def("setUniform", &FDTD::App::setUniformMaterial)
      .def("setUniformMaterial", &FDTD::App::setUniformMaterial)
      .def("runVisualization", [](FDTD::App& a) { run_released([&] { a.runVisualization(); }); })   // headless: no OpenGL window here
      .def("runSimulation", [](FDTD::App& a) { run_released([&] { a.runSimulation(); }); })
      .def("runCapture", [](FDTD::App& a) { run_released([&] { a.runCapture(); }); })
      .def("getResponse", &FDTD::App::getResponse)
      .def("getResponseDouble", &FDTD::App::getResponseDouble)
      .def("forcePartitionTo", &FDTD::App::setForcePartitionTo)
      .def("addSliceToCapture", &FDTD::App::addSliceToCapture)
      .def("addMeshToCapture", &FDTD::App::addMeshToCapture)
      .def("setDouble", &FDTD::App::setDouble)
      .def("setCapturedB", &FDTD::App::setCapturedB)
      .def("close", [](FDTD::App& a) { guarded([&] { a.close(); }); })
      .def("getMvox", &FDTD::App::getMvoxPerSec)
      .def("getNumElems", &FDTD::App::getNumElements)
      // ---- additions (not in the reference module): numpy access to what the reference only hands to MATLAB
      .def("setVoxelVolumes", &setVoxelVolumes, "voxelizer-style [z][y][x] uint8 volumes (bid 0..27, material index) instead of a triangle mesh")
      .def("getNumberOfSliceCaptures", &FDTD::App::getNumberOfSliceCaptures)
      .def("getSliceCapture", &getSliceCapture)
      .def("getNumberOfMeshCaptures", &FDTD::App::getNumberOfMeshCaptures)
      .def("getMeshCapture", &getMeshCapture)
      .def("getTimePerStep", &FDTD::App::getTimePerStep)
      .def("getDx", [](FDTD::App& a) { return a.m_parameters.getDx(); })
      .def("getDims", [](FDTD::App& a) { return py::make_tuple(a.m_mesh.getDimX(), a.m_mesh.getDimY(), a.m_mesh.getDimZ()); })
      .def("getVolume", &FDTD::App::getVolume)
      .def("getSabine", &FDTD::App::getSabine)
      .def("getEyring", &FDTD::App::getEyring);
}
