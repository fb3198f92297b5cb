"""CPU: the Python module `libPyFDTD` builds, imports without a GPU, exposes every method of the reference's
boost::python module (reference src/AppPy.cpp:106-133) and fails loudly without a device."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the .def(...) names of BOOST_PYTHON_MODULE(libPyFDTD), reference src/AppPy.cpp:106-133
REFERENCE_APP_METHODS = [
    "initializeDevices", "initializeGeometryFromFile", "initializeGeometryPy", "setLayerIndices", "addSource", "addSourceDataFloat",
    "addSourceDataDouble", "addReceiver", "addSurfaceMaterials", "setSpatialFs", "setNumSteps", "setUpdateType", "setUniform",
    "runVisualization", "runSimulation", "runCapture", "setUniformMaterial", "getResponse", "getResponseDouble", "forcePartitionTo",
    "addSliceToCapture", "setDouble", "setCapturedB", "close", "getMvox", "getNumElems"]


@pytest.fixture(scope="module")
def pf():
    from parallelfdtd_b200 import build
    path = build.build_py_module()
    assert os.path.exists(path)
    sys.path.insert(0, os.path.dirname(path))
    import libPyFDTD
    return libPyFDTD


def test_module_exposes_the_reference_api(pf):
    app = pf.App()
    missing = [m for m in REFERENCE_APP_METHODS if not hasattr(app, m)]
    assert not missing, missing


def test_host_side_setters_work_without_a_device(pf, capi):
    app = pf.App()
    app.initializeGeometryPy([0, 1, 2, 0, 2, 3], [0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0])
    app.setLayerIndices([0, 1], "floor")
    app.setUpdateType(0)
    app.setNumSteps(10)
    app.setSpatialFs(7000)
    app.addSurfaceMaterials([0.5] * 40, 2, 20)
    app.addSource(0.5, 0.5, 0.5, 0, 0, 0)
    app.addReceiver(0.6, 0.6, 0.6)
    app.addSourceDataFloat([0.0, 1.0, 0.0], 3, 1)
    app.addSourceDataDouble([0.0, 1.0, 0.0], 3, 1)
    with pytest.raises(IndexError):
        app.addSourceDataFloat([0.0], 3, 1)
    if capi.device_count() == 0:
        with pytest.raises(RuntimeError):                # no CPU fallback behind the module either
            app.initializeDevices()
        with pytest.raises(RuntimeError):
            app.runSimulation()
    with pytest.raises(RuntimeError):
        app.initializeGeometryFromFile("x.vtk")                # unreadable file


def test_geometry_from_a_vtk_file(pf, tmp_path):
    """initializeGeometryFromFile: legacy ASCII VTK POLYDATA in inches (reference FileReader.cpp:41-101)"""
    inch = 1.0 / 0.0254
    pts = [(0, 0, 0), (2, 0, 0), (2, 1, 0), (0, 1, 0), (0, 0, 1.5), (2, 0, 1.5), (2, 1, 1.5), (0, 1, 1.5)]
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]
    tris = [t for a, b, c, d in quads for t in ((a, b, c), (a, c, d))]
    p = tmp_path / "room.vtk"
    p.write_text("# vtk DataFile Version 3.0\nroom\nASCII\nDATASET POLYDATA\nPOINTS 8 float\n"
                 + "".join(f"{x * inch} {y * inch} {z * inch}\n" for x, y, z in pts)
                 + f"POLYGONS {len(tris)} {4 * len(tris)}\n" + "".join(f"3 {a} {b} {c}\n" for a, b, c in tris))
    app = pf.App()
    app.initializeGeometryFromFile(str(p))
    app.setLayerIndices(list(range(len(tris))), "all")          # accepted: the file gave 12 triangles
    app.setUniformMaterial(0.9)
    (tmp_path / "quad.vtk").write_text("DATASET POLYDATA\nPOINTS 4 float\n0 0 0 1 0 0 1 1 0 0 1 0\nPOLYGONS 1 5\n4 0 1 2 3\n")
    with pytest.raises(RuntimeError):
        pf.App().initializeGeometryFromFile(str(tmp_path / "quad.vtk"))


def test_filter_materials_through_the_module(pf):
    """additions for the frequency-dependent boundaries: per-surface rows [b0..bN, a1..aN], order 1..4"""
    app = pf.App()
    app.addSurfaceFilters([0.1, 0.02, 0.01, -0.5, 0.1] * 3, 3, 2)
    with pytest.raises(IndexError):
        app.addSurfaceFilters([0.1, 0.02], 3, 2)            # list shorter than surfaces * (2 order + 1)
    with pytest.raises(IndexError):
        app.addSurfaceFilters([0.1] * 11, 1, 5)             # order > 4
    with pytest.raises(Exception):
        app.addSurfaceFilters([0.1, 0.02, -0.5], 1, 1)      # another order once surfaces exist
    other = pf.App()
    other.setUniformFilter([0.1, 0.02, 0.01], [-0.5, 0.1])
    with pytest.raises(IndexError):
        pf.App().setUniformFilter([0.1, 0.02, 0.01], [-0.5])


def test_example_workflow_fails_loudly_without_a_device(pf, capi, tmp_path):
    """examples/test_bench.py (the reference's python/testBench.py workflow) has no CPU path either"""
    import subprocess
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "test_bench.py"), "--steps", "10"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and "ParallelFDTD error" in r.stderr
