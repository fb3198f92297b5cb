"""The C++ host layer (parallelfdtd_b200/host: reference class names over the C ABI) against the reference's
own unit tests, re-expressed in tests/cpp/host_tests.cpp.  `cpu` part: SimulationParameters / SrcRec /
MaterialHandler / partition indexing / geometry + voxelizer.  `gpu` part: CudaMesh set/get/halo semantics,
launchFDTD3d[Double] partition invariance, FDTD::App runSimulation vs runCapture."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "host_tests")


def _bin():
    from parallelfdtd_b200 import build
    build.build_host()
    assert os.path.exists(BIN)
    return BIN


def test_host_layer_cpu(tmp_path):
    r = subprocess.run([_bin(), "cpu"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_host_library_exports_reference_class_names():
    _bin()
    so = os.path.join(ROOT, "parallelfdtd_b200", "libpfdtd_host.so")
    syms = subprocess.run(["nm", "-DC", so], capture_output=True, text=True).stdout
    for name in ("FDTD::App::runSimulation()", "FDTD::App::initializeMesh(unsigned int)", "FDTD::App::executeStep()",
                 "SimulationParameters::getSourceSample(unsigned int, unsigned int)", "SimulationParameters::getParameterPtrDouble()",
                 "MaterialHandler::getMaterialCoefficientPtr()", "MaterialHandler::addMaterials(float*, unsigned int, unsigned int)"):
        assert name in syms, name


@pytest.mark.gpu
def test_host_layer_gpu(tmp_path, gpu):
    r = subprocess.run([_bin(), "gpu"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
