"""The drop-in boundary, checked with the reference's OWN front ends: matlab/mex_FDTD.cpp (against the stand-in MEX
API of tests/cpp/mex_stub/mex.h) and src/main.cpp are compiled UNCHANGED, where they lie under /root/reference,
against parallelfdtd_b200/host/ and linked with libpfdtd_host.so / libpfdtd_b200.so.  Nothing is run (main.cpp opens
data files and waits on stdin); what is checked is that every member of FDTD::App / CudaMesh / SimulationParameters /
MaterialHandler those sources use exists here with a compatible signature.

Only meaningful where the reference tree is mounted (this container); skipped on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HOST = os.path.join(ROOT, "parallelfdtd_b200", "host")
PKG = os.path.join(ROOT, "parallelfdtd_b200")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not mounted")


def _libs():
    from parallelfdtd_b200 import build
    build.build_host()
    return ["-L", PKG, "-l:libpfdtd_host.so", "-l:libpfdtd_b200.so", "-Wl,-rpath," + PKG]


def test_reference_mex_gateway_compiles_and_links_unchanged(tmp_path):
    # mex_includes.h (next to mex_FDTD.cpp) asks for "./includes/App.h": the layout the reference's install step creates
    os.symlink(HOST, tmp_path / "includes")
    out = tmp_path / "mex_FDTD_ref.so"
    cmd = [CXX, "-std=c++17", "-O0", "-fPIC", "-shared", "-w", "-I", os.path.join(ROOT, "tests", "cpp", "mex_stub"), "-I", str(tmp_path),
           "-I", HOST, "-o", str(out), os.path.join(REF, "matlab", "mex_FDTD.cpp")] + _libs()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-4000:]
    syms = subprocess.run(["nm", "-D", str(out)], capture_output=True, text=True).stdout
    assert " T mexFunction" in syms
    for needed in ("runVisualization", "runSimulation", "runCapture", "addSliceToCapture", "addMeshToCapture"):
        assert needed in syms, needed                               # undefined here = resolved by libpfdtd_host.so
    undefined = subprocess.run(["ldd", "-r", str(out)], capture_output=True, text=True)
    missing = [l for l in (undefined.stdout + undefined.stderr).splitlines() if "undefined symbol" in l and "mx" not in l and "mex" not in l]
    assert not missing, missing


def test_reference_main_compiles_and_links_unchanged(tmp_path):
    # fed through stdin so that `#include "App.h"` resolves through -I (a file compiled in place would pick up the
    # reference's own src/App.h, which sits next to it)
    out = tmp_path / "ref_main"
    src = open(os.path.join(REF, "src", "main.cpp"), "rb").read()
    cmd = [CXX, "-std=c++17", "-O0", "-w", "-x", "c++", "-", "-I", HOST, "-o", str(out)] + _libs()
    r = subprocess.run(cmd, input=src, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-4000:]
    assert os.path.exists(out)
