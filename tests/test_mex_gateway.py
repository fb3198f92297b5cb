"""The MATLAB gateway `mex_FDTD` (parallelfdtd_b200/host/mex_FDTD.cpp; calling convention of the reference's
matlab/mex_FDTD.cpp:30-368 as driven by matlab/runFDTD.m) compiled against the stand-in MEX API of
tests/cpp/mex_stub/mex.h and driven by tests/cpp/mex_tests.cpp.  `cpu`: argument checking and the loud failure
without a device; `gpu`: the 1 / 3 / 8 output forms, double precision, captures, DATA sources and filter materials
against the same jobs run directly on FDTD::App."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bin():
    from parallelfdtd_b200 import build
    b = build.build_mex_tests()
    assert os.path.exists(b)
    return b


def test_gateway_exports_mexfunction_with_c_linkage():
    syms = subprocess.run(["nm", _bin()], capture_output=True, text=True).stdout
    assert " T mexFunction" in syms


def test_mex_gateway_cpu(tmp_path):
    r = subprocess.run([_bin(), "cpu"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_mex_gateway_gpu(tmp_path, gpu):
    r = subprocess.run([_bin(), "gpu"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
