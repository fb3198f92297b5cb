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


def test_runfdtd_helper_mex_functions():
    """mem_check / device_reset (matlab/runFDTD.m:19,42) through the stand-in MEX API, called via ctypes"""
    import ctypes as C
    from parallelfdtd_b200 import build, capi

    class MxArray(C.Structure):
        _fields_ = [("cls", C.c_int), ("m", C.c_size_t), ("n", C.c_size_t), ("data", C.c_void_p)]

    mem_so, reset_so = build.build_mex_helpers()
    plhs = (C.POINTER(MxArray) * 1)()
    L = C.CDLL(mem_so)
    L.mexFunction(C.c_int(1), plhs, C.c_int(0), None)
    out = plhs[0].contents
    ndev = capi.device_count()
    assert out.cls == 6 and out.n == 1 and out.m == ndev                      # mxDOUBLE_CLASS column, one entry per device
    if ndev:
        mem = (C.c_double * ndev).from_address(out.data)
        assert all(v > 1e9 for v in mem)
    C.CDLL(reset_so).mexFunction(C.c_int(0), None, C.c_int(0), None)
