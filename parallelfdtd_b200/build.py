"""Build recipe for libpfdtd_b200.so (hand-written CUDA for sm_100a, no torch dependency).

``python -m parallelfdtd_b200.build`` or ``build_lib()``; the library is built in-tree so
that it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpfdtd_b200.so")
SOURCES = ["pfdtd_api.cu", "update_kernels.cu", "interp_kernels.cu", "mesh_kernels.cu", "srcrec_kernels.cu", "capture_kernels.cu", "voxelize_kernels.cu"]
HEADERS = ["pfdtd_internal.h", "update_math.cuh", "tma_common.cuh", "update_host.cuh", os.path.join("..", "..", "include", "pfdtd.h")]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(src: str, dst: str) -> bool:
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    hdr_paths = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(op)
        if force or _newer(sp, op) or any(_newer(h, op) for h in hdr_paths):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", op]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
        subprocess.check_call(cmd)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_LIB = os.path.join(HERE, "libpfdtd_host.so")
HOST_SOURCES = [os.path.join("base", "SimulationParameters.cpp"), os.path.join("base", "MaterialHandler.cpp"), "App.cpp"]
HOST_TEST_SRC = os.path.normpath(os.path.join(HERE, "..", "tests", "cpp", "host_tests.cpp"))
HOST_TEST_BIN = os.path.normpath(os.path.join(HERE, "..", "tests", "cpp", "host_tests"))


def _host_cxx() -> str:
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++")


def build_host(force: bool = False) -> str:
    """C++ host layer (reference class names over the C ABI) -> libpfdtd_host.so, plus its test binary."""
    build_lib()
    srcs = [os.path.join(HOST_DIR, s) for s in HOST_SOURCES]
    deps = []
    for root, _, files in os.walk(HOST_DIR):
        deps += [os.path.join(root, f) for f in files]
    deps.append(os.path.normpath(os.path.join(HERE, "..", "include", "pfdtd.h")))
    flags = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-unused-function", "-I", os.path.join(HERE, "host")]
    link = ["-L", HERE, "-l:libpfdtd_b200.so", "-Wl,-rpath,$ORIGIN"]
    if force or any(_newer(d, HOST_LIB) for d in deps):
        subprocess.check_call([_host_cxx()] + flags + ["-shared", "-o", HOST_LIB] + srcs + link)
    if os.path.exists(HOST_TEST_SRC) and (force or _newer(HOST_TEST_SRC, HOST_TEST_BIN) or _newer(HOST_LIB, HOST_TEST_BIN)):
        subprocess.check_call([_host_cxx()] + flags + ["-o", HOST_TEST_BIN, HOST_TEST_SRC, "-L", HERE, "-l:libpfdtd_host.so",
                                                       "-l:libpfdtd_b200.so", "-Wl,-rpath," + HERE])
    return HOST_LIB


MEX_SRC = os.path.join(HOST_DIR, "mex_FDTD.cpp")
MEX_STUB_DIR = os.path.normpath(os.path.join(HERE, "..", "tests", "cpp", "mex_stub"))
MEX_TEST_SRC = os.path.normpath(os.path.join(HERE, "..", "tests", "cpp", "mex_tests.cpp"))
MEX_TEST_BIN = os.path.normpath(os.path.join(HERE, "..", "tests", "cpp", "mex_tests"))


def build_mex_tests(force: bool = False) -> str:
    """The MATLAB gateway (host/mex_FDTD.cpp) compiled against the stand-in MEX API of tests/cpp/mex_stub (MATLAB is
    not in this image) together with its driver -> tests/cpp/mex_tests.  Inside MATLAB the same source builds with
    `mex` against the real mex.h."""
    build_host()
    deps = [MEX_SRC, MEX_TEST_SRC, os.path.join(MEX_STUB_DIR, "mex.h"), HOST_LIB, os.path.join(HOST_DIR, "App.h")]
    if force or any(_newer(d, MEX_TEST_BIN) for d in deps):
        subprocess.check_call([_host_cxx(), "-std=c++17", "-O2", "-Wall", "-Wno-unused-function", "-I", MEX_STUB_DIR, "-I", HOST_DIR,
                               "-o", MEX_TEST_BIN, MEX_SRC, MEX_TEST_SRC, "-L", HERE, "-l:libpfdtd_host.so", "-l:libpfdtd_b200.so",
                               "-Wl,-rpath," + HERE])
    return MEX_TEST_BIN


MEX_HELPERS = ["mem_check", "device_reset"]


def mex_helper_path(name: str) -> str:
    return os.path.normpath(os.path.join(HERE, "..", "tests", "cpp", f"mex_{name}.so"))


def build_mex_helpers(force: bool = False):
    """matlab/runFDTD.m's two helper MEX functions (host/mem_check.cpp, host/device_reset.cpp) against the stand-in mex.h,
    as shared objects the tests call through ctypes."""
    build_lib()
    out = []
    for name in MEX_HELPERS:
        src, dst = os.path.join(HOST_DIR, name + ".cpp"), mex_helper_path(name)
        if force or _newer(src, dst) or _newer(os.path.join(MEX_STUB_DIR, "mex.h"), dst):
            subprocess.check_call([_host_cxx(), "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", MEX_STUB_DIR, "-o", dst, src,
                                   "-L", HERE, "-l:libpfdtd_b200.so", "-Wl,-rpath," + HERE])
        out.append(dst)
    return out


PY_MODULE_SRC = os.path.join(HOST_DIR, "AppPy.cpp")


def py_module_path() -> str:
    import sysconfig
    return os.path.join(HERE, "libPyFDTD" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_py_module(force: bool = False) -> str:
    """The Python module `libPyFDTD` (API of the reference's boost::python module, src/AppPy.cpp) -> in-tree
    CPython extension linked against libpfdtd_host.so / libpfdtd_b200.so."""
    import sysconfig
    import pybind11
    build_host()
    out = py_module_path()
    deps = [PY_MODULE_SRC, HOST_LIB, os.path.join(HOST_DIR, "App.h")]
    if force or any(_newer(d, out) for d in deps):
        subprocess.check_call([_host_cxx(), "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
                               "-I", HOST_DIR, "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
                               "-o", out, PY_MODULE_SRC, "-L", HERE, "-l:libpfdtd_host.so", "-l:libpfdtd_b200.so",
                               "-Wl,-rpath,$ORIGIN"])
    return out


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
    print(build_py_module(force="--force" in sys.argv))
    print(build_mex_tests(force="--force" in sys.argv))
    print(build_mex_helpers(force="--force" in sys.argv))
