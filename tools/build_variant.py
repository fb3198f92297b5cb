"""A/B builds of libpfdtd_b200.so with extra -D flags -> build/variants/libpfdtd_b200_<name>.so (git-ignored, travels with gpurun).
Select at run time with PFDTD_LIB_PATH=<that file>.   usage: python tools/build_variant.py <name> [-DX=Y ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelfdtd_b200 import build as b  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "build", "variants", name)
os.makedirs(out_dir, exist_ok=True)
procs, objs = [], []
for src in b.SOURCES:
    o = os.path.join(out_dir, src.replace(".cu", ".o"))
    objs.append(o)
    procs.append(subprocess.Popen(["nvcc"] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + defs + ["-c", os.path.join(b.CSRC, src), "-o", o]))
if any(p.wait() for p in procs):
    raise SystemExit("nvcc failed")
lib = os.path.join(ROOT, "build", "variants", f"libpfdtd_b200_{name}.so")
subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"])
print(lib)
