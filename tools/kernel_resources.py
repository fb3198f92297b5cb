"""Registers / stack / spill instructions of every TMA update kernel in the built objects (cuobjdump, no GPU needed).
usage: python tools/kernel_resources.py [object ...]   (default: the in-tree csrc/*.o)"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sys.argv[1:] or [os.path.join(ROOT, "parallelfdtd_b200", "csrc", n) for n in ("update_kernels.o", "interp_kernels.o")]
for o in objs:
    txt = subprocess.run(f"cuobjdump --dump-resource-usage {o} | c++filt", shell=True, capture_output=True, text=True).stdout
    name = None
    for line in txt.splitlines():
        m = re.search(r"Function void pfdtd::(\w+<[^>]*>)", line)
        if m:
            name = m.group(1).replace(" ", "")
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            print(f"{name:70s} reg {m.group(1):>3s} stack {m.group(2):>4s} smem {m.group(3)}")
            name = None
