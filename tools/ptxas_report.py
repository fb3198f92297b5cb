"""Registers / stack / spills of every kernel in one .cu (nvcc -Xptxas -v, sm_100a).  usage: python tools/ptxas_report.py file.cu [filter]"""
import re, subprocess, sys
out = subprocess.run(["nvcc", "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xptxas", "-v", "-c", sys.argv[1],
                      "-o", "/tmp/_ptxas_report.o"], capture_output=True, text=True).stderr
flt = sys.argv[2] if len(sys.argv) > 2 else ""
name = None
for line in out.splitlines():
    m = re.search(r"Function properties for (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        stack = m.groups()
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and name and flt in name:
        print(f"{name[:110]:110s} regs {m.group(1):>3s} stack {stack[0]:>4s} spill st/ld {stack[1]:>4s}/{stack[2]:>4s}")
