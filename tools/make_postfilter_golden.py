"""Golden vectors for parallelfdtd_b200/postfilter.py from the reference's own Python post-filter
(/root/reference/python/FDTDfilter.py, imported here; it cannot travel to the GPU box, the vectors can):
    python tools/make_postfilter_golden.py  ->  tests/golden/postfilter_fdtdfilter.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/python")
from FDTDfilter import FDTDfilter  # noqa: E402  (the reference)

rng = np.random.default_rng(0x5EED0F17)
x1 = np.zeros(600)
x1[3] = 1.0                                                    # impulse: the output is the tap vector
x2 = rng.standard_normal((900, 3))
out = {}
for k, (x, sfs, cut) in enumerate([(x1, 100000.0, 0.2), (x2, 100000.0, 0.2), (x2, 48000.0, 0.05)]):
    out[f"x{k}"] = x
    out[f"arg{k}"] = np.array([sfs, cut])
    out[f"y{k}"] = FDTDfilter(x, sfs, 0, cut)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "postfilter_fdtdfilter.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
