"""Turn the outputs of the reference itself into the committed fixtures.

tools/gpu_check.py, run on a GPU box (`gpurun -- python tools/gpu_check.py`), drives oracle/_ref/ref_fdtd -- the
reference's own CUDA sources, compiled unmodified -- through the cases of tests/fdtd_cases.parity_cases() and writes
gpurun_out/golden/<case>.npz (responses, padded dims, node counts, slab index sets, node bytes per slab and their
checksums).  This script checks each file against the case it claims to be (shapes, slab rule, checksums) and copies
it to tests/golden/<case>.npz, which tests/test_oracle_golden.py (CPU) and tests/test_gpu_parity.py (GPU) read.

    python tools/make_golden.py [--src gpurun_out/golden] [--check-only]
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from tests import fdtd_cases as fc  # noqa: E402


def check(case, g):
    """-> list of problems (empty = consistent with the case definition)."""
    bad = []
    n_rec = len(case["receivers"])
    if g["responses"].shape != (n_rec, case["steps"]):
        bad.append(f"responses {g['responses'].shape} != ({n_rec}, {case['steps']})")
    want_dt = np.float64 if case["double"] else np.float32
    if g["responses"].dtype != want_dt:
        bad.append(f"responses dtype {g['responses'].dtype}")
    X, Y, Z = (int(v) for v in g["dims"])
    first, size = oracle.partition_indexing(Z, case["n_parts"])
    parts = [tuple(int(v) for v in p) for p in g["partitions"]]
    if parts != list(zip(first, size)):
        bad.append(f"slabs {parts} != rule {list(zip(first, size))}")
    for k, (f, s) in enumerate(parts):
        pos, mat = g[f"pos_{k}"], g[f"mat_{k}"]
        if pos.shape != (s, Y, X) or mat.shape != (s, Y, X):
            bad.append(f"slab {k} node shape {pos.shape}")
        if fc.node_checksum(pos) != int(g["pos_crc"][k]) or fc.node_checksum(mat) != int(g["mat_crc"][k]):
            bad.append(f"slab {k} checksum")
    if not (np.isfinite(g["responses"]).all() and np.abs(g["responses"]).max() > 0):
        bad.append("responses not finite / all zero")
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.path.join(ROOT, "gpurun_out", "golden"))
    ap.add_argument("--check-only", action="store_true", help="verify the fixtures already in tests/golden")
    a = ap.parse_args()
    src = fc.GOLDEN_DIR if a.check_only else a.src
    rc = 0
    for case in fc.parity_cases():
        p = os.path.join(src, case["name"] + ".npz")
        if not os.path.exists(p):
            print(f"{case['name']}: no file under {src}")
            rc = 1
            continue
        bad = check(case, dict(np.load(p)))
        print(f"{case['name']}: {'ok' if not bad else '; '.join(bad)}")
        if bad:
            rc = 1
        elif not a.check_only:
            shutil.copyfile(p, os.path.join(fc.GOLDEN_DIR, case["name"] + ".npz"))
    return rc


if __name__ == "__main__":
    sys.exit(main())
