"""CPU: the oracle against the golden vectors produced by the REFERENCE'S OWN CUDA kernels.

tests/golden/<case>.npz were written by tools/gpu_check.py on a B200: it runs oracle/_ref/ref_fdtd
(the reference's kernels3d.cu / cudaMesh.cu / cudaUtils.cu + host classes, compiled unmodified for
sm_100 by oracle/Makefile) on the cases of tests/fdtd_cases.parity_cases() and stores its receiver
responses, padded dims, node counts, partition index sets and node bytes.  This pins the oracle:
bit-exact responses (fp32 and fp64), bit-exact node bytes and partition layout.
"""
import os

import numpy as np
import pytest

from oracle import oracle
from tests import fdtd_cases as fc

CASES = {c["name"]: c for c in fc.parity_cases()}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_cuda_build(name):
    g = fc.load_golden(name)
    assert g is not None, f"tests/golden/{name}.npz missing"
    case = CASES[name]
    r, (pos, mat, air, bnd), _ = fc.run_oracle(case, double_pad=True)
    Z, Y, X = pos.shape
    assert (X, Y, Z) == tuple(int(v) for v in g["dims"])
    assert (air, bnd) == (int(g["n_air"]), int(g["n_boundary"]))
    first, size = oracle.partition_indexing(Z, case["n_parts"])
    assert [tuple(int(v) for v in p) for p in g["partitions"]] == list(zip(first, size))
    for k in range(case["n_parts"]):
        assert np.array_equal(pos[first[k]:first[k] + size[k]], g[f"pos_{k}"])
        assert np.array_equal(mat[first[k]:first[k] + size[k]], g[f"mat_{k}"])
        assert fc.node_checksum(g[f"pos_{k}"]) == int(g["pos_crc"][k])
        assert fc.node_checksum(g[f"mat_{k}"]) == int(g["mat_crc"][k])
    ref = g["responses"]
    assert ref.shape == r.shape and np.abs(ref).max() > 0
    assert np.array_equal(r, ref.astype(r.dtype)), f"rel-L2 {fc.rel_l2(r, ref):.3e}"


def test_double_padding_quirk_does_not_change_responses():
    # y padded to block.x (64) instead of block.y multiples (40): extra rows are solid
    case = CASES["shoebox_48x40x49_ctr_f64_6mat_5parts"]
    a, (pa, _, _, _), _ = fc.run_oracle(case, double_pad=True)
    b, (pb, _, _, _), _ = fc.run_oracle(case, double_pad=False)
    assert pa.shape == (49, 64, 64) and pb.shape == (49, 40, 64)
    assert np.array_equal(a, b)


def test_constructor_lambda_differs_in_double_only():
    case = dict(CASES["shoebox64_fwd_f64"])
    g = fc.load_golden(case["name"])
    lam_ctor = 1.0 / np.sqrt(3.0)
    assert lam_ctor != fc.LAM and np.float32(lam_ctor) == np.float32(fc.LAM)
    assert np.float32(lam_ctor * lam_ctor) == np.float32(fc.LAM * fc.LAM)
    saved = fc.LAM
    try:
        fc.LAM = lam_ctor
        r, _, _ = fc.run_oracle(case)
    finally:
        fc.LAM = saved
    assert not np.array_equal(r, g["responses"]) and fc.rel_l2(r, g["responses"]) < 1e-11


def test_golden_fixtures_are_consistent_with_their_case_definitions():
    """tools/make_golden.py --check-only: shapes, dtypes, the slab rule and the node checksums of every fixture."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "make_golden.py"), "--check-only"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(": ok") == len(fc.parity_cases())
