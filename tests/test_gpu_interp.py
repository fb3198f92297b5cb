"""GPU: the 27-point interpolated schemes (IISO / IWB) through the C ABI against the CPU oracle.
The reference does not contain these schemes (parity unpinned by it); the oracle restates the
equation documented in csrc/update_math.cuh operation by operation, so results are compared bit for bit."""
import numpy as np
import pytest

from oracle import oracle
from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu
CASES = {c["name"]: c for c in fc.interp_cases()}
TOL = {False: 1e-5, True: 1e-12}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("kernel", ["tma", "plain"])
def test_interp_parity_with_oracle(capi, gpu, name, kernel):
    case = CASES[name]
    kern = capi.KERNEL_TMA if kernel == "tma" else capi.KERNEL_PLAIN
    r_or, (pos, mat, air, bnd), _ = fc.run_oracle(case)
    r, nodes, info = fc.run_ours(capi, case, kernel=kern, matidx=0)
    assert "interp" in info["kernel"] and kernel in info["kernel"]
    f_, s_ = oracle.partition_indexing(pos.shape[0], case["n_parts"])
    for k in range(case["n_parts"]):
        assert np.array_equal(nodes[k][0], pos[f_[k]:f_[k] + s_[k]]) and np.array_equal(nodes[k][1], mat[f_[k]:f_[k] + s_[k]])
    assert np.abs(r_or).max() > 0
    assert fc.rel_l2(r, r_or) <= TOL[case["double"]]
    assert np.array_equal(r, r_or)


@pytest.mark.parametrize("name", list(CASES))
def test_interp_partition_invariance_and_tiles(capi, gpu, name):
    case = CASES[name]
    base, _, _ = fc.run_ours(capi, case, n_parts=1, matidx=0)
    for n in (2, 5):
        r, _, _ = fc.run_ours(capi, case, n_parts=n, matidx=0)
        assert np.array_equal(r, base), n
    for tile in (1, 3, 7, 9):
        for chunk in (0, 1, 7):
            r, _, info = fc.run_ours(capi, case, n_parts=1, matidx=0, kernel=capi.KERNEL_TMA,
                                     opts=[(capi.OPT_TMA_TILE, tile), (capi.OPT_TMA_CHUNK, chunk)])
            assert np.array_equal(r, base), info["kernel"]


def test_interp_kernel_with_srl_weights_matches_the_pinned_scheme(capi, gpu):
    """(d1,d2,d3,d4) = (lam2, 0, 0, 2-6 lam2): the 27-point kernel computes the reference's SRL_FORWARD equation."""
    srl = {c["name"]: c for c in fc.parity_cases()}["hall_96x128x64_fwd_f32_5mat_oct1"]
    r_srl, _, _ = fc.run_ours(capi, srl, matidx=0)
    case = dict(srl, update_type=3)
    lam2 = np.float32(fc.LAM * fc.LAM)
    case["dcoef"] = [float(lam2), 0.0, 0.0, float(2 - 6 * np.float64(lam2))]
    # run at the SRL Courant number: the parameter vector is built from fc.LAM by overriding lam_of
    saved = oracle.interp_lambda
    try:
        oracle.interp_lambda = lambda ut: fc.LAM
        r_int, _, info = fc.run_ours(capi, case, matidx=0)
        r_or, _, _ = fc.run_oracle(case)
    finally:
        oracle.interp_lambda = saved
    assert "interp" in info["kernel"]
    assert np.array_equal(r_int, r_or)
    assert fc.rel_l2(r_int, r_srl) < 2e-5
