"""GPU: parity AT SIZE.  BASELINE config 2 (512^3 shoebox, 6 wall materials) at its full size and, for the part the
reference implements, its full length:

  * frequency-independent boundaries, forward (fp32, fp64) and centred (fp32) schemes: 2000 steps against the
    reference's own CUDA build run LIVE on the same box (oracle/_ref/ref_fdtd: the reference's kernels3d.cu /
    cudaMesh.cu compiled unmodified) -- tolerance of BASELINE.json's north_star: receiver responses within relative L2
    1e-5 (fp32) / 1e-12 (fp64) over the full run; asserted stricter: bit for bit.
  * what the reference does not contain (order-2 digital impedance filter boundaries, the 27-point IISO scheme and their
    combination): 100 steps against the CPU oracle (oracle/fdtd_oracle.cpp) at 512^3, bit for bit, with the source in a
    corner of the room so that walls, edges and corners are all excited inside the run.
  * one slab of more than 2^31 voxels (1024 x 1024 x 2112): TMA kernel vs plain kernel with the source, the receivers and
    a captured plane all beyond element 2^31 (the reference indexes with 32-bit ints, cudaMesh.h:142).
"""
import numpy as np
import pytest

from oracle import casefile, oracle
from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu

DIMS = (512, 512, 512)
TOL = {False: 1e-5, True: 1e-12}       # north_star: relative L2 of the receiver responses over the full run


def _pulse(n, t0=40.0, w=6.0):
    k = np.arange(n, dtype=np.float64)
    return np.exp(-0.5 * ((k - t0) / w) ** 2)


def _gpu_mem_gb(capi):
    import ctypes as C
    tot, free = C.c_int(0), C.c_int(0)
    capi.lib().pfdtd_device_mem_mb(0, C.byref(tot), C.byref(free))
    return free.value / 1024.0


@pytest.mark.parametrize("update_type,double", [(0, False), (0, True), (2, False)])
def test_config2_full_size_full_length_against_the_live_reference_build(capi, gpu, tmp_path, update_type, double):
    if not casefile.ref_available():
        pytest.skip("oracle/_ref/ref_fdtd (the reference's CUDA build) is not in this snapshot")
    steps = 2000
    case = fc.make_case(f"c2_512_ut{update_type}_{'f64' if double else 'f32'}", DIMS, update_type, double, steps, 6, 1,
                        [(256, 256, 256, 0, 3, 0)], [(273, 261, 259), (30, 40, 50), (256, 256, 3), (500, 300, 200)],
                        input_data=[_pulse(steps)])
    ref = casefile.run_reference(case, str(tmp_path), timeout=1200)
    ours, _, info = fc.run_ours(capi, case)
    assert "tma" in info["kernel"], info["kernel"]
    assert tuple(info["dims"]) == tuple(ref["dims"])
    assert info["counts"][1] == ref["n_air"] and info["counts"][2] == ref["n_boundary"]
    r = ref["responses"]
    assert np.abs(r).max() > 0 and all(np.abs(r[i]).max() > 0 for i in range(r.shape[0])), "every receiver must be reached"
    err = fc.rel_l2(ours, r)
    assert err <= TOL[double], err
    assert np.array_equal(ours, r), (err, float(np.abs(ours.astype(np.float64) - r).max()))


# (update_type, double, dif_order)
UNPINNED = [(0, False, 2), (0, True, 2), (2, False, 2), (3, False, 0), (3, False, 2), (3, True, 2)]


@pytest.mark.parametrize("update_type,double,order", UNPINNED)
def test_config2_full_size_filter_and_interpolated_variants_against_the_cpu_oracle(capi, gpu, update_type, double, order):
    steps = 100
    src = [(24, 20, 28, 0, 3, 0)]                                                            # 100 steps reach ~57 voxels
    rec = [(5, 20, 28), (24, 2, 30), (30, 26, 1), (1, 1, 1), (50, 40, 45), (2, 14, 2)]       # faces, a corner, open air, an edge
    case = fc.make_case(f"c2_512_ut{update_type}_dif{order}", DIMS, update_type, double, steps, 6, 1, src, rec, input_data=[_pulse(steps, 12.0, 3.0)])
    if order:
        case["dif_order"] = order
        case["materials"] = fc.dif_table(6, order).astype(np.float64 if double else np.float32)
    want, _, _ = fc.run_oracle(case, matidx=0)
    got, _, info = fc.run_ours(capi, case, matidx=0)
    assert "tma" in info["kernel"], info["kernel"]
    assert all(np.abs(want[i]).max() > 0 for i in range(want.shape[0])), "every receiver must be reached"
    assert np.array_equal(got, want), (info["kernel"], fc.rel_l2(got, want))


def test_single_slab_beyond_2_to_31_voxels(capi, gpu):
    dims = (1024, 1024, 2112)                       # 2.21e9 voxels: fields 2 x 8.9 GB, node volumes 3 x 2.2 GB
    if _gpu_mem_gb(capi) < 40:
        pytest.skip("needs 40 GB of free device memory")
    steps = 48
    bid, mat = synth.shoebox(dims, 6)
    tab = synth.material_table(list(np.linspace(0.99, 0.5, 6)))
    prm = oracle.params(fc.LAM, 0)
    zs = 2100                                       # element index of (x, y, zs) > 2^31 for every x, y
    assert zs * dims[0] * dims[1] > 2 ** 31
    src = [[500, 520, zs]]
    rec = [[512, 512, zs + 3], [490, 530, 2109], [500, 520, 2080]]
    out = {}
    for kern in (capi.KERNEL_PLAIN, capi.KERNEL_TMA):
        s = capi.Solver()
        try:
            s.set_option(capi.OPT_KERNEL, kern)
            s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
            s.setup_mesh(bid, mat, (32, 4, 1), capi.SRL_FORWARD, capi.F32, prm, tab)
            s.make_partition(1, [0])
            s.set_sources(src, [capi.SRC_HARD], _pulse(steps, 10.0, 3.0)[None, :])
            s.set_receivers(rec)
            resp, _ = s.run(steps)
            plane = s.capture_slice(2105, 0)
            assert s.element_idx_and_partition(500, 520, zs)[1] > 2 ** 31
            assert s.get_sample(500, 520, zs) == plane_value(s, 500, 520, zs)
            out[kern] = (resp, plane, s.kernel_name())
        finally:
            s.close()
    (r0, p0, n0), (r1, p1, n1) = out[capi.KERNEL_PLAIN], out[capi.KERNEL_TMA]
    assert "plain" in n0 and "tma" in n1
    assert all(np.abs(r0[i]).max() > 0 for i in range(3)) and np.abs(p0).max() > 0
    assert np.array_equal(r0, r1) and np.array_equal(p0, p1)


def plane_value(s, x, y, z):
    """the same voxel through the slice capture (a second, independent 64-bit addressing path)"""
    return float(s.capture_slice(z, 0)[y, x])
