"""GPU: the five BASELINE.json configurations, each at the scale the CPU oracle finishes in seconds (SURVEY 8d says
which reduced copy stands for which configuration), through the C ABI against the oracle, bit for bit.

  C1  64^3 shoebox, SRL_FORWARD, one admittance, impulse source, 1 receiver, 500 steps, fp32      (full size)
  C2  shoebox, 6 materials, order-2 DIF boundaries, fp32 and fp64                                  (128^3 for 512^3)
  C3  IISO on the synthetic hall, 5 materials, 16 receivers on a seating grid, 8 z-slabs           (1/8 scale)
  C4  weak-scaling shoebox, SRL fp32: slab count must not change the responses                      (8 slabs of 24 planes)
  C5  fp64 IISO, 20 materials in z-bands, 10 source positions as 10 separate runs                   (1/16 scale)
"""
import numpy as np
import pytest

from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu


def _case(name, bid, mat, update_type, double, steps, n_mat, n_parts, sources, receivers, dif_order=None):
    refl = [0.9] if n_mat == 1 else list(np.linspace(0.99, 0.5, n_mat))
    npdt = np.float64 if double else np.float32
    tab = fc.dif_table(n_mat, dif_order).astype(npdt) if dif_order else synth.material_table(refl).astype(np.float32)
    c = dict(name=name, bid=bid, mat=mat, block=(32, 4, 1), update_type=update_type, double=double, steps=steps, octave=0,
             n_parts=n_parts, devices=[0] * n_parts, materials=tab, sources=list(sources), receivers=list(receivers), input_data=[])
    if dif_order:
        c["dif_order"] = dif_order
    return c


def _check(capi, case, matidx=0):
    r_or, (pos, mat, _, _), _ = fc.run_oracle(case, matidx=matidx)
    r, nodes, info = fc.run_ours(capi, case, matidx=matidx)
    assert "tma" in info["kernel"], info["kernel"]
    assert np.abs(r_or).max() > 0
    assert fc.rel_l2(r, r_or) <= (1e-12 if case["double"] else 1e-5)
    assert np.array_equal(r, r_or), info["kernel"]
    return r, info


def test_c1_shoebox64_forward_fp32_500_steps(capi, gpu):
    bid, mat = synth.shoebox((64, 64, 64), 1)
    c = _case("c1", bid, mat, 0, False, 500, 1, 1, [(32, 32, 32, 0, 0, 0)], [(40, 36, 28)])
    c["materials"] = synth.material_table([0.9]).astype(np.float32)
    _check(capi, c, matidx=1)


@pytest.mark.parametrize("double", [False, True])
def test_c2_shoebox_six_materials_dif2(capi, gpu, double):
    dims = (128, 128, 128)
    bid, mat = synth.shoebox(dims, 6)
    rec = [(64 + 17, 64 + 5, 3 + (i * (128 - 6)) // 3) for i in range(4)]
    c = _case("c2", bid, mat, 0, double, 160, 6, 1, [(64, 64, 64, 0, 1, 0)], rec, dif_order=2)
    _check(capi, c)


def test_c3_iiso_hall_16_receivers_8_slabs(capi, gpu):
    dims = (192, 128, 120)                                   # 1/8 scale of 1536 x 1024 x 960
    bid, mat = synth.hall(dims, 5)
    rec = [(40 + 16 * i, 66 + 14 * j, 34 + 6 * i) for i in range(8) for j in range(2)]   # seating grid over the raked floor, all 8 slabs
    c = _case("c3", bid, mat, 3, False, 140, 5, 8, [(96, 60, 60, 0, 0, 0)], rec)
    r8, info = _check(capi, c)
    assert (np.abs(r8).max(axis=1) > 0).all()                # the wave reached every receiver
    assert [p[1] for p in info["partitions"]] == [16, 17, 17, 17, 17, 17, 17, 16]     # getPartitionIndexing of 120 slices over 8
    r1, _, _ = fc.run_ours(capi, c, n_parts=1, matidx=0)
    assert np.array_equal(r1, r8)


def test_c4_weak_scaling_slabs_do_not_change_the_answer(capi, gpu):
    dims = (256, 128, 192)                                   # 8 slabs of 24 planes, like 8 x 960 planes at full size
    bid, mat = synth.shoebox(dims, 6)
    rec = [(130, 70, 5 + 26 * i) for i in range(8)]          # one receiver per slab
    c = _case("c4", bid, mat, 0, False, 120, 6, 8, [(128, 64, 96, 0, 1, 0)], rec)
    r8, _ = _check(capi, c)
    for n in (1, 2, 4):
        r, _, _ = fc.run_ours(capi, c, n_parts=n, matidx=0)
        assert np.array_equal(r, r8), n


def test_c5_fp64_iiso_20_materials_10_sources_as_separate_runs(capi, gpu):
    dims = (128, 64, 120)                                    # 1/16 scale of 2048 x 1024 x 1920
    bid, mat = synth.banded_shoebox(dims, 20)
    rng = np.random.default_rng(5)
    srcs = [(int(rng.integers(30, 98)), int(rng.integers(16, 48)), int(rng.integers(30, 90))) for _ in range(10)]
    rec = [(64, 32, 60), (20, 50, 100)]
    seen = set()
    for k, (x, y, z) in enumerate(srcs):
        c = _case(f"c5_{k}", bid, mat, 3, True, 130, 20, 8, [(x, y, z, 0, 0, 0)], rec)
        r, _ = _check(capi, c)
        assert (np.abs(r).max(axis=1) > 0).all()
        seen.add(r.tobytes())
    assert len(seen) == 10                                   # ten different responses


# More (position, material) pairs than a class byte can name -- or more lossy ones than the kernels' shared filter table
# holds: the mesh is classified by position alone and boundary voxels look their material up (update_math.cuh WideArgs).
# The reference has no such ceiling (separate material byte, cudaMesh.h:90-91, kernels3d.cu:639-640).
@pytest.mark.parametrize("update_type,double,order,n_parts", [(2, False, 2, 1), (3, True, 2, 2), (2, True, 0, 3), (3, False, 0, 1), (4, False, 1, 2)])
def test_c5_hall_20_materials_beyond_256_node_classes(capi, gpu, update_type, double, order, n_parts):
    dims = (96, 128, 64)
    bid, mat = synth.hall(dims, 20)
    # a material per boundary voxel instead of per surface: every (position, material) pair occurs
    rng = np.random.default_rng(11)
    mat = np.where((bid > 0) & (bid < 27), rng.integers(0, 20, size=bid.shape, dtype=np.uint8), 0).astype(np.uint8)
    c = _case(f"wide_{update_type}_{order}", bid, mat, update_type, double, 150, 20, n_parts, [(40, 20, 20, 0, 0, 0)],
              [(50, 60, 30), (20, 100, 40), (3, 64, 30)], dif_order=order or None)
    r, info = _check(capi, c)
    if update_type == 2:      # 26 position bytes x 20 materials = 384 pairs on this mesh; the interpolated cases go wide when their
        assert "position classes" in info["kernel"], info["kernel"]      # K12 / K8 combinations or lossy-class count ask for it
    assert (np.abs(r).max(axis=1) > 0).all()
