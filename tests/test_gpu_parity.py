"""GPU: the CUDA path (through the C ABI) against the CPU oracle and against the golden vectors
the reference's own CUDA build produced (tests/golden/, made by tools/gpu_check.py on a B200).

Tolerances (BASELINE.json north_star): node bytes and partition layout bit-exact; receiver
responses rel-L2 <= 1e-5 (fp32) / 1e-12 (fp64).  The kernels write every rounding step explicitly,
so the stronger property -- bit-exact responses -- is asserted as well.
"""
import numpy as np
import pytest

from oracle import oracle
from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu

CASES = {c["name"]: c for c in fc.parity_cases()}
TOL = {False: 1e-5, True: 1e-12}


def _kernels(capi):
    return [("tma", capi.KERNEL_TMA), ("plain", capi.KERNEL_PLAIN)]


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("kernel", ["tma", "plain"])
def test_parity_with_oracle(capi, gpu, name, kernel):
    case = CASES[name]
    kern = dict(_kernels(capi))[kernel]
    r_or, (pos, mat, air, bnd), _ = fc.run_oracle(case)
    r, nodes, info = fc.run_ours(capi, case, kernel=kern)
    assert kernel in info["kernel"]
    # layout: padded dims, node counts, partition index sets, node bytes -- all bit-exact
    Z, Y, X = pos.shape
    assert info["dims"] == (X, Y, Z)
    assert info["counts"] == (X * Y * Z, air, bnd)
    f_, s_ = oracle.partition_indexing(Z, case["n_parts"])
    assert [tuple(p) for p in info["partitions"]] == list(zip(f_, s_))
    for k in range(case["n_parts"]):
        assert np.array_equal(nodes[k][0], pos[f_[k]:f_[k] + s_[k]])
        assert np.array_equal(nodes[k][1], mat[f_[k]:f_[k] + s_[k]])
    # responses
    assert np.linalg.norm(r_or) > 0
    assert fc.rel_l2(r, r_or) <= TOL[case["double"]]
    assert np.array_equal(r, r_or), "response is within tolerance but not bit-identical to the oracle"


@pytest.mark.parametrize("name", list(CASES))
def test_parity_with_reference_golden(capi, gpu, name):
    """Golden vectors = outputs of the reference's own kernels (oracle/_ref/ref_fdtd) on a B200."""
    g = fc.load_golden(name)
    if g is None:
        pytest.fail(f"tests/golden/{name}.npz is missing: regenerate with tools/gpu_check.py + tools/make_golden.py")
    case = CASES[name]
    # the reference pads y with block.x in double precision (setupMeshDouble, cudaMesh.cu:106-111)
    r, nodes, info = fc.run_ours(capi, case, opts=[(capi.OPT_DOUBLE_PAD_AS_WRITTEN, 1)])
    assert tuple(int(v) for v in g["dims"]) == info["dims"]
    assert (int(g["n_air"]), int(g["n_boundary"])) == info["counts"][1:]
    assert [tuple(int(v) for v in p) for p in g["partitions"]] == [tuple(p) for p in info["partitions"]]
    for k in range(case["n_parts"]):
        assert np.array_equal(nodes[k][0], g[f"pos_{k}"]) and np.array_equal(nodes[k][1], g[f"mat_{k}"])
    ref = g["responses"]
    assert fc.rel_l2(r, ref) <= TOL[case["double"]]
    assert np.array_equal(r, ref.astype(r.dtype)), "not bit-identical to the reference CUDA build"


@pytest.mark.parametrize("name", ["shoebox_48x40x49_ctr_f32_6mat_2parts", "shoebox_48x40x49_fwd_f32_6mat_2parts",
                                  "shoebox_48x40x49_ctr_f64_6mat_5parts", "hall_96x128x64_ctr_f32_5mat"])
@pytest.mark.parametrize("kernel", ["tma", "plain"])
def test_partition_count_does_not_change_results(capi, gpu, name, kernel):
    """reference tests/CudaMeshTest.cpp:472-575: 1, 2 and 5 partitions give bitwise equal responses."""
    case = CASES[name]
    kern = dict(_kernels(capi))[kernel]
    base, _, _ = fc.run_ours(capi, case, n_parts=1, kernel=kern)
    assert np.abs(base).max() > 0
    for n in (2, 3, 5, 8):
        r, _, info = fc.run_ours(capi, case, n_parts=n, kernel=kern)
        assert np.array_equal(r, base), (n, info["kernel"])
    # overlap off / graph off must not matter either
    r, _, _ = fc.run_ours(capi, case, n_parts=2, kernel=kern, opts=[(capi.OPT_OVERLAP, 0), (capi.OPT_USE_GRAPH, 0)])
    assert np.array_equal(r, base)


def test_tma_and_plain_kernels_agree_on_fields(capi, gpu):
    """Whole pressure field after 40 steps, not only receiver points."""
    case = CASES["hall_96x128x64_fwd_f32_5mat_oct1"]
    fields = []
    for kern in (capi.KERNEL_TMA, capi.KERNEL_PLAIN):
        s = capi.Solver()
        s.set_option(capi.OPT_KERNEL, kern)
        s.setup_mesh(case["bid"], case["mat"], case["block"], 0, capi.F32, oracle.params(fc.LAM, 1), case["materials"])
        s.make_partition(1, [0])
        src = np.asarray(case["sources"], dtype=np.int32).reshape(-1, 6)
        s.set_sources(src[:, :3], src[:, 3], fc.source_table(case))
        s.set_receivers(case["receivers"])
        s.enqueue_steps(0, 40)
        s.sync()
        fields.append((s.export_partition_pressure(0, 0), s.export_partition_pressure(0, 1)))
        s.close()
    assert np.abs(fields[0][0]).max() > 0
    assert np.array_equal(fields[0][0], fields[1][0]) and np.array_equal(fields[0][1], fields[1][1])


@pytest.mark.parametrize("tile", [1, 2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("double", [False, True])
def test_every_tma_tile_variant(capi, gpu, tile, double):
    case = CASES["shoebox_48x40x49_ctr_f64_6mat_5parts" if double else "shoebox_48x40x49_fwd_f32_6mat_2parts"]
    r_or, _, _ = fc.run_oracle(case)
    for chunk in (0, 1, 7):
        r, _, info = fc.run_ours(capi, case, n_parts=1, kernel=capi.KERNEL_TMA,
                                 opts=[(capi.OPT_TMA_TILE, tile), (capi.OPT_TMA_CHUNK, chunk)])
        assert np.array_equal(r, r_or), info["kernel"]


def test_matidx_intended_mode(capi, gpu):
    case = CASES["hall_96x128x64_fwd_f32_5mat_oct1"]
    a_or, _, _ = fc.run_oracle(case, matidx=0)
    b_or, _, _ = fc.run_oracle(case, matidx=1)
    assert not np.array_equal(a_or, b_or)
    a, _, _ = fc.run_ours(capi, case, matidx=0)
    assert np.array_equal(a, a_or)


def test_soft_source_accumulate_option(capi, gpu):
    case = dict(CASES["shoebox_48x40x49_fwd_f32_6mat_2parts"])
    r_or, _, _ = fc.run_oracle(case, soft=1)
    r, _, _ = fc.run_ours(capi, case, opts=[(capi.OPT_SOFT_ACCUMULATE, 1)])
    assert np.array_equal(r, r_or)
    r0_or, _, _ = fc.run_oracle(case, soft=0)
    assert not np.array_equal(r_or, r0_or)


def test_odd_dimensions_fall_back_to_plain_kernel_and_match(capi, gpu):
    # block (8,4,1) gives X = 24: not a multiple of 16 -> the TMA path is not eligible
    bid, mat = synth.shoebox((22, 18, 21), 6)
    tab = synth.material_table(list(np.linspace(0.99, 0.5, 6)))
    case = dict(name="odd", bid=bid, mat=mat, block=(8, 4, 1), update_type=2, double=False, steps=100, octave=0, n_parts=3,
                devices=[0, 0, 0], materials=tab, sources=[(5, 6, 7, 0, 0, 0)], receivers=[(12, 9, 10), (3, 3, 18)], input_data=[])
    r_or, (pos, m, _, _), _ = fc.run_oracle(case)
    r, nodes, info = fc.run_ours(capi, case)
    assert "plain" in info["kernel"] and info["dims"] == (24, 20, 21)
    assert np.array_equal(r, r_or)
    with pytest.raises(capi.PfdtdError):
        fc.run_ours(capi, case, kernel=capi.KERNEL_TMA)


def test_step_api_matches_run(capi, gpu):
    """launchFDTD3dStep (kernels3d.cu:376-482): stepping one at a time gives the same responses."""
    case = CASES["shoebox_48x40x49_fwd_f32_6mat_2parts"]
    r_run, _, _ = fc.run_ours(capi, case)
    s = capi.Solver()
    s.setup_mesh(case["bid"], case["mat"], case["block"], case["update_type"], capi.F32, oracle.params(fc.LAM, 0), case["materials"])
    s.make_partition(2, [0, 0])
    src = np.asarray(case["sources"], dtype=np.int32).reshape(-1, 6)
    s.set_sources(src[:, :3], src[:, 3], fc.source_table(case))
    s.set_receivers(case["receivers"])
    n = 60
    out = np.zeros((len(case["receivers"]), n), np.float32)
    for i in range(n):
        s.step(i, 1, out, n)
    s.close()
    assert np.array_equal(out, r_run[:, :n])


def test_run_in_blocks_and_interrupt(capi, gpu):
    case = CASES["c1_shoebox64_fwd_f32"]
    r_or, _, _ = fc.run_oracle(case)
    s = capi.Solver()
    s.setup_mesh(case["bid"], case["mat"], case["block"], 0, capi.F32, oracle.params(fc.LAM, 0), case["materials"])
    s.make_partition(1, [0])
    s.set_sources([[32, 32, 32]], [0], fc.source_table(case))
    s.set_receivers(case["receivers"])
    seen = []
    r, sps = s.run(500, progress=lambda step, mx, t: seen.append((step, mx)))
    assert np.array_equal(r, r_or) and sps > 0
    assert seen == [(0, 500), (100, 500), (200, 500), (300, 500), (400, 500)]    # PROGRESS_MOD 100 (kernels3d.h:36)
    s.close()
    # interrupt: polled before the first block
    s = capi.Solver()
    s.setup_mesh(case["bid"], case["mat"], case["block"], 0, capi.F32, oracle.params(fc.LAM, 0), case["materials"])
    s.make_partition(1, [0])
    s.set_receivers(case["receivers"])
    r, _ = s.run(50, interrupt=lambda: 1)
    assert not r.any()
    s.close()
    # an installed callback that never fires changes nothing (the run is cut into ~25 ms blocks, drained before each poll);
    # one that fires at its 4th poll stops the run early: the steps before it are there, the rest is not
    for fire_at in (None, 4):
        s = capi.Solver()
        s.setup_mesh(case["bid"], case["mat"], case["block"], 0, capi.F32, oracle.params(fc.LAM, 0), case["materials"])
        s.make_partition(1, [0])
        s.set_sources([[32, 32, 32]], [0], fc.source_table(case))
        s.set_receivers(case["receivers"])
        polls, seen = [], []
        try:
            r, _ = s.run(500, interrupt=lambda: (polls.append(1), int(fire_at is not None and len(polls) >= fire_at))[1],
                         progress=lambda step, mx, t: seen.append(step))
        finally:
            s.close()
        if fire_at is None:
            assert np.array_equal(r, r_or) and seen == [0, 100, 200, 300, 400] and len(polls) >= 6
        else:
            done = 8 + 92 + 100                                       # blocks before the 4th poll: 8, 92, 100
            assert len(polls) == 4 and np.array_equal(r[:, :done], r_or[:, :done]) and not r[:, done + 1:].any()


@pytest.mark.parametrize("name", ["shoebox_48x40x49_ctr_f64_6mat_5parts", "hall_96x128x64_fwd_f32_5mat_oct1"])
def test_halo_by_peer_stores_equals_halo_by_copies(capi, gpu, name):
    """Slabs in one process: the edge launches store their plane into the neighbour's halo plane themselves
    (PFDTD_OPT_PEER_STORES, default) or the planes are copied afterwards -- same responses, and equal to one slab."""
    case = CASES[name]
    base, _, _ = fc.run_ours(capi, case, n_parts=1)
    for n in (2, 3, 5):
        for peer in (1, 0):
            for graph in (1, 0):
                r, _, _ = fc.run_ours(capi, case, n_parts=n, opts=[(capi.OPT_PEER_STORES, peer), (capi.OPT_USE_GRAPH, graph)])
                assert np.array_equal(r, base), (n, peer, graph)


def test_step_api_time_reversal(capi, gpu):
    """launchFDTD3dStep's flip rule (kernels3d.cu:386,462-465): the pointers are flipped only when the direction did
    not change, so a change of direction runs the lossless scheme backwards -- the field retraces its steps."""
    dims = (64, 48, 40)
    from parallelfdtd_b200 import synth
    bid, mat = synth.shoebox(dims, 1)
    tab = np.zeros((1, 20), dtype=np.float32)                       # admittance 0: lossless walls
    s = capi.Solver()
    s.setup_mesh(bid, mat, (32, 4, 1), capi.SRL_FORWARD, capi.F32, oracle.params(fc.LAM, 0), tab)
    s.make_partition(2, [0, 0])
    n_total = 200
    src = np.zeros((1, n_total), dtype=np.float32)
    src[0, 1] = 1.0                                                  # impulse at step 1, nothing afterwards
    s.set_sources([[30, 20, 18]], [capi.SRC_HARD], src)
    s.set_receivers([[40, 25, 20]])
    out = np.zeros((1, n_total), np.float32)
    k, n = 30, 70
    snap = None
    for i in range(n):
        s.step(i, 1, out, n_total)
        if i == k - 1:
            snap = s.capture_mesh()                                  # P^k
    for m in range(n - k + 1):                                       # first reversed call does not flip, then one step back per call
        s.step(100 + m, -1, out, n_total)                            # step indices with a zero source sample
    back = s.capture_mesh()
    s.close()
    assert np.abs(snap).max() > 0
    assert fc.rel_l2(back, snap) < 1e-4


def test_shared_update_type_runs_the_forward_equations(capi, gpu):
    """UpdateType SHARED (1): the reference's sliced kernel is numerically broken as written (unmasked position byte,
    kernels3d.cu:559, SURVEY C-3); its equation is the forward one, so type 1 gives the responses of type 0."""
    case = {c["name"]: c for c in fc.parity_cases()}["shoebox_48x40x49_fwd_f32_6mat_2parts"]
    r0, _, _ = fc.run_ours(capi, case)
    r1, _, info = fc.run_ours(capi, dict(case, update_type=1))
    assert "forward" in info["kernel"] and np.abs(r0).max() > 0 and np.array_equal(r0, r1)
