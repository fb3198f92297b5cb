"""GPU: frequency-dependent boundaries (per-material digital impedance filters, states of the sparse boundary
nodes updated inside the same kernel pass) against the CPU oracle.  Not in the reference (parity unpinned by it);
anchor: filter order 0 with b0 = Y reproduces the pinned frequency-independent path bit for bit."""
import numpy as np
import pytest

from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu
CASES = {c["name"]: c for c in fc.dif_cases()}
TOL = {False: 1e-5, True: 1e-12}


@pytest.mark.parametrize("name", list(CASES))
def test_dif_parity_with_oracle(capi, gpu, name):
    case = CASES[name]
    r_or, _, _ = fc.run_oracle(case)
    r, _, info = fc.run_ours(capi, case, matidx=0)
    assert "tma" in info["kernel"]
    assert np.abs(r_or).max() > 0
    assert fc.rel_l2(r, r_or) <= TOL[case["double"]]
    assert np.array_equal(r, r_or)


@pytest.mark.parametrize("name", list(CASES))
def test_dif_partition_invariance(capi, gpu, name):
    case = CASES[name]
    base, _, _ = fc.run_ours(capi, case, n_parts=1, matidx=0)
    for n in (2, 5):
        r, _, _ = fc.run_ours(capi, case, n_parts=n, matidx=0)
        assert np.array_equal(r, base), n
    r, _, _ = fc.run_ours(capi, case, n_parts=1, matidx=0, opts=[(capi.OPT_TMA_CHUNK, 5), (capi.OPT_USE_GRAPH, 0)])
    assert np.array_equal(r, base)


@pytest.mark.parametrize("ut,double", [(0, False), (2, True), (3, False)])
def test_order_zero_filter_is_the_frequency_independent_boundary(capi, gpu, ut, double):
    srl = dict({c["name"]: c for c in fc.parity_cases()}["shoebox_48x40x49_fwd_f32_6mat_2parts"], update_type=ut, double=double)
    plain, _, _ = fc.run_ours(capi, srl, matidx=0)                       # scalar admittance materials[m*20 + 0]
    for order in (0, 2):
        c = dict(srl, dif_order=order)
        t = np.zeros_like(srl["materials"])
        t[:, 0] = srl["materials"][:, 0]                                  # b0 = Y, all other coefficients 0
        c["materials"] = t
        r, _, info = fc.run_ours(capi, c, matidx=0)
        assert np.array_equal(r, plain), (order, info["kernel"])


def test_filters_change_the_response_and_reset_clears_state(capi, gpu):
    case = CASES["dif2_shoebox_48x40x49_fwd_f32"]
    r, _, _ = fc.run_ours(capi, case, matidx=0)
    flat = dict(case, dif_order=0)
    r0, _, _ = fc.run_ours(capi, flat, matidx=0)
    assert not np.array_equal(r, r0)


def test_python_module_filter_materials_match_the_c_abi(capi, gpu):
    """Filter boundaries through the front end (libPyFDTD.App.addSurfaceFilters -> MaterialHandler -> App::initializeMesh
    -> PFDTD_OPT_DIF_ORDER) give the responses of the same run made directly on the C ABI."""
    import os
    import sys
    from oracle import oracle
    from parallelfdtd_b200 import build, synth
    build.build_py_module()
    sys.path.insert(0, os.path.dirname(build.py_module_path()))
    import libPyFDTD as pf

    steps, order = 150, 2
    bid, mat = synth.shoebox((48, 40, 49), 6)
    tab = synth.filter_material_table([0.99, 0.95, 0.9, 0.8, 0.7, 0.5], order).astype(np.float32)
    app = pf.App()
    app.initializeDevices()
    app.setVoxelVolumes(bid, mat)
    app.setUpdateType(0)
    app.setNumSteps(steps)
    app.setSpatialFs(7000)
    app.forcePartitionTo(1)
    app.addSurfaceFilters(tab[:, :2 * order + 1].flatten().tolist(), 6, order)
    dx = app.getDx()
    sxyz, rxyz = (20, 18, 22), (30, 25, 12)
    app.addSource(*[(c - 1) * dx for c in sxyz], 0, 0, 0)           # element = round(pos / dx) + 1
    app.addReceiver(*[(c - 1) * dx for c in rxyz])
    app.runSimulation()
    r_app = np.array(app.getResponse(0), dtype=np.float32)
    app.close()

    s = capi.Solver()
    s.set_option(capi.OPT_DIF_ORDER, order)
    s.setup_mesh(bid, mat, (32, 4, 1), capi.SRL_FORWARD, capi.F32, oracle.params(fc.LAM, 0), tab)
    s.make_partition(1, [0])
    s.set_sources([list(sxyz)], [capi.SRC_HARD], oracle.source_samples(0, steps)[None, :])
    s.set_receivers([list(rxyz)])
    r, _ = s.run(steps)
    s.close()
    s0 = capi.Solver()                                                # the same room without filters differs
    s0.setup_mesh(bid, mat, (32, 4, 1), capi.SRL_FORWARD, capi.F32, oracle.params(fc.LAM, 0), tab)
    s0.make_partition(1, [0])
    s0.set_sources([list(sxyz)], [capi.SRC_HARD], oracle.source_samples(0, steps)[None, :])
    s0.set_receivers([list(rxyz)])
    r0, _ = s0.run(steps)
    s0.close()
    assert np.abs(r).max() > 0 and not np.array_equal(r0, r)
    assert np.array_equal(r_app, r[0])
