"""GPU: the capture path (pfdtd_capture_slice / pfdtd_capture_mesh; reference captureSliceFast / captureMesh,
src/kernels/visualizationUtils.cu:111-254) against the full field read back slab by slab, for 1, 2 and 5 slabs,
both dtypes, and the Python module `libPyFDTD` (API of the reference's src/AppPy.cpp) end to end."""
import os
import sys

import numpy as np
import pytest

from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {c["name"]: c for c in fc.parity_cases()}


def _solver_after(capi, case, n_parts, steps):
    s = capi.Solver()
    s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 1)
    dt = capi.F64 if case["double"] else capi.F32
    s.setup_mesh(case["bid"], case["mat"], case["block"], case["update_type"], dt, fc.params_of(case, False), case["materials"])
    s.make_partition(n_parts, [0] * n_parts)
    src = np.asarray(case["sources"], dtype=np.int32).reshape(-1, 6)
    s.set_sources(src[:, :3], src[:, 3], fc.source_table(case))
    s.set_receivers(case["receivers"])
    s.run(steps)
    return s


def _assembled_field(s, n_parts):
    X, Y, Z = s.dims()
    field = np.zeros((Z, Y, X), dtype=s.np_dtype)
    pos = np.zeros((Z, Y, X), dtype=np.uint8)
    for k in range(n_parts):
        first, nz, _ = s.partition(k)
        lo, hi = (0 if k == 0 else 1), (nz if k == n_parts - 1 else nz - 1)
        field[first + lo:first + hi] = s.export_partition_pressure(k)[lo:hi]
        pos[first + lo:first + hi] = s.export_partition_nodes(k)[0][lo:hi]
    return field, pos


@pytest.mark.parametrize("name", ["shoebox_48x40x49_fwd_f32_6mat_2parts", "shoebox_48x40x49_ctr_f64_6mat_5parts"])
@pytest.mark.parametrize("n_parts", [1, 2, 5])
def test_slice_and_mesh_capture_match_the_field(capi, gpu, name, n_parts):
    case = CASES[name]
    s = _solver_after(capi, case, n_parts, 60)
    try:
        field, pos = _assembled_field(s, n_parts)
        assert np.abs(field).max() > 0
        assert np.array_equal(s.capture_mesh(), field)
        X, Y, Z = s.dims()
        for orientation, lim, take in ((0, Z, lambda a, i: a[i]), (1, Y, lambda a, i: a[:, i, :]), (2, X, lambda a, i: a[:, :, i])):
            for idx in (0, 1, lim // 2, lim - 1):
                p, b = s.capture_slice(idx, orientation, with_position=True)
                assert np.array_equal(p, take(field, idx)), (orientation, idx)
                assert np.array_equal(b, take(pos, idx)), (orientation, idx)
            with pytest.raises(capi.PfdtdError) as e:
                s.capture_slice(lim, orientation)
            assert e.value.code == 4          # PFDTD_ERR_RANGE
    finally:
        s.close()


def test_python_module_runs_like_the_reference_test_bench(capi, gpu):
    """reference python/testBench.py:110-149 with a 1 m box: initializeGeometryPy -> runSimulation / runCapture."""
    from parallelfdtd_b200 import build
    build.build_py_module()
    sys.path.insert(0, os.path.join(ROOT, "parallelfdtd_b200"))
    import libPyFDTD as pf

    L = 1.0
    v = np.array([[0, 0, 0], [L, 0, 0], [L, L, 0], [0, L, 0], [0, 0, L], [L, 0, L], [L, L, L], [0, L, L]], dtype=np.float32)
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]
    tri = np.array([t for a, b, c, d in quads for t in ((a, b, c), (a, c, d))], dtype=np.uint32)

    def make(double, captures):
        app = pf.App()
        app.initializeDevices()
        app.initializeGeometryPy(tri.flatten().tolist(), v.flatten().tolist())
        app.setUpdateType(0)
        app.setNumSteps(120)
        app.setSpatialFs(7000)
        app.setDouble(double)
        app.forcePartitionTo(1)
        app.addSurfaceMaterials([0.9] * (len(tri) * 20), len(tri), 20)
        app.addSource(0.5, 0.5, 0.5, 0, 0, 0)
        app.addReceiver(0.6, 0.6, 0.6)
        app.addReceiver(0.3, 0.4, 0.7)
        if captures:
            app.addSliceToCapture(10, 50, 1)
            app.addSliceToCapture(12, 60, 0)
        return app

    app = make(False, False)
    app.runSimulation()
    r = np.array([app.getResponse(i) for i in range(2)])
    assert r.shape == (2, 120) and np.isfinite(r).all() and np.abs(r).max() > 0
    assert app.getNumElems() > 0 and app.getMvox() > 0
    app.close()

    cap = make(False, True)
    cap.runCapture()                                   # step-by-step path: launchFDTD3dStep + captures
    rc = np.array([cap.getResponse(i) for i in range(2)])
    assert np.array_equal(rc, r)                       # same responses as the batched run
    assert cap.getNumberOfSliceCaptures() == 2
    X, Y, Z = cap.getDims()
    s1, s0 = cap.getSliceCapture(0), cap.getSliceCapture(1)
    assert s1.shape == (Z, X) and s0.shape == (Y, X) and np.abs(s1).max() > 0
    cap.close()

    dbl = make(True, False)
    dbl.runSimulation()
    rd = np.array([dbl.getResponseDouble(i) for i in range(2)])
    assert fc.rel_l2(r, rd) < 1e-4 and rd.dtype == np.float64
    dbl.close()
    with pytest.raises(RuntimeError):                     # no geometry: fails loudly, like every other entry point
        pf.App().runVisualization()
    vis = make(False, False)                              # headless viewer session: 2 x fs steps through executeStep
    vis.runVisualization()
    rv = np.array(vis.getResponse(0))
    assert rv.shape == (14000,) and np.array_equal(rv[:120], r[0])
    vis.close()
