"""GPU: the TMA z-march kernels against the plain per-voxel kernel on grids large enough that several CTAs share
an SM and the shared-memory stage ring wraps many times -- the regime the small parity cases (<= 148 CTAs, one per
SM) never reach.  Whole pressure fields are compared bit for bit after a few dozen steps.

Regression for a shared-memory write-after-read hazard found on B200 (profiles/README.md, "stage release"): a stage
must not be handed back to the TMA producer before a store that depends on the values loaded from it; an early
release let the refill overtake loads that had been issued but not performed (fp64, two CTAs per SM).
"""
import numpy as np
import pytest

from oracle import oracle
from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu


def _fields(capi, dims, geom, update_type, double, steps, opts, dif_order=0, n_mat=5):
    bid, mat = (synth.hall if geom == "hall" else synth.shoebox)(dims, n_mat)
    npdt = np.float64 if double else np.float32
    if dif_order:
        tab = fc.dif_table(n_mat, dif_order).astype(npdt)
    else:
        tab = synth.material_table(list(np.linspace(0.99, 0.5, n_mat))).astype(npdt)
    lam = oracle.interp_lambda(update_type) if update_type >= 3 else fc.LAM
    s = capi.Solver()
    try:
        s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
        for k, v in opts:
            s.set_option(k, v)
        if dif_order:
            s.set_option(capi.OPT_DIF_ORDER, dif_order)
        s.setup_mesh(bid, mat, (32, 4, 1), update_type, capi.F64 if double else capi.F32, oracle.params(lam, 0, double), tab)
        s.make_partition(1, [0])
        n = np.arange(steps + 8, dtype=np.float64)
        src = np.exp(-0.5 * ((n - 12.0) / 3.0) ** 2)[None, :]
        s.set_sources([[dims[0] // 2 - 3, dims[1] // 3, dims[2] // 3]], [capi.SRC_HARD], src)
        s.set_receivers([[dims[0] // 2, dims[1] // 2, dims[2] // 2]])
        s.reserve_steps(steps)
        s.enqueue_steps(0, steps)
        s.sync()
        f = (s.export_partition_pressure(0, 0), s.export_partition_pressure(0, 1))
        name = s.kernel_name()
    finally:
        s.close()
    return f, name


# (dims, geometry, update_type, double, steps, chunk option, repeats)
GRIDS = [
    ((96, 128, 64), "hall", 0, True, 80, 4, 4),      # 256 CTAs of the fp64 128x8 tile: two per SM, ring wraps after 4 planes
    ((96, 128, 64), "hall", 2, True, 80, 4, 2),
    ((96, 128, 64), "hall", 0, False, 80, 4, 2),
    ((256, 256, 128), "hall", 0, True, 40, 4, 2),    # 2048 CTAs, several waves
    ((256, 256, 128), "hall", 0, True, 40, 0, 1),    # auto chunk
    ((256, 256, 128), "hall", 0, False, 40, 0, 1),
    ((256, 256, 128), "hall", 2, False, 40, 5, 1),
    ((256, 256, 128), "hall", 3, False, 30, 4, 2),   # 27-point kernels
    ((256, 256, 128), "hall", 4, True, 30, 0, 1),
]


@pytest.mark.parametrize("dims,geom,ut,double,steps,chunk,reps", GRIDS)
def test_tma_equals_plain_on_multi_cta_per_sm_grids(capi, gpu, dims, geom, ut, double, steps, chunk, reps):
    (b0, b1), _ = _fields(capi, dims, geom, ut, double, steps, [(capi.OPT_KERNEL, capi.KERNEL_PLAIN)])
    assert np.abs(b0).max() > 0
    for _ in range(reps):
        (f0, f1), name = _fields(capi, dims, geom, ut, double, steps,
                                 [(capi.OPT_KERNEL, capi.KERNEL_TMA), (capi.OPT_USE_GRAPH, 0), (capi.OPT_TMA_CHUNK, chunk)])
        assert "tma" in name
        assert np.array_equal(f0, b0) and np.array_equal(f1, b1), name


@pytest.mark.parametrize("ut,double,order", [(0, True, 2), (0, False, 2), (3, False, 1)])
def test_filter_boundaries_do_not_depend_on_the_chunking_on_large_grids(capi, gpu, ut, double, order):
    """DIF boundaries exist only in the TMA kernels: compare chunkings (and graph replay) with each other."""
    dims, steps = (256, 256, 128), 40
    (b0, b1), _ = _fields(capi, dims, "hall", ut, double, steps, [(capi.OPT_TMA_CHUNK, 126)], dif_order=order)
    assert np.abs(b0).max() > 0
    for opts in ([(capi.OPT_TMA_CHUNK, 4), (capi.OPT_USE_GRAPH, 0)], [(capi.OPT_TMA_CHUNK, 0)], [(capi.OPT_TMA_CHUNK, 7), (capi.OPT_USE_GRAPH, 0)]):
        (f0, f1), name = _fields(capi, dims, "hall", ut, double, steps, opts, dif_order=order)
        assert np.array_equal(f0, b0) and np.array_equal(f1, b1), name


def test_config2_room_512_cubed_tma_equals_plain(capi, gpu):
    """BASELINE config 2 at full size (512^3, 6 materials), fp32 and fp64 (half height in fp64 to bound memory/time)."""
    for dims, double in (((512, 512, 512), False), ((512, 512, 256), True)):
        (b0, b1), _ = _fields(capi, dims, "shoebox", 0, double, 24, [(capi.OPT_KERNEL, capi.KERNEL_PLAIN)], n_mat=6)
        (f0, f1), name = _fields(capi, dims, "shoebox", 0, double, 24, [(capi.OPT_KERNEL, capi.KERNEL_TMA)], n_mat=6)
        assert np.abs(b0).max() > 0
        assert np.array_equal(f0, b0) and np.array_equal(f1, b1), name
