"""GPU: receivers recorded and sources injected inside the update launch (single slab, PFDTD_OPT_FUSE_SRCREC, the
default) against the separate source/receiver launch per step and against the oracle: same responses, same fields.
Reference order per step: source(n) -> update -> receiver(n) (kernels3d.cu:93-104,164-173)."""
import numpy as np
import pytest

from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu
CASES = {c["name"]: c for c in fc.parity_cases() + fc.interp_cases() + fc.dif_cases()}


def _run(capi, case, fuse, graph, blocks, soft=0):
    s = capi.Solver()
    try:
        s.set_option(capi.OPT_FUSE_SRCREC, fuse)
        s.set_option(capi.OPT_USE_GRAPH, graph)
        s.set_option(capi.OPT_SOFT_ACCUMULATE, soft)
        s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0 if case["update_type"] >= 3 else 1)
        if case.get("dif_order") is not None:
            s.set_option(capi.OPT_DIF_ORDER, case["dif_order"])
        s.setup_mesh(case["bid"], case["mat"], case["block"], case["update_type"], capi.F64 if case["double"] else capi.F32,
                     fc.params_of(case, False), case["materials"])
        s.make_partition(1, [0])
        src = np.asarray(case["sources"], dtype=np.int32).reshape(-1, 6)
        s.set_sources(src[:, :3], src[:, 3], fc.source_table(case))
        s.set_receivers(case["receivers"])
        steps = case["steps"]
        s.reserve_steps(steps)
        first = 0
        for b in blocks:                                  # several enqueue blocks: the hand-over of the step counter
            n = min(b, steps - first)
            s.enqueue_steps(first, n)
            first += n
        s.enqueue_steps(first, steps - first)
        s.sync()
        launches = s.launch_count()
        return s.fetch_responses(steps), s.export_partition_pressure(0, 0), s.export_partition_pressure(0, 1), launches
    finally:
        s.close()


@pytest.mark.parametrize("name", ["c1_shoebox64_fwd_f32", "shoebox_48x40x49_ctr_f64_6mat_5parts", "hall_96x128x64_fwd_f32_5mat_oct1",
                                  "iwb_shoebox_48x40x49_f64_6mat", "dif2_shoebox_48x40x49_fwd_f32", "dif2_hall_96x128x64_iiso_f32"])
def test_fused_sources_and_receivers_equal_the_separate_launch(capi, gpu, name):
    case = CASES[name]
    want, _, _ = fc.run_oracle(case, n_parts=1, matidx=0 if case["update_type"] >= 3 else 1)
    base = _run(capi, case, 0, 0, [])
    assert np.array_equal(base[0], want)
    for graph, blocks in ((1, []), (0, [7, 1, 30]), (1, [40, 9])):
        got = _run(capi, case, 1, graph, blocks)
        assert np.array_equal(got[0], base[0]), (name, graph, blocks)
        assert np.array_equal(got[1], base[1]) and np.array_equal(got[2], base[2]), (name, graph, blocks)
    fused_launches = _run(capi, case, 1, 0, [])[3]
    if case.get("dif_order"):                             # the filter kernels keep the separate launch
        assert fused_launches == base[3]
    else:
        assert fused_launches < base[3]                   # one launch per step instead of two


def test_fused_soft_sources_accumulate_once(capi, gpu):
    case = dict(CASES["shoebox_48x40x49_fwd_f32_6mat_2parts"])
    a = _run(capi, case, 0, 0, [], soft=1)
    b = _run(capi, case, 1, 1, [50, 3], soft=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0], _run(capi, case, 1, 1, [], soft=0)[0])      # the option is observable on this case
