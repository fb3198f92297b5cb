"""GPU: the edges of the C ABI's argument space -- empty source / receiver lists, zero steps, slabs as thin as the
partition rule allows, a step index beyond the response buffer, sources outside the mesh -- behave like the
reference where it defines the behaviour and fail with a status code (never a crash) where it does not."""
import numpy as np
import pytest

from oracle import oracle
from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu


def _solver(capi, dims=(48, 40, 33), n_parts=1, n_mat=3, double=False, ut=0, opts=()):
    bid, mat = synth.shoebox(dims, n_mat)
    npdt = np.float64 if double else np.float32
    tab = synth.material_table(list(np.linspace(0.9, 0.6, n_mat))).astype(npdt)
    s = capi.Solver()
    s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
    for k, v in opts:
        s.set_option(k, v)
    s.setup_mesh(bid, mat, (32, 4, 1), ut, capi.F64 if double else capi.F32, oracle.params(fc.LAM, 0, double), tab)
    s.make_partition(n_parts, [0] * n_parts)
    return s


def test_no_sources_and_no_receivers(capi, gpu):
    s = _solver(capi)
    s.set_sources(np.zeros((0, 3)), [], np.zeros((0, 0)))
    s.set_receivers(np.zeros((0, 3)))
    r, sps = s.run(20)
    assert r.shape == (0, 20) and sps >= 0
    assert not s.capture_mesh().any()                    # nothing was injected: the field stays zero
    s.close()


def test_receivers_without_sources_record_zeros_and_zero_steps_is_a_no_op(capi, gpu):
    s = _solver(capi, n_parts=2)
    s.set_receivers([[10, 10, 10], [20, 20, 20]])
    r, _ = s.run(15)
    assert r.shape == (2, 15) and not r.any()
    r0, _ = s.run(0)
    assert r0.shape == (2, 0)
    s.close()


@pytest.mark.parametrize("double", [False, True])
def test_thinnest_slabs_match_one_slab(capi, gpu, double):
    """Z = 16 in 8 slabs: floor(Z / N) = 2 owned planes per slab, 3-4 stored (cudaMesh.h:280-307); the first slab updates
    a single plane."""
    steps = 60
    src = oracle.source_samples(1, steps, double=double)[None, :]
    out = []
    for n in (1, 8):
        s = _solver(capi, dims=(48, 40, 16), n_parts=n, double=double)
        s.set_sources([[20, 18, 7]], [capi.SRC_HARD], src)
        s.set_receivers([[30, 25, 2], [12, 9, 13], [20, 18, 8]])
        r, _ = s.run(steps)
        out.append(r)
        if n == 8:
            assert [s.partition(k)[1] for k in range(8)] == [3, 4, 4, 4, 4, 4, 4, 3]
        s.close()
    assert np.abs(out[0]).max() > 0 and np.array_equal(out[0], out[1])


def test_one_owned_plane_per_slab(capi, gpu):
    """Z = 8 in 8 slabs: one owned plane each; the first slab stores two planes and updates none."""
    steps = 40
    src = oracle.source_samples(1, steps)[None, :]
    out = []
    for n, opts in ((1, []), (8, []), (8, [(capi.OPT_PEER_STORES, 0), (capi.OPT_USE_GRAPH, 0)])):
        s = _solver(capi, dims=(48, 40, 8), n_parts=n, opts=opts)
        s.set_sources([[20, 18, 4]], [capi.SRC_SOFT], src)
        s.set_receivers([[30, 25, 2], [12, 9, 6], [20, 18, 5]])
        r, _ = s.run(steps)
        out.append(r)
        if n == 8:
            assert [s.partition(k)[1] for k in range(8)] == [2, 3, 3, 3, 3, 3, 3, 2]
        s.close()
    assert np.abs(out[0]).max() > 0 and np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])


def test_more_slabs_than_the_rule_allows_is_an_error(capi, gpu):
    bid, mat = synth.shoebox((48, 40, 6), 1)
    s = capi.Solver()
    s.setup_mesh(bid, mat, (32, 4, 1), 0, capi.F32, oracle.params(fc.LAM, 0), synth.material_table([0.9]))
    with pytest.raises(capi.PfdtdError):
        s.make_partition(12, [0] * 12)                   # floor(6 / 12) = 0 planes per slab
    s.close()


def test_step_beyond_the_response_buffer_and_bad_direction(capi, gpu):
    s = _solver(capi)
    s.set_sources([[20, 18, 12]], [capi.SRC_HARD], oracle.source_samples(0, 10)[None, :])
    s.set_receivers([[22, 18, 12]])
    out = np.zeros((1, 10), np.float32)
    s.step(0, 1, out, 10)
    with pytest.raises(capi.PfdtdError) as e:
        s.step(10, 1, out, 10)
    assert e.value.code == 4                             # PFDTD_ERR_RANGE
    with pytest.raises(capi.PfdtdError):
        s.step(1, 0, out, 10)
    s.step(1, 1, None, 0)                                # no response buffer: any step index is fine
    s.close()


def test_source_or_receiver_outside_the_mesh(capi, gpu):
    s = _solver(capi)
    with pytest.raises(capi.PfdtdError) as e:
        s.set_sources([[20, 18, 400]], [capi.SRC_HARD], np.ones((1, 4), np.float32))
        s.run(4)
    assert e.value.code == 4
    s.close()
    s = _solver(capi)
    with pytest.raises(capi.PfdtdError):
        s.set_receivers([[4000, 18, 4]])
        s.run(4)
    s.close()
