"""CPU: the oracle's digital-impedance-filter boundary (not in the reference; parity unpinned).  Anchors:
order 0 == the pinned locally-reacting boundary bit for bit; a filter with only b0 set == order 0; slab count
does not matter; a passive filter keeps a closed room bounded."""
import numpy as np
import pytest

from oracle import oracle
from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc


@pytest.mark.parametrize("scheme_type", [0, 2])
@pytest.mark.parametrize("double", [False, True])
def test_order_zero_is_the_reference_boundary(scheme_type, double):
    dt = np.float64 if double else np.float32
    bid, mat = synth.shoebox((24, 20, 22), 3)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (8, 4, 1), scheme_type, double)
    prm = oracle.params(fc.LAM, 0, double)
    tab = synth.material_table([0.9, 0.8, 0.7]).astype(dt)
    steps = 200
    src = oracle.source_samples(1, steps, double=double)
    a, _ = oracle.run(pos, m, scheme_type, prm, tab, [(8, 8, 8)], [0], src, [(15, 12, 10)], steps, 1, 0)
    t0 = np.zeros((3, 20), dt)
    t0[:, 0] = tab[:, 0]
    for order in (0, 1, 4):
        b, _ = oracle.run_dif(pos, m, scheme_type, prm, t0, order, [(8, 8, 8)], [0], src, [(15, 12, 10)], steps, 1)
        assert np.array_equal(a, b), order


@pytest.mark.parametrize("name", [c["name"] for c in fc.dif_cases()][:3])
def test_dif_partition_invariance(name):
    case = {c["name"]: c for c in fc.dif_cases()}[name]
    case = dict(case, steps=120)
    base, _, _ = fc.run_oracle(case, n_parts=1)
    assert np.abs(base).max() > 0
    for n in (2, 5):
        r, _, _ = fc.run_oracle(case, n_parts=n)
        assert np.array_equal(r, base)


def test_passive_filter_room_stays_bounded():
    bid, mat = synth.shoebox((20, 18, 16), 2)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (4, 2, 1), 0, True)
    prm = oracle.params(fc.LAM, 0, True)
    t = np.zeros((2, 20))
    t[:, 0] = [0.05, 0.08]; t[:, 1] = [0.02, 0.01]; t[:, 2] = [-0.4, -0.3]     # Y(z) = (b0 + b1 z^-1)/(1 + a1 z^-1), positive real
    steps = 3000
    g = oracle.source_samples(1, steps, double=True)
    src = np.zeros(steps); src[1:] = np.diff(g); src[80:] = 0
    r, _ = oracle.run_dif(pos, m, 0, prm, t, 1, [(7, 8, 6)], [1], src, [(12, 9, 8)], steps, 1)
    assert np.isfinite(r).all()
    assert np.abs(r[:, -300:]).max() < np.abs(r[:, 100:400]).max()      # absorbing walls: the field decays
