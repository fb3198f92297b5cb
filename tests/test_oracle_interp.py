"""CPU: the oracle's interpolated 27-point schemes (IISO / IWB).  The reference does not contain them
 -- parity unpinned -- so the anchors are: (1) with the SRL weights the equation IS the
pinned SRL_FORWARD update, (2) the standard weights satisfy the consistency condition
6 d1 + 12 d2 + 8 d3 + d4 = 2, (3) rigid rooms neither gain nor lose energy at the stability limit,
(4) slab count does not change results, (5) wave speed: first arrival along an axis after distance/lambda steps."""
import numpy as np
import pytest

from oracle import oracle
from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc


@pytest.mark.parametrize("ut", [3, 4])
def test_weights_are_consistent(ut):
    lam = oracle.interp_lambda(ut)
    d = oracle.interp_coefficients(ut, lam * lam)
    assert abs(6 * d[0] + 12 * d[1] + 8 * d[2] + d[3] - 2) < 1e-15
    assert np.allclose(d, [0.25, 0.125, 0.0, -1.0] if ut == 3 else [0.25, 0.125, 0.0625, -1.5], rtol=0, atol=1e-15)
    # in float the squared Courant number is exact (0.75 / 1.0), so the weights are exact binary fractions
    d32 = oracle.interp_coefficients(ut, np.float32(lam * lam))
    assert d32 == ([0.25, 0.125, 0.0, -1.0] if ut == 3 else [0.25, 0.125, 0.0625, -1.5])


@pytest.mark.parametrize("double", [False, True])
def test_srl_weights_reproduce_the_pinned_forward_scheme(double):
    bid, mat = synth.shoebox((24, 20, 22), 3)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (8, 4, 1), 0, double)
    prm = oracle.params(fc.LAM, 0, double)
    tab = synth.material_table([0.9, 0.8, 0.7])
    steps = 200
    src = oracle.source_samples(1, steps, double=double)
    a, _ = oracle.run(pos, m, 0, prm, tab, [(8, 8, 8)], [0], src, [(15, 12, 10), (3, 3, 3)], steps, 1, 0)
    # 2 - 6*lam2 with ONE rounding, as the reference's contracted fma computes it (exact in double for a float lam2)
    d = [float(prm[1]), 0.0, 0.0, float(2 - 6 * np.float64(prm[1]))]
    b, _ = oracle.run(pos, m, 3, oracle.params_interp(fc.LAM, 0, d, double), tab, [(8, 8, 8)], [0], src, [(15, 12, 10), (3, 3, 3)], steps, 1, 0)
    assert np.abs(a).max() > 0
    # same equation, different summation order of the six axial taps: rounding-level agreement
    assert fc.rel_l2(b, a) < (1e-12 if double else 2e-5)


@pytest.mark.parametrize("ut", [3, 4])
def test_rigid_room_is_stable_and_lossless(ut):
    bid, mat = synth.shoebox((20, 18, 16), 1)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (4, 2, 1), 0, True)
    lam = oracle.interp_lambda(ut)
    p8 = oracle.params_interp(lam, 0, oracle.interp_coefficients(ut, lam * lam), True)
    tab = np.zeros((1, 20))                                   # Y = 0: rigid walls
    steps = 3000
    g = oracle.source_samples(1, steps, double=True)          # gaussian pulse
    src = np.zeros(steps)
    src[1:] = np.diff(g)                                      # zero-mean (a net injected volume would excite the linear-in-time DC mode)
    src[80:] = 0                                              # accumulate for 80 steps, then leave the node free
    r, _ = oracle.run(pos, m, 3, p8, tab, [(7, 8, 6)], [1], src, [(12, 9, 8), (4, 4, 4)], steps, 1, 0, 1)
    early = np.abs(r[:, 100:600]).max()
    late = np.abs(r[:, -500:]).max()
    assert np.isfinite(r).all() and early > 0
    assert 0.2 * early < late < 5 * early                     # no blow-up, no decay


@pytest.mark.parametrize("ut", [3, 4])
def test_partition_invariance_and_first_arrival(ut):
    bid, mat = synth.shoebox((40, 24, 49), 6)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (8, 4, 1), 0, False)
    lam = oracle.interp_lambda(ut)
    p8 = oracle.params_interp(lam, 0, oracle.interp_coefficients(ut, np.float32(lam * lam)), False)
    tab = synth.material_table(list(np.linspace(0.99, 0.5, 6)))
    steps = 120
    src = oracle.source_samples(0, steps)
    rx = [(20, 12, 5), (20, 12, 23), (20, 12, 24), (30, 12, 24)]
    base, _ = oracle.run(pos, m, 3, p8, tab, [(20, 12, 24)], [0], src, rx, steps, 1, 0)
    for n in (2, 5, 7):
        r, _ = oracle.run(pos, m, 3, p8, tab, [(20, 12, 24)], [0], src, rx, steps, n, 0)
        assert np.array_equal(r, base), n
    # a 27-point stencil reaches a node 10 voxels away along x after 10 updates (impulse at step 1 -> response index 10)
    first = np.flatnonzero(base[3] != 0)[0]
    assert first == 10


@pytest.mark.parametrize("ut,a,b", [(3, 1.0 / 6.0, 0.0), (4, 1.0 / 4.0, 1.0 / 16.0)])
def test_weights_follow_the_published_family_and_land_on_the_right_neighbours(ut, a, b):
    """Literature anchor (Kowalczyk & van Walstijn, "Room acoustics simulation using 3-D compact explicit FDTD schemes",
    IEEE TASLP 2011): d1 = lam^2 (1 - 4a + 4b), d2 = lam^2 (a - 2b), d3 = lam^2 b, d4 = 2 (1 - 3 lam^2 + 6 lam^2 a - 4 lam^2 b),
    IISO (a, b, lam) = (1/6, 0, sqrt(3)/2), IWB = (1/4, 1/16, 1).  And a known answer for WHERE the weights act: one step
    after a unit impulse in open air the six axial neighbours hold d1, the twelve edge neighbours d2, the eight corner
    neighbours d3, the voxel itself d4 and everything else 0."""
    lam = {3: np.sqrt(3.0) / 2.0, 4: 1.0}[ut]
    assert oracle.interp_lambda(ut) == pytest.approx(lam, abs=1e-15)
    l2 = lam * lam
    lit = [l2 * (1 - 4 * a + 4 * b), l2 * (a - 2 * b), l2 * b, 2 * (1 - 3 * l2 + 6 * l2 * a - 4 * l2 * b)]
    d = oracle.interp_coefficients(ut, l2)
    assert np.allclose(d, lit, rtol=0, atol=1e-15)

    bid, mat = synth.shoebox((16, 16, 16), 1)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (4, 2, 1), 0, True)
    p8 = oracle.params_interp(lam, 0, d, True)
    c = 8
    offs = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)] + [(2, 0, 0), (0, -2, 0), (2, 1, 0), (1, 1, 2)]
    rec = [(c + dx, c + dy, c + dz) for dx, dy, dz in offs]
    src = np.zeros((1, 2))
    src[0, 0] = 1.0
    r, _ = oracle.run(pos, m, 3, p8, np.zeros((1, 20)), [(c, c, c)], [0], src, rec, 2, 1, 0)
    for (dx, dy, dz), v in zip(offs, r[:, 0]):                # response[step 0] = the field after the first update
        n_off = abs(dx) + abs(dy) + abs(dz)
        if max(abs(dx), abs(dy), abs(dz)) > 1:
            want = 0.0
        else:
            want = {0: d[3], 1: d[0], 2: d[1], 3: d[2]}[n_off]
        assert v == want, ((dx, dy, dz), v, want)


@pytest.mark.parametrize("ut,a,b", [(3, 1.0 / 6.0, 0.0), (4, 1.0 / 4.0, 1.0 / 16.0)])
@pytest.mark.parametrize("k", [(0.31, 0.0, 0.0), (0.4, 0.4, 0.0), (0.23, -0.37, 0.52), (1.1, 0.9, 1.3)])
def test_plane_waves_follow_the_published_dispersion_relation(ut, a, b, k):
    """Second literature anchor, independent of how the weights are written: the numerical dispersion relation of the compact
    explicit family (Kowalczyk & van Walstijn 2011),
        sin^2(w T / 2) = lam^2 [ (sx + sy + sz) - 4 a (sx sy + sx sz + sy sz) + 16 b sx sy sz ],   s_i = sin^2(k_i h / 2).
    A sampled plane wave cos(w n T - k.x) with that w must be carried one step forward exactly by the oracle's update in
    open air (k in radians per voxel; the last case is close to the grid's Nyquist limit)."""
    lam = oracle.interp_lambda(ut)
    l2 = lam * lam
    sx, sy, sz = (np.sin(0.5 * q) ** 2 for q in k)
    s2 = l2 * ((sx + sy + sz) - 4 * a * (sx * sy + sx * sz + sy * sz) + 16 * b * sx * sy * sz)
    assert 0.0 <= s2 <= 1.0                                    # inside the stability limit
    wT = 2.0 * np.arcsin(np.sqrt(s2))

    X, Y, nz = 24, 20, 12
    bid, mat = synth.shoebox((X, Y, nz + 8), 1)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (4, 2, 1), 0, True)
    pos, m = pos[4:4 + nz], m[4:4 + nz]                        # an all-air run of planes (walls only in x and y)
    assert pos.shape == (nz, Y, X)
    zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(Y), np.arange(X), indexing="ij")
    phase = k[0] * xx + k[1] * yy + k[2] * zz
    p8 = oracle.params_interp(lam, 0, oracle.interp_coefficients(ut, l2), True)
    new = oracle.step_slab(pos, m, 3, p8, np.zeros((1, 20)), np.cos(phase), np.cos(wT + phase))
    want = np.cos(wT - phase)
    inner = (slice(1, nz - 1), slice(3, Y - 3), slice(3, X - 3))   # away from the walls; planes 0 / nz-1 are halos
    assert np.abs(new[inner] - want[inner]).max() < 1e-13
    # and the relation discriminates: the leapfrog frequency of the 7-point scheme (a = b = 0) is measurably off
    wT7 = 2.0 * np.arcsin(np.sqrt(min(1.0, l2 * (sx + sy + sz))))
    if abs(wT7 - wT) > 1e-6:
        assert np.abs(new[inner] - np.cos(wT7 - phase)[inner]).max() > 1e-8
