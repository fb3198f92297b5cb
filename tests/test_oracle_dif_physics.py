"""CPU: a physical known answer for the frequency-dependent boundaries, which the reference does not contain and which
are therefore "parity unpinned" (DESIGN.md section 2): the NORMAL-INCIDENCE REFLECTANCE of a wall carrying the
digital impedance filter Y(z), measured with a plane pulse in a rigid-walled duct, against the closed form of the
discretised boundary condition

    R(w) = - (lam^2 (e^{-jk} - 1) + j lam Y(e^{jw}) sin w) / (lam^2 (e^{jk} - 1) + j lam Y(e^{jw}) sin w),
    sin^2(w/2) = lam^2 sin^2(k/2)            (axial dispersion relation of the 7-point scheme)

which follows from inserting p_i^n = e^{jwn} (e^{jki} + R e^{-jki}) into the boundary-node update
(1 + b) p0^{n+1} = (2 - lam^2) p0^n + lam^2 p1^n - (1 - b) p0^{n-1}, b = lam Y / 2, and tends to the textbook
(1 - Y) / (1 + Y) for w -> 0.  What this pins: the sign and the delay conventions of the filter recursion (transposed
direct form II on u = p^{n+1} - p^{n-1}), the coupling coefficient, and that order 0 is the scalar admittance -- for
filter orders 0, 1, 2 and 4.  The centred-difference boundary of the SRL scheme (the interior neighbour counted twice,
b = lam Y) gives, the same way,

    R(w) = (lam sin k - Y(e^{jw}) sin w) / (lam sin k + Y(e^{jw}) sin w)

and is checked for orders 0, 2 and 3.  The CUDA kernels are bit-equal to this oracle (tests/test_gpu_dif.py).
"""
import numpy as np
import pytest

from oracle import oracle
from parallelfdtd_b200 import synth

LAM = float(np.sqrt(1.0 / 3.0))
NX, NY, Z = 6, 6, 420            # air cross-section, duct length (voxels)
ZS, ZR = 10, 200                 # source plane, receiver plane
STEPS = 1500                     # incident pulse passes ZR around step 330, the reflection around 1080


def _duct(order, b, a, scheme=0):
    """-> pos, mat node volumes and the [4][20] material table: 0 rigid side walls, 1 / 2 / 3 the end wall's face, edge
    and corner voxels, whose admittance is Y / (number of missing neighbours) so that the cross-section stays uniform
    (the boundary term is proportional to 6 - K, and only one of an edge voxel's missing neighbours is the end wall)."""
    dims = (32, 8, Z)

    def inside(x, y, z):
        return (x >= 1) & (x <= NX) & (y >= 1) & (y <= NY) & (z >= 1) & (z <= Z - 2)
    bid = synth.bid_from_inside(inside, dims)
    mat = np.zeros_like(bid)
    x = np.arange(dims[0])[None, :]
    y = np.arange(dims[1])[:, None]
    missing = 1 + (x == 1) + (x == NX) + (y == 1) + (y == NY)
    end = bid[Z - 2] > 0
    mat[Z - 2][end] = np.broadcast_to(missing, end.shape)[end].astype(np.uint8)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (32, 4, 1), scheme, True)
    tab = np.zeros((4, 20), dtype=np.float64)
    for k in (1, 2, 3):
        tab[k, :order + 1] = np.asarray(b) / k
        tab[k, order + 1:2 * order + 1] = a
    return pos, m, tab


def _measured_reflectance(order, b, a, scheme=0):
    pos, m, tab = _duct(order, b, a, scheme)
    src_xyz = [[x, y, ZS] for y in range(1, NY + 1) for x in range(1, NX + 1)]
    n = np.arange(STEPS, dtype=np.float64)
    pulse = np.exp(-0.5 * ((n - 40.0) / 6.0) ** 2)        # band-limited well below the axial cut-off (w = 1.23 rad / sample)
    prm = oracle.params(LAM, 0, True)
    r, _ = oracle.run_dif(pos, m, scheme, prm, tab, order, src_xyz, [0] * len(src_xyz), np.tile(pulse, (len(src_xyz), 1)),
                          [[3, 3, ZR], [1, 1, ZR], [NX, 4, ZR]], STEPS, 1)
    assert np.max(np.abs(r[0] - r[1])) < 1e-12 and np.max(np.abs(r[0] - r[2])) < 1e-12      # the wave stays plane
    split = 700
    inc, ref = r[0].copy(), r[0].copy()
    inc[split:] = 0.0
    ref[:split] = 0.0
    assert np.abs(r[0][split - 20:split + 20]).max() < 1e-7 * np.abs(r[0]).max()                                     # the two pulses are separated
    nfft = 4096
    Fi, Fr = np.fft.rfft(inc, nfft), np.fft.rfft(ref, nfft)
    w = 2 * np.pi * np.arange(nfft // 2 + 1) / nfft
    return w, np.abs(Fr) / np.maximum(np.abs(Fi), 1e-300), np.abs(Fi) / np.abs(Fi).max()


def _closed_form(w, b, a, scheme=0):
    zi = np.exp(-1j * w)
    order = len(b) - 1
    Y = sum(b[i] * zi ** i for i in range(order + 1)) / (1.0 + sum(a[i] * zi ** (i + 1) for i in range(order)))
    k = 2.0 * np.arcsin(np.clip(np.sin(w / 2.0) / LAM, -1.0, 1.0))
    D = 1j * LAM * Y * np.sin(w)
    lam2 = LAM * LAM
    with np.errstate(invalid="ignore", divide="ignore"):              # w = 0 is 0 / 0 and is not used
        if scheme == 2:                                                # centred-difference boundary, see the module docstring
            return (LAM * np.sin(k) - Y * np.sin(w)) / (LAM * np.sin(k) + Y * np.sin(w)), Y
        return -(lam2 * (np.exp(-1j * k) - 1.0) + D) / (lam2 * (np.exp(1j * k) - 1.0) + D), Y


def _filter(order, y0=0.25):
    """Y0 times a cascade of real one-pole / one-zero sections (the family synth.filter_material_table uses)."""
    b, a = np.array([1.0]), np.array([1.0])
    for i in range(order):
        p = 0.55 - 0.12 * i
        q = p - 0.18 / (i + 1)
        b, a = np.convolve(b, [1.0, -q]), np.convolve(a, [1.0, -p])
    return y0 * b, a[1:]


@pytest.mark.parametrize("scheme,order", [(0, 0), (0, 1), (0, 2), (0, 4), (2, 0), (2, 2), (2, 3)])
def test_normal_incidence_reflectance_matches_the_closed_form(scheme, order):
    b, a = _filter(order)
    w, R_meas, weight = _measured_reflectance(order, b, a, scheme)
    R_exact, Y = _closed_form(w, b, a, scheme)
    band = (weight > 1e-4) & (w > 0.02)                                  # where the pulse has energy (w < 0.72 rad / sample)
    assert band.sum() > 400
    err = np.max(np.abs(R_meas[band] - np.abs(R_exact[band])))
    assert err < 1e-3, err                                             # measured: 3e-4 (window leakage); the textbook formula is off by up to 0.09 here
    # the filters really are frequency dependent here, and the low-frequency limit is the textbook reflectance
    if order:
        assert np.ptp(np.abs(R_exact[band])) > 0.02
    lo = (w > 0.02) & (w < 0.06)
    textbook = np.abs((1.0 - Y[lo]) / (1.0 + Y[lo]))
    assert np.max(np.abs(R_meas[lo] - textbook)) < 5e-3


def test_order_zero_is_the_scalar_admittance_of_the_pinned_path():
    """the same duct through the frequency-independent entry point (the reference's own boundary, pinned bit for bit by
    the golden vectors) gives the same response as an order-0 filter with b0 = Y"""
    b, a = _filter(0)
    pos, m, tab = _duct(0, b, a)
    src_xyz = [[x, y, ZS] for y in range(1, NY + 1) for x in range(1, NX + 1)]
    n = np.arange(600, dtype=np.float64)
    pulse = np.tile(np.exp(-0.5 * ((n - 16.0) / 3.0) ** 2), (len(src_xyz), 1))
    prm = oracle.params(LAM, 0, True)
    rec = [[3, 3, Z - 3], [2, 5, Z - 2]]
    r_dif, _ = oracle.run_dif(pos, m, 0, prm, tab, 0, src_xyz, [0] * 36, pulse, rec, 600, 1)
    r_ref, _ = oracle.run(pos, m, 0, prm, np.repeat(tab[:, :1], 20, axis=1), src_xyz, [0] * 36, pulse, rec, 600, 1, 0, 0, 0)
    assert np.abs(r_ref).max() > 0 and np.array_equal(r_dif, r_ref)
