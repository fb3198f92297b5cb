"""Synthetic voxelized rooms: voxelizer-style ``bid`` + material byte volumes.

The reference obtains these two byte volumes from the third-party Voxelizer
(reference src/kernels/voxelizationUtils.cu:55-136), which is not vendored.  Here
they are generated directly from an inside/outside predicate, using the ``bid``
semantics recovered from the reference's own comments (src/kernels/cudaMesh.cu:372-476;
SURVEY.md Appendix B): ``bid`` 0 = solid, 27 = air with six air neighbours, 1..26 =
boundary node whose listed neighbours are air.  Axis names: Left=x-1, Right=x+1,
In=y-1, Out=y+1, Down=z-1, Up=z+1.

Volumes are C-ordered ``[z][y][x]`` uint8, i.e. linear index ``z*X*Y + y*X + x`` like
the reference (src/kernels/cudaMesh.h:247-249).  Every generator can emit a z-range
of the global volume so that each z-slab owner builds only its own part.
"""
from __future__ import annotations

import numpy as np

_L, _R, _IN, _OUT, _D, _U = 1, 2, 4, 8, 16, 32

# air-neighbour sets per bid (cudaMesh.cu:372-476)
_BID_SETS = {
    1: _D | _L | _IN, 2: _D | _R | _IN, 3: _D | _L | _OUT, 4: _D | _R | _OUT,
    5: _U | _L | _IN, 6: _U | _R | _IN, 7: _U | _L | _OUT, 8: _U | _R | _OUT,
    9: _D | _L | _R | _IN, 10: _D | _L | _R | _OUT, 11: _D | _L | _IN | _OUT, 12: _D | _R | _IN | _OUT,
    13: _U | _L | _R | _IN, 14: _U | _L | _R | _OUT, 15: _U | _L | _IN | _OUT, 16: _U | _R | _IN | _OUT,
    17: _U | _D | _L | _IN, 18: _U | _D | _R | _IN, 19: _U | _D | _L | _OUT, 20: _U | _D | _R | _OUT,
    21: _L | _R | _IN | _OUT | _D, 22: _L | _R | _OUT | _D | _U, 23: _L | _R | _IN | _D | _U,
    24: _R | _IN | _OUT | _D | _U, 25: _L | _IN | _OUT | _D | _U, 26: _L | _R | _IN | _OUT | _U,
    27: _L | _R | _IN | _OUT | _D | _U,
}
_MASK_TO_BID = np.full(64, 255, dtype=np.uint8)
for _b, _m in _BID_SETS.items():
    _MASK_TO_BID[_m] = _b


def bid_from_inside(inside_fn, dims, z0=0, z1=None, strict=True):
    """``bid`` volume for global slices ``[z0, z1)``.

    ``inside_fn(x, y, z)`` takes broadcastable int arrays of *global* voxel coordinates and
    returns a bool array (True = air).  Coordinates outside ``dims`` count as solid.
    Non-representable nodes (air on neither side of an axis, etc.) raise when ``strict``,
    else they become solid (bid 0).
    """
    X, Y, Z = dims
    z1 = Z if z1 is None else z1
    x = np.arange(-1, X + 1)[None, None, :]
    y = np.arange(-1, Y + 1)[None, :, None]
    z = np.arange(z0 - 1, z1 + 1)[:, None, None]
    ins = np.asarray(inside_fn(x, y, z), dtype=bool)
    ins = np.broadcast_to(ins, (z1 - z0 + 2, Y + 2, X + 2)).copy()
    ins &= (x >= 0) & (x < X) & (y >= 0) & (y < Y) & (z >= 0) & (z < Z)
    while True:
        c = ins[1:-1, 1:-1, 1:-1]
        mask = (ins[1:-1, 1:-1, :-2] * np.uint8(_L)
                + ins[1:-1, 1:-1, 2:] * np.uint8(_R)
                + ins[1:-1, :-2, 1:-1] * np.uint8(_IN)
                + ins[1:-1, 2:, 1:-1] * np.uint8(_OUT)
                + ins[:-2, 1:-1, 1:-1] * np.uint8(_D)
                + ins[2:, 1:-1, 1:-1] * np.uint8(_U)).astype(np.uint8)
        bid = _MASK_TO_BID[mask]
        bid[~c] = 0
        bad = bid == 255
        if not bad.any():
            break
        if strict:
            raise ValueError(f"{int(bad.sum())} air voxels have a neighbour set with no bid code")
        # make them solid and re-classify the neighbours (only exact when the whole z-range is
        # generated at once; slab-wise callers should use geometry that needs no repair)
        ins[1:-1, 1:-1, 1:-1][bad] = False
    return np.ascontiguousarray(bid)


# ---------------------------------------------------------------------------------------
# rooms
# ---------------------------------------------------------------------------------------
def shoebox_inside(dims, shell=1):
    X, Y, Z = dims

    def fn(x, y, z):
        return ((x >= shell) & (x < X - shell) & (y >= shell) & (y < Y - shell)
                & (z >= shell) & (z < Z - shell))
    return fn


def shoebox_generic(dims, n_materials=1, z0=0, z1=None, shell=1):
    """The shoebox through the general predicate path (bid_from_inside); `shoebox` is its separable fast form."""
    X, Y, Z = dims
    z1 = Z if z1 is None else z1
    bid = bid_from_inside(shoebox_inside(dims, shell), dims, z0, z1)
    mat = np.zeros_like(bid)
    if n_materials > 1:
        x = np.arange(X)[None, None, :]
        y = np.arange(Y)[None, :, None]
        z = np.arange(z0, z1)[:, None, None]
        face = np.full(bid.shape, 255, dtype=np.uint8)
        for idx, cond in reversed(list(enumerate([x == shell, x == X - shell - 1, y == shell, y == Y - shell - 1,
                                                   z == shell, z == Z - shell - 1]))):
            face = np.where(np.broadcast_to(cond, bid.shape), np.uint8(idx), face)
        bnd = (bid > 0) & (bid < 27)
        mat[bnd] = face[bnd] % n_materials
    return bid, mat


def shoebox(dims, n_materials=1, z0=0, z1=None, shell=1):
    """Shoebox with a ``shell``-voxel solid rim (matches the +1 source/receiver padding,
    reference SimulationParameters.cpp:200-208).  Returns (bid, mat) for slices [z0,z1).
    Materials: with ``n_materials`` == 6 one per face (x-, x+, y-, y+, z-, z+; edges/corners take
    the first face in that order); otherwise face index modulo ``n_materials``.

    A box is separable: a voxel is air iff it is inside on every axis, and its air-neighbour set is the union of
    three per-axis bit pairs, so the volumes are three 1-D tables combined by broadcasting (a few passes over the
    volume instead of the general path's twenty; a 1e9-voxel room takes seconds)."""
    X, Y, Z = dims
    z1 = Z if z1 is None else z1

    def axis(n, lo, hi, bit_minus, bit_plus, first_face):
        c = np.arange(lo, hi)
        ins = (c >= shell) & (c < n - shell)
        minus = (c - 1 >= shell) & (c - 1 < n - shell)          # neighbour c-1 is air (given the other axes are inside)
        plus = (c + 1 >= shell) & (c + 1 < n - shell)
        bits = (minus * np.uint8(bit_minus) + plus * np.uint8(bit_plus)).astype(np.uint8)
        face = np.full(c.shape, 255, dtype=np.uint8)
        face[c == n - shell - 1] = first_face + 1
        face[c == shell] = first_face
        return ins, bits, face

    ix, bx, fx = axis(X, 0, X, _L, _R, 0)
    iy, by, fy = axis(Y, 0, Y, _IN, _OUT, 2)
    iz, bz, fz = axis(Z, z0, z1, _D, _U, 4)
    mask = bx[None, None, :] | by[None, :, None] | bz[:, None, None]
    bid = _MASK_TO_BID[mask]
    inside = ix[None, None, :] & iy[None, :, None] & iz[:, None, None]
    bid *= inside                                               # solid outside; inside masks always have a code for shell >= 1 rooms
    if (bid == 255).any():
        return shoebox_generic(dims, n_materials, z0, z1, shell)   # degenerate (one-voxel-wide) rooms
    mat = np.zeros_like(bid)
    if n_materials > 1:
        face = np.minimum(np.minimum(fx[None, None, :], fy[None, :, None]), fz[:, None, None])
        bnd = (bid > 0) & (bid < 27)
        np.copyto(mat, face % np.uint8(n_materials), where=bnd)
    return np.ascontiguousarray(bid), mat


def hall_inside(dims):
    """Concert-hall-like room inside ``dims``: a shoebox shell, a raked (stepped) floor rising
    towards +y, a balcony slab on the back wall with its own solid underside, and four
    rectangular columns.  All features are >= 2 voxels thick/wide so every air node has a
    representable ``bid``."""
    X, Y, Z = dims

    def fn(x, y, z):
        box = (x >= 1) & (x < X - 1) & (y >= 1) & (y < Y - 1) & (z >= 1) & (z < Z - 1)
        # raked floor: steps of 4 voxels depth in y, rising 1 voxel each, over the rear 60 % of the hall
        y_start = int(0.4 * Y)
        rise = np.maximum(0, (y - y_start) // 4 * 2) // 2
        rise = np.minimum(rise, Z // 4)
        floor = z >= 1 + np.where(y >= y_start, 2 * ((rise + 1) // 2), 0)
        # balcony: solid slab between z in [bz0, bz0+bt) for y >= by0
        bz0, bt, by0 = int(0.6 * Z), max(2, Z // 32), int(0.8 * Y)
        balcony = (y >= by0) & (z >= bz0) & (z < bz0 + bt)
        # columns: 4 square pillars, side cw, full height, in the front half
        cw = max(2, X // 32)
        cols = np.zeros(np.broadcast(x, y, z).shape, dtype=bool)
        for cx in (X // 4, 3 * X // 4):
            for cy in (Y // 8, Y // 4 + Y // 16):
                cols = cols | ((x >= cx) & (x < cx + cw) & (y >= cy) & (y < cy + cw))
        return box & floor & ~balcony & ~cols
    return fn


def hall(dims, n_materials=5, z0=0, z1=None):
    """(bid, mat) of the synthetic hall; materials banded by height and orientation."""
    X, Y, Z = dims
    z1 = Z if z1 is None else z1
    bid = bid_from_inside(hall_inside(dims), dims, z0, z1, strict=False)
    mat = np.zeros_like(bid)
    if n_materials > 1:
        z = np.arange(z0, z1)[:, None, None]
        band = (z * n_materials // max(Z, 1)).astype(np.uint8)
        bnd = (bid > 0) & (bid < 27)
        horiz = (bid == 21) | (bid == 26)     # floor / ceiling faces
        m = np.where(horiz, np.uint8(0), np.broadcast_to(band, bid.shape) % np.uint8(n_materials))
        mat[bnd] = m[bnd]
    return bid, mat


def banded_shoebox(dims, n_materials=20, z0=0, z1=None):
    """Shoebox whose wall material changes in ``n_materials`` z-bands (BASELINE config 5)."""
    X, Y, Z = dims
    z1 = Z if z1 is None else z1
    bid = bid_from_inside(shoebox_inside(dims), dims, z0, z1)
    mat = np.zeros_like(bid)
    z = np.arange(z0, z1)[:, None, None]
    band = np.minimum(z * n_materials // max(Z, 1), n_materials - 1).astype(np.uint8)
    bnd = (bid > 0) & (bid < 27)
    mat[bnd] = np.broadcast_to(band, bid.shape)[bnd]
    return bid, mat


def reflection_to_admittance(r):
    """reference src/global_includes.h:29-32 (float arithmetic)."""
    r = np.float32(r)
    return np.float32((np.float32(1) - r) / (np.float32(1) + r))


def material_table(reflectances, n_coef=20):
    """[n_mat][20] float32 admittance table, every octave slot the same value
    (layout: reference src/base/MaterialHandler.cpp:100-130)."""
    t = np.zeros((len(reflectances), 20), dtype=np.float32)
    for i, r in enumerate(reflectances):
        t[i, :n_coef] = reflection_to_admittance(r)
    return t


def filter_material_table(reflectances, order, n_coef=20):
    """[n_mat][20] float64 table of digital-impedance-filter admittances, row = [b0..bN, a1..aN] (PFDTD_OPT_DIF_ORDER).

    Material m is a one-pole-per-order low-shelf around its frequency-independent admittance Y0 (derived from the
    reflectance like `material_table`): Y(z) = Y0 * prod_i (1 - q_i z^-1) / (1 - p_i z^-1) with real poles/zeros
    well inside the unit circle, so that Re Y(e^jw) > 0 (passive) and the DC / Nyquist values stay within a factor
    of two of Y0.  Synthetic stand-in for fitted wall impedances (BASELINE config 2).
    """
    assert 1 <= order <= 4 and 2 * order + 1 <= n_coef
    t = np.zeros((len(reflectances), n_coef), dtype=np.float64)
    for m, r in enumerate(reflectances):
        y0 = float(reflection_to_admittance(np.float32(r)))
        b = np.array([1.0])
        a = np.array([1.0])
        for i in range(order):
            p = 0.55 - 0.12 * i + 0.02 * (m % 3)
            q = p - 0.18 / (i + 1)
            b = np.convolve(b, [1.0, -q])
            a = np.convolve(a, [1.0, -p])
        t[m, 0:order + 1] = y0 * b
        t[m, order + 1:2 * order + 1] = a[1:]
    return t
