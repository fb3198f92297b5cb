"""GPU: the device voxeliser (pfdtd_voxelize*; reference src/kernels/voxelizationUtils.cu:47-146, whose third-party
back end is not vendored -- parity unpinned by the reference) against an independent numpy classification of
analytically known rooms, and through FDTD::App against the same room given as node volumes."""
import numpy as np
import pytest

from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu


def box(lo, hi):
    """12 triangles of the axis-aligned box [lo, hi], outward orientation not required (parity fill)."""
    (x0, y0, z0), (x1, y1, z1) = lo, hi
    v = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], dtype=np.float32)
    q = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]
    t = np.array([tri for a, b, c, d in q for tri in ((a, b, c), (a, c, d))], dtype=np.uint32)
    return v, t


def expected_bid(shape, inside):
    """bid volume from an inside predicate over lattice points p = (i-1) dx, classified by synth.bid_from_inside."""
    vz, vy, vx = shape
    return synth.bid_from_inside(inside, (vx, vy, vz), strict=False)


@pytest.mark.parametrize("dims_m,dx", [((1.0, 1.0, 1.0), 0.1), ((2.0, 1.3, 0.9), 0.05), ((0.64, 0.48, 0.8), 0.02)])
def test_box_room_matches_the_analytic_classification(capi, gpu, dims_m, dx):
    v, t = box((0, 0, 0), dims_m)
    tri_mat = (np.arange(len(t)) // 2).astype(np.uint8)          # one material per face
    bid, mat = capi.voxelize(v, t, dx, tri_mat)
    assert bid.shape == tuple(int(np.ceil(np.float32(d) / np.float32(dx))) + 3 for d in dims_m[::-1])
    f32 = np.float32

    def inside(x, y, z):     # lattice point strictly inside the box (points on the surface are solid)
        px, py, pz = (x - 1).astype(f32) * f32(dx), (y - 1).astype(f32) * f32(dx), (z - 1).astype(f32) * f32(dx)
        e = f32(3e-4) * f32(dx)
        return ((px > e) & (px < f32(dims_m[0]) - e) & (py > e) & (py < f32(dims_m[1]) - e) & (pz > e) & (pz < f32(dims_m[2]) - e))

    exp = expected_bid(bid.shape, inside)
    assert np.array_equal(bid, exp)
    assert (bid == 27).sum() > 0 and ((bid > 0) & (bid < 27)).sum() > 0
    # materials: boundary voxels of a face carry that face's material (corners/edges: nearest triangle), others 0
    assert (mat[(bid == 0) | (bid == 27)] == 0).all()
    zc, yc = bid.shape[0] // 2, bid.shape[1] // 2
    assert mat[zc, yc, 2] == tri_mat[10] and mat[zc, yc, np.nonzero(bid[zc, yc])[0][-1]] == tri_mat[8]   # x = 0 / x = L faces
    assert mat[2, yc, bid.shape[2] // 2] == tri_mat[0] and mat[np.nonzero(bid[:, yc, bid.shape[2] // 2])[0][-1], yc, bid.shape[2] // 2] == tri_mat[2]


@pytest.mark.parametrize("dims_m,dx", [((10.0, 10.0, 3.0), 0.25), ((6.0, 2.0, 9.0), 0.2)])
def test_every_face_voxel_of_a_flat_room_carries_its_own_face_material(capi, gpu, dims_m, dx):
    """Coarse room mesh (two triangles per face), low aspect ratio: a floor voxel next to a wall is nearer to the wall
    triangles' CENTROIDS than to the floor triangles' -- the material must come from the nearest TRIANGLE.  Every face
    voxel is checked (bid 21..26 = exactly one missing neighbour), not only the face centres."""
    v, t = box((0, 0, 0), dims_m)
    tri_mat = (1 + np.arange(len(t)) // 2).astype(np.uint8)      # faces in box() order: z=0, z=top, y=0, y=max, x=max, x=0
    bid, mat = capi.voxelize(v, t, dx, tri_mat)
    face_of_bid = {26: 1, 21: 2, 22: 3, 23: 4, 25: 5, 24: 6}     # missing Down / Up / In (y-1) / Out (y+1) / Right (x+1) / Left (x-1)
    for b, m in face_of_bid.items():
        sel = bid == b
        assert sel.sum() > 0, b
        assert (mat[sel] == m).all(), (b, m, np.unique(mat[sel], return_counts=True))
    # edges and corners belong to one of the faces they touch
    edge = (bid > 0) & (bid < 21)
    assert edge.sum() > 0 and (mat[edge] >= 1).all() and (mat[edge] <= 6).all()


def test_l_shaped_room_and_thin_features(capi, gpu):
    """Union of two boxes sharing a face region is not a closed 2-manifold; use an L-shaped prism built from its outline."""
    # L-shaped outline in xy (metres), extruded 0..1 in z
    pts = np.array([[0, 0], [2, 0], [2, 1], [1, 1], [1, 2], [0, 2]], dtype=np.float32)
    n = len(pts)
    v = np.array([[x, y, 0] for x, y in pts] + [[x, y, 1] for x, y in pts], dtype=np.float32)
    floor = [(0, 1, 2), (0, 2, 3), (0, 3, 4), (0, 4, 5)]                 # fan is valid for this outline (star-shaped from vertex 0)
    t = [tri for tri in floor] + [(a + n, b + n, c + n) for a, b, c in floor]
    for i in range(n):
        j = (i + 1) % n
        t += [(i, j, j + n), (i, j + n, i + n)]
    t = np.array(t, dtype=np.uint32)
    dx = 0.05
    bid, _ = capi.voxelize(v, t, dx)
    f32 = np.float32

    def inside(x, y, z):
        px, py, pz = (x - 1).astype(f32) * f32(dx), (y - 1).astype(f32) * f32(dx), (z - 1).astype(f32) * f32(dx)
        e = f32(3e-4) * f32(dx)
        in_z = (pz > e) & (pz < 1 - e)
        a = (px > e) & (px < 2 - e) & (py > e) & (py < 1 - e)          # lower bar
        b = (px > e) & (px < 1 - e) & (py > e) & (py < 2 - e)          # left bar
        return (a | b) & in_z

    exp = expected_bid(bid.shape, inside)
    assert np.array_equal(bid, exp)


def test_app_voxelizes_on_the_device_and_matches_node_volumes(capi, gpu):
    """libPyFDTD: a box given as triangles and the same box given as node volumes produce identical responses."""
    import os
    import sys
    from parallelfdtd_b200 import build
    build.build_py_module()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "parallelfdtd_b200"))
    import libPyFDTD as pf
    v, t = box((0, 0, 0), (1.2, 1.0, 0.8))
    fs = 8000
    dx = np.float32(344.0) / np.float32(fs) * np.float32(np.sqrt(3.0))
    out = []
    for as_volumes in (False, True):
        app = pf.App()
        app.initializeDevices()
        app.initializeGeometryPy(t.flatten().tolist(), v.flatten().tolist())
        app.setUpdateType(0)
        app.setNumSteps(100)
        app.setSpatialFs(fs)
        app.forcePartitionTo(1)
        app.addSurfaceMaterials([0.8] * (len(t) * 20), len(t), 20)
        app.addSource(0.5, 0.5, 0.4, 0, 0, 0)
        app.addReceiver(0.7, 0.6, 0.5)
        if as_volumes:
            bid, mat = capi.voxelize(v, t, float(app.getDx()), np.zeros(len(t), dtype=np.uint8))
            app.setVoxelVolumes(bid, mat)
        app.runSimulation()
        out.append(np.array(app.getResponse(0)))
        app.close()
    assert np.abs(out[0]).max() > 0 and np.array_equal(out[0], out[1])
