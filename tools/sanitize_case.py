"""Small solves through the C ABI for `compute-sanitizer` (memcheck / racecheck / synccheck / initcheck):
every kernel family once -- 7-point and 27-point TMA kernels with and without filter boundaries, fp32 and fp64,
the plain kernels, two slabs with peer-stored halos, slice / mesh capture, the device voxeliser.  No torch import.
usage (GPU box): compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_case.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from parallelfdtd_b200 import capi, synth  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 12
LAM = {0: float(np.sqrt(1.0 / 3.0)), 2: float(np.sqrt(1.0 / 3.0)), 3: float(np.sqrt(3.0) / 2.0), 4: 1.0}


def solve(dims, geom, ut, double, dif, parts, kernel, graph=1):
    npdt = np.float64 if double else np.float32
    n_mat = 5
    bid, mat = (synth.hall if geom == "hall" else synth.shoebox)(dims, n_mat)
    refl = list(np.linspace(0.99, 0.5, n_mat))
    tab = (synth.filter_material_table(refl, dif) if dif else synth.material_table(refl)).astype(npdt)
    lam = LAM[ut]
    prm = np.array([lam, lam * lam, 1.0 / 3.0, 0.0], dtype=npdt)
    s = capi.Solver()
    s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
    s.set_option(capi.OPT_KERNEL, kernel)
    s.set_option(capi.OPT_USE_GRAPH, graph)
    s.set_option(capi.OPT_DIF_ORDER, dif)
    s.setup_mesh(bid, mat, (32, 4, 1), ut, capi.F64 if double else capi.F32, prm, tab)
    s.make_partition(parts, [0] * parts)
    n = np.arange(STEPS, dtype=np.float64)
    src = np.exp(-0.5 * ((n - 1.0) / 1.0) ** 2).astype(npdt)[None, :]
    s.set_sources([[dims[0] // 2, dims[1] // 2, dims[2] // 2]], [capi.SRC_SOFT], src)
    s.set_receivers([[dims[0] // 2 + 3, dims[1] // 2, dims[2] // 2 + 2], [5, 6, dims[2] - 4]])
    r, _ = s.run(STEPS)
    p, b = s.capture_slice(dims[1] // 2, 1, with_position=True)
    f = s.capture_mesh()
    name = s.kernel_name()
    s.close()
    assert np.isfinite(r).all() and np.isfinite(f).all() and np.abs(f).max() > 0 and p.shape[0] == f.shape[0]
    print(f"ok {name:60s} {geom} {dims} dif={dif} parts={parts} graph={graph}", flush=True)


if __name__ == "__main__":
    T, P = capi.KERNEL_TMA, capi.KERNEL_PLAIN
    for args in [((96, 40, 30), "shoebox", 0, False, 0, 1, T), ((96, 40, 30), "shoebox", 0, False, 2, 2, T),
                 ((96, 128, 24), "hall", 0, True, 4, 2, T), ((96, 40, 30), "shoebox", 2, False, 2, 1, T),
                 ((96, 40, 30), "shoebox", 2, True, 0, 3, T), ((96, 128, 24), "hall", 3, False, 2, 2, T),
                 ((96, 40, 30), "shoebox", 4, True, 3, 1, T), ((96, 40, 30), "shoebox", 3, True, 0, 2, T),
                 ((48, 40, 30), "shoebox", 0, False, 0, 2, P), ((48, 40, 30), "shoebox", 3, True, 0, 1, P),
                 ((96, 40, 30), "shoebox", 0, False, 2, 2, T, 0)]:
        solve(*args)
    # device voxeliser: 1 m box at dx = 0.1
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.float32)
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]
    tri = np.array([t for a, b, c, d in quads for t in ((a, b, c), (a, c, d))], dtype=np.uint32)
    bid, mat = capi.voxelize(v, tri, 0.1, np.arange(12, dtype=np.uint8) // 2)
    assert bid.shape == (13, 13, 13) and int((bid == 27).sum()) > 0
    print("ok voxelize", bid.shape, flush=True)
