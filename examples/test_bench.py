"""The workflow of the reference's python/testBench.py:110-149 on this build: JSON geometry -> per-layer materials ->
libPyFDTD.App -> responses -> post-filter.  Differences from the reference script: the geometry / filter helpers come
from parallelfdtd_b200.postfilter, `--filters` switches the walls to order-2 digital impedance filters (not in the
reference), captures come back as numpy arrays instead of TGA files, and nothing is plotted.

    python examples/test_bench.py [--double] [--filters] [--captures] [--scheme 0|2|3|4] [--steps N] [--fs HZ]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "parallelfdtd_b200"))


def reflection2admittance(r):
    return (1.0 - r) / (1.0 + r)


def run(double=False, filters=False, captures=False, scheme=0, steps=400, fs=20000, geometry=None, verbose=True):
    from parallelfdtd_b200 import build, postfilter, synth
    build.build_py_module()
    import libPyFDTD as pf

    vertices, indices, layers = postfilter.load_json_geometry(geometry or os.path.join(ROOT, "examples", "unit_box.json"))
    n_tri = len(indices)
    app = pf.App()
    app.initializeDevices()
    app.initializeGeometryPy(indices.flatten().tolist(), vertices.flatten().tolist())
    for name, tris in layers.items():
        app.setLayerIndices(tris, name)
    app.setUpdateType(scheme)
    app.setNumSteps(int(steps))
    app.setSpatialFs(int(fs))
    app.setDouble(bool(double))
    app.forcePartitionTo(1)
    refl = {"floor": 0.8, "ceiling": 0.9, "walls": 0.95}
    if filters:
        order = 2
        names = list(layers)
        table = synth.filter_material_table([refl.get(n, 0.99) for n in names], order).astype(np.float32)
        rows = np.zeros((n_tri, 2 * order + 1), dtype=np.float32)
        for k, n in enumerate(names):
            rows[layers[n], :] = table[k, :2 * order + 1]
        app.addSurfaceFilters(rows.flatten().tolist(), n_tri, order)
    else:
        materials = np.ones((n_tri, 20)) * reflection2admittance(0.99)
        for n, tris in layers.items():
            materials[tris, :] = reflection2admittance(refl.get(n, 0.99))
        app.addSurfaceMaterials(materials.flatten().tolist(), n_tri, 20)
    app.addSource(0.5, 0.5, 0.5, 0, 1, 0)                       # hard source, Gaussian pulse
    rec = [[0.6, 0.6, 0.6], [0.4, 0.4, 0.4]]
    for r in rec:
        app.addReceiver(*r)
    if captures:
        app.addSliceToCapture(int(0.5 / app.getDx()) + 1, steps // 4, 1)
        app.runCapture()
    else:
        app.runSimulation()
    get = app.getResponseDouble if (double and not captures) else app.getResponse
    ret = np.transpose(np.array([get(i) for i in range(len(rec))]))
    out = {"responses": ret, "filtered": postfilter.FDTDfilter(ret, float(fs), 0, 0.2), "mvox": app.getMvox(), "elements": app.getNumElems(),
           "dims": app.getDims(), "slices": [app.getSliceCapture(i) for i in range(app.getNumberOfSliceCaptures())]}
    app.close()
    if verbose:
        print(f"{out['elements']} voxels {out['dims']}, {steps} steps, {out['mvox']:.0f} Mvox/s; response peak {np.abs(ret).max():.4g}, "
              f"filtered peak {np.abs(out['filtered']).max():.4g}, {len(out['slices'])} slice capture(s)")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--double", action="store_true")
    ap.add_argument("--filters", action="store_true")
    ap.add_argument("--captures", action="store_true")
    ap.add_argument("--scheme", type=int, default=0)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--fs", type=int, default=20000)
    ap.add_argument("--geometry", default=None)
    a = ap.parse_args()
    run(a.double, a.filters, a.captures, a.scheme, a.steps, a.fs, a.geometry)
