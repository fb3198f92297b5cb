"""Shared case definitions and runners for the parity tests (CUDA path vs CPU oracle vs golden
vectors produced by the reference's own CUDA build).  Test infrastructure."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from parallelfdtd_b200 import synth  # noqa: E402

# Courant number as the reference's callers get it: every front-end calls setUpdateType, which sets
# lambda = sqrt(1/3) (reference SimulationParameters.cpp:123-137) -- one ulp below the constructor's
# 1/sqrt(3) (SimulationParameters.h:47); its square is exactly the double nearest 1/3.  Identical in
# float, observable in double: the reference-CUDA golden vectors are only reproduced with this value.
LAM = float(np.sqrt(1.0 / 3.0))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def make_case(name, dims, update_type, double, steps, n_mat, n_parts, sources, receivers, geometry="shoebox", octave=0,
              input_data=()):
    """sources: (x, y, z, src_type, input_type, data_idx) final element coordinates."""
    if geometry == "shoebox":
        bid, mat = synth.shoebox(dims, n_mat)
    elif geometry == "hall":
        bid, mat = synth.hall(dims, n_mat)
    else:
        raise ValueError(geometry)
    refl = [0.9] if n_mat == 1 else list(np.linspace(0.99, 0.5, n_mat))
    # octave slots differ so that the material-index arithmetic is observable
    tab = synth.material_table(refl) * (1 + 0.01 * np.arange(20, dtype=np.float32))[None, :]
    return dict(name=name, bid=bid, mat=mat, block=(32, 4, 1), update_type=update_type, double=double, steps=steps,
                octave=octave, n_parts=n_parts, devices=[0] * n_parts, materials=tab.astype(np.float32),
                sources=list(sources), receivers=list(receivers), input_data=list(input_data))


_DATA = [np.sin(np.arange(64) * 0.3)]
_SRC3 = [(10, 10, 3, 0, 0, 0), (10, 12, 23, 0, 1, 0), (11, 10, 40, 1, 3, 0)]
_REC3 = [(20, 12, 5), (20, 12, 24), (20, 12, 42)]


def parity_cases():
    """The cases whose reference-CUDA outputs are committed under tests/golden/."""
    return [
        make_case("c1_shoebox64_fwd_f32", (64, 64, 64), 0, False, 500, 1, 1, [(32, 32, 32, 0, 0, 0)], [(40, 36, 28)]),
        make_case("shoebox64_fwd_f64", (64, 64, 64), 0, True, 300, 1, 1, [(32, 32, 32, 0, 0, 0)], [(40, 36, 28)]),
        make_case("shoebox_48x40x49_ctr_f32_6mat_2parts", (48, 40, 49), 2, False, 300, 6, 2, _SRC3, _REC3, input_data=_DATA),
        make_case("shoebox_48x40x49_ctr_f64_6mat_5parts", (48, 40, 49), 2, True, 300, 6, 5, _SRC3, _REC3, input_data=_DATA),
        make_case("shoebox_48x40x49_fwd_f32_6mat_2parts", (48, 40, 49), 0, False, 300, 6, 2, _SRC3, _REC3, input_data=_DATA),
        make_case("hall_96x128x64_fwd_f32_5mat_oct1", (96, 128, 64), 0, False, 200, 5, 1, [(40, 20, 20, 0, 0, 0)],
                  [(50, 60, 30), (20, 100, 40)], geometry="hall", octave=1),
        make_case("hall_96x128x64_ctr_f32_5mat", (96, 128, 64), 2, False, 200, 5, 2, [(40, 20, 20, 0, 0, 0)],
                  [(50, 60, 30), (20, 100, 40)], geometry="hall"),
        make_case("hall_96x128x64_ctr_f64_5mat_oct2", (96, 128, 64), 2, True, 200, 5, 1, [(40, 20, 20, 0, 0, 0)],
                  [(50, 60, 30), (20, 100, 40)], geometry="hall", octave=2),
    ]


def source_table(case):
    steps = case["steps"]
    dt = np.float64 if case["double"] else np.float32
    tab = np.zeros((len(case["sources"]), steps), dtype=dt)
    for i, s in enumerate(case["sources"]):
        data = case["input_data"][s[5]] if s[4] == 3 else None
        tab[i] = oracle.source_samples(s[4], steps, 7000, data, case["double"])
    return tab


def interp_cases():
    """Interpolated 27-point schemes (IISO = 3, IWB = 4): not in the reference, checked against our own oracle."""
    return [
        make_case("iiso_shoebox_48x40x49_f32_6mat", (48, 40, 49), 3, False, 300, 6, 2, _SRC3, _REC3, input_data=_DATA),
        make_case("iwb_shoebox_48x40x49_f64_6mat", (48, 40, 49), 4, True, 300, 6, 3, _SRC3, _REC3, input_data=_DATA),
        make_case("iiso_hall_96x128x64_f64_5mat_oct1", (96, 128, 64), 3, True, 200, 5, 1, [(40, 20, 20, 0, 0, 0)],
                  [(50, 60, 30), (20, 100, 40)], geometry="hall", octave=1),
        make_case("iwb_hall_96x128x64_f32_5mat", (96, 128, 64), 4, False, 200, 5, 2, [(40, 20, 20, 0, 0, 0)],
                  [(50, 60, 30), (20, 100, 40)], geometry="hall"),
    ]


def dif_table(n_mat, order, seed=7):
    """[n_mat][20] rows [b0..bN, a1..aN]: passive-looking low-order admittance filters (poles well inside the unit circle)."""
    rng = np.random.default_rng(seed)
    t = np.zeros((n_mat, 20), dtype=np.float64)
    for m in range(n_mat):
        y0 = 0.02 + 0.3 * (m + 1) / (n_mat + 1)
        t[m, 0] = y0 * (0.6 + 0.3 * rng.random())
        for i in range(1, order + 1):
            t[m, i] = y0 * 0.25 * (rng.random() - 0.3) / i
            t[m, order + i] = 0.5 * (rng.random() - 0.5) / i
    return t


def dif_cases():
    """Frequency-dependent (digital impedance filter) boundaries: not in the reference, checked against our oracle."""
    out = []
    for name, dims, ut, dbl, steps, n_mat, parts, geom, order in [
            ("dif2_shoebox_48x40x49_fwd_f32", (48, 40, 49), 0, False, 300, 6, 2, "shoebox", 2),
            ("dif4_shoebox_48x40x49_ctr_f64", (48, 40, 49), 2, True, 300, 6, 3, "shoebox", 4),
            ("dif1_hall_96x128x64_fwd_f64", (96, 128, 64), 0, True, 200, 5, 1, "hall", 1),
            ("dif2_hall_96x128x64_iiso_f32", (96, 128, 64), 3, False, 200, 5, 2, "hall", 2),
            ("dif3_shoebox_48x40x49_iwb_f64", (48, 40, 49), 4, True, 300, 6, 1, "shoebox", 3)]:
        if geom == "shoebox":
            c = make_case(name, dims, ut, dbl, steps, n_mat, parts, _SRC3, _REC3, input_data=_DATA)
        else:
            c = make_case(name, dims, ut, dbl, steps, n_mat, parts, [(40, 20, 20, 0, 0, 0)], [(50, 60, 30), (20, 100, 40)], geometry="hall")
        c["dif_order"] = order
        c["materials"] = dif_table(n_mat, order).astype(np.float64 if dbl else np.float32)
        out.append(c)
    return out


def lam_of(case):
    return oracle.interp_lambda(case["update_type"]) if case["update_type"] >= 3 else LAM


def params_of(case, for_oracle):
    """4-entry reference parameter vector; the oracle's interpolated path takes 4 more (d1..d4)."""
    lam = lam_of(case)
    prm = oracle.params(lam, case["octave"], case["double"])
    if case["update_type"] >= 3 and for_oracle:
        d = case.get("dcoef") or oracle.interp_coefficients(case["update_type"], float(prm[1]))
        return oracle.params_interp(lam, case["octave"], d, case["double"])
    return prm


def scheme_of(case):
    if case["update_type"] >= 3:
        return 3
    if case["double"]:
        return 0 if case["update_type"] in (0, 1) else 2          # setupMeshDouble, cudaMesh.cu:134-137
    return 0 if case["update_type"] in (0, 1, 3) else 2           # setupMesh, cudaMesh.cu:70-73


def run_oracle(case, n_parts=None, matidx=1, soft=0, double_pad=False):
    pos, mat, air, bnd = oracle.setup_mesh(case["bid"], case["mat"], case["block"], case["update_type"], case["double"],
                                           double_pad)
    prm = params_of(case, True)
    if case["update_type"] >= 3:
        matidx = 0          # the new schemes index the material table as intended (mat*20 + octave)
    src = np.asarray(case["sources"], dtype=np.int32).reshape(-1, 6)
    if case.get("dif_order") is not None:
        r, secs = oracle.run_dif(pos, mat, scheme_of(case), prm, case["materials"], case["dif_order"], src[:, :3], src[:, 3],
                                 source_table(case), case["receivers"], case["steps"], n_parts or case["n_parts"])
        return r, (pos, mat, air, bnd), secs
    r, secs = oracle.run(pos, mat, scheme_of(case), prm, case["materials"], src[:, :3], src[:, 3], source_table(case),
                         case["receivers"], case["steps"], n_parts or case["n_parts"], matidx, soft)
    return r, (pos, mat, air, bnd), secs


def run_ours(capi, case, n_parts=None, kernel=None, matidx=1, opts=(), devices=None):
    s = capi.Solver()
    try:
        s.set_option(capi.OPT_KERNEL, capi.KERNEL_AUTO if kernel is None else kernel)
        s.set_option(capi.OPT_MATIDX_AS_WRITTEN, matidx)
        for k, v in opts:
            s.set_option(k, v)
        dt = capi.F64 if case["double"] else capi.F32
        prm = params_of(case, False)
        if case.get("dcoef"):
            s.set_scheme_coefficients(case["dcoef"])
        if case.get("dif_order") is not None:
            s.set_option(capi.OPT_DIF_ORDER, case["dif_order"])
        s.setup_mesh(case["bid"], case["mat"], case["block"], case["update_type"], dt, prm, case["materials"])
        n = n_parts or case["n_parts"]
        s.make_partition(n, devices or [0] * n)
        src = np.asarray(case["sources"], dtype=np.int32).reshape(-1, 6)
        s.set_sources(src[:, :3], src[:, 3], source_table(case))
        s.set_receivers(case["receivers"])
        t0 = time.time()
        r, sps = s.run(case["steps"])
        wall = time.time() - t0
        nodes = [s.export_partition_nodes(k) for k in range(n)]
        info = dict(kernel=s.kernel_name(), wall=wall, counts=s.counts(), dims=s.dims(), launches=s.launch_count(),
                    partitions=[s.partition(k)[:2] for k in range(n)])
    finally:
        s.close()
    return r, nodes, info


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))


def node_checksum(a):
    """Position-weighted 64-bit checksum of a node-byte volume (used in the golden fixtures)."""
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    w = (1 + np.arange(a.size, dtype=np.uint64) % np.uint64(251))
    return int(np.sum(a.astype(np.uint64) * w, dtype=np.uint64))


def load_golden(name):
    p = os.path.join(GOLDEN_DIR, name + ".npz")
    if not os.path.exists(p):
        return None
    return dict(np.load(p))
