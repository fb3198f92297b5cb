"""Stress tool: TMA kernel vs plain kernel, whole field, repeated, on grids large enough for >1 CTA per SM and
multi-wave launches (the permanent version of this check is tests/test_gpu_large_grids.py).
usage: python tools/race_check.py [all|small|big]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle
from parallelfdtd_b200 import capi, synth
from tests import fdtd_cases as fc

def run(dims, geom, ut, double, steps, opts, n_mat=5):
    bid, mat = (synth.hall if geom == "hall" else synth.shoebox)(dims, n_mat)
    tab = synth.material_table(list(np.linspace(0.99, 0.5, n_mat))).astype(np.float64 if double else np.float32)
    s = capi.Solver()
    s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
    for k, v in opts:
        s.set_option(k, v)
    lam = fc.LAM
    s.setup_mesh(bid, mat, (32, 4, 1), ut, capi.F64 if double else capi.F32, oracle.params(lam, 0, double), tab)
    s.make_partition(1, [0])
    n = np.arange(steps + 8, dtype=np.float64)
    src = np.exp(-0.5 * ((n - 12.0) / 3.0) ** 2)[None, :]
    s.set_sources([[dims[0] // 2 - 3, dims[1] // 3, dims[2] // 3]], [capi.SRC_HARD], src)
    s.set_receivers([[dims[0] // 2, dims[1] // 2, dims[2] // 2]])
    s.reserve_steps(steps)
    s.enqueue_steps(0, steps)
    s.sync()
    f = (s.export_partition_pressure(0, 0), s.export_partition_pressure(0, 1))
    nm = s.kernel_name()
    s.close()
    return f, nm

def compare(label, dims, geom, ut, double, steps, variants, reps=3):
    (b0, b1), nmb = run(dims, geom, ut, double, steps, [(capi.OPT_KERNEL, capi.KERNEL_PLAIN)])
    for vl, opts in variants:
        bad = 0
        info = ""
        for rep in range(reps):
            (f0, f1), nm = run(dims, geom, ut, double, steps, [(capi.OPT_KERNEL, capi.KERNEL_TMA), (capi.OPT_USE_GRAPH, 0)] + opts)
            d = (f0 != b0) | (f1 != b1)
            if d.any():
                bad += 1
                zz, yy, xx = np.nonzero(d)
                info = f"ndiff={int(d.sum())} z[{zz.min()},{zz.max()}] y[{yy.min()},{yy.max()}] x[{xx.min()},{xx.max()}] max={np.abs(f0-b0).max():.2e}"
        print(f"{label:28s} {vl:26s} bad {bad}/{reps} {info} [{nm}]", flush=True)

H = capi.OPT_TMA_HINTS
C = capi.OPT_TMA_CHUNK
T = capi.OPT_TMA_TILE
if __name__ != "__main__":
    which = "none"
else:
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "small"):
    compare("hall96x128x64 f64 fwd", (96, 128, 64), "hall", 0, True, 80,
            [("chunk4", [(C, 4)]), ("chunk4 tile3(128x16)", [(C, 4), (T, 3)]),
             ("chunk4 tile4(128x8 s6)", [(C, 4), (T, 4)]), ("chunk8", [(C, 8)])], reps=5)
    compare("hall96x128x64 f32 fwd", (96, 128, 64), "hall", 0, False, 80,
            [("chunk4 tile1(128x8)", [(C, 4), (T, 1)]), ("chunk4 tile3(128x16)", [(C, 4), (T, 3)]),
             ("chunk2 tile3", [(C, 2), (T, 3)])], reps=5)
    compare("hall96x128x64 f64 ctr", (96, 128, 64), "hall", 2, True, 80, [("chunk4", [(C, 4)])], reps=5)
if which in ("all", "big"):
    compare("hall256x256x128 f32 fwd", (256, 256, 128), "hall", 0, False, 60, [("auto", []), ("chunk4", [(C, 4)])])
    compare("hall256x256x128 f64 fwd", (256, 256, 128), "hall", 0, True, 60, [("auto", []), ("chunk4", [(C, 4)])])
    compare("shoebox512 f32 fwd", (512, 512, 512), "shoebox", 0, False, 30, [("auto", [])], reps=2)
    compare("shoebox512x512x256 f64 fwd", (512, 512, 256), "shoebox", 0, True, 30, [("auto", [])], reps=2)
    compare("hall256x256x128 f32 iiso", (256, 256, 128), "hall", 3, False, 40, [("auto", []), ("chunk4", [(C, 4)])])
