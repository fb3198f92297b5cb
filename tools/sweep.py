"""Kernel-variant sweep on one GPU (developer tool): geometry is built once, every option set is timed
with per-launch CUDA events.  usage: python tools/sweep.py [--dims 512 512 512] [--dtype f32] [--steps 40]
 --sets "tile=3,chunk=0,hints=0;tile=3,chunk=64,hints=1;..."
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallelfdtd_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=3, default=[512, 512, 512])
ap.add_argument("--dtype", default="f32")
ap.add_argument("--update-type", type=int, default=0)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--sets", default="tile=0")
a = ap.parse_args()

dims = tuple(a.dims)
double = a.dtype == "f64"
dt = capi.F64 if double else capi.F32
npdt = np.float64 if double else np.float32
bid, mat = synth.shoebox(dims, 6)
tab = synth.material_table(list(np.linspace(0.99, 0.5, 6))).astype(npdt)
lam = float(np.sqrt(1 / 3)) if a.update_type < 3 else (float(np.sqrt(3.0) / 2) if a.update_type == 3 else 1.0)
prm = np.array([lam, lam * lam, 1 / 3, 0], dtype=npdt)
bpv = 25 if double else 13
for spec in a.sets.split(";"):
    kv = dict(x.split("=") for x in spec.split(",") if x)
    try:
        s = capi.Solver()
        kern = kv.get("kernel", "tma")
        s.set_option(capi.OPT_KERNEL, capi.KERNEL_PLAIN if kern == "plain" else capi.KERNEL_TMA)
        s.set_option(capi.OPT_TMA_TILE, int(kv.get("tile", 0)))
        s.set_option(capi.OPT_TMA_CHUNK, int(kv.get("chunk", 0)))
        s.set_option(capi.OPT_TMA_HINTS, int(kv.get("hints", 0)))
        s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
        s.set_option(capi.OPT_TIME_KERNELS, 1)
        dif = int(kv.get("dif", 0))
        s.set_option(capi.OPT_DIF_ORDER, dif)
        if dif:
            t = np.zeros((6, 20), dtype=npdt)
            t[:, 0] = tab[:, 0] * 0.7
            for i in range(1, dif + 1):
                t[:, i] = tab[:, 0] * 0.1 / i
                t[:, dif + i] = -0.3 / i
            tab_use = t
        else:
            tab_use = tab
        s.setup_mesh(bid, mat, (32, 4, 1), a.update_type, dt, prm, tab_use)
        s.make_partition(1, [0])
        c = [d // 2 for d in dims]
        src = np.zeros((1, a.steps + 16), dtype=npdt)
        src[0, 1] = 1
        s.set_sources([c], [0], src)
        s.set_receivers([[c[0] + 5, c[1], c[2]]])
        s.enqueue_steps(0, 8)
        s.sync()
        s.enqueue_steps(8, a.steps)
        s.sync()
        tot, kms, nk = s.last_timing()
        X, Y, Z = s.dims()
        upd = X * Y * (Z - 2)
        print(json.dumps(dict(spec=spec, kernel=s.kernel_name(), kernel_ms=round(kms / max(nk, 1), 4),
                              gbs=round(upd * bpv / (kms / max(nk, 1) * 1e-3) / 1e9, 1),
                              mvox_s=round(X * Y * Z * a.steps / (tot * 1e-3) / 1e6))), flush=True)
        s.close()
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(spec=spec, error=repr(e))), flush=True)
