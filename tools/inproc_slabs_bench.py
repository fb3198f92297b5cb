"""Slabs in ONE process over several GPUs (the reference's own multi-GPU mode, CudaMesh::makePartition(n, devices)):
step time with the halo planes stored by the edge launches (peer-mapped stores) vs copied afterwards.
usage: python tools/inproc_slabs_bench.py [n_gpus] [steps]"""
import json
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from parallelfdtd_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dims = (512, 512, 512 * n)
bid, mat = synth.shoebox(dims, 6)
lam = float(np.sqrt(1.0 / 3.0))
prm = np.array([lam, lam * lam, 1.0 / 3.0, 0.0], dtype=np.float32)
out = {"n_gpus": n, "dims": dims, "steps": steps}
for dif in (2, 0):
    tab = (synth.filter_material_table(list(np.linspace(0.99, 0.5, 6)), dif) if dif else synth.material_table(list(np.linspace(0.99, 0.5, 6)))).astype(np.float32)
    resp = {}
    for peer in (1, 0):
        s = capi.Solver()
        s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
        s.set_option(capi.OPT_DIF_ORDER, dif)
        s.set_option(capi.OPT_PEER_STORES, peer)
        s.setup_mesh(bid, mat, (32, 4, 1), capi.SRL_FORWARD, capi.F32, prm, tab)
        s.make_partition(n, list(range(n)))
        t = np.arange(steps + 40, dtype=np.float64)
        s.set_sources([[256, 256, dims[2] // 2]], [capi.SRC_HARD], np.exp(-0.5 * ((t - 40.0) / 6.0) ** 2).astype(np.float32)[None, :])
        s.set_receivers([[270, 260, 3 + (i * (dims[2] - 6)) // 3] for i in range(4)])
        s.reserve_steps(steps + 40)
        s.enqueue_steps(0, 20)
        s.sync()
        s.enqueue_steps(20, steps)
        s.sync()
        ms, _, _ = s.last_timing()
        resp[peer] = s.fetch_responses(steps + 20) if hasattr(s, "fetch_responses") else None
        X, Y, Z = s.dims()
        out[f"dif{dif}_peer{peer}"] = {"ms_per_step": ms / steps, "Mvox_s": X * Y * Z * steps / (ms * 1e-3) / 1e6, "kernel": s.kernel_name()}
        s.close()
    if resp[0] is not None:
        out[f"dif{dif}_identical"] = bool(np.array_equal(resp[0], resp[1]))
print(json.dumps(out))
