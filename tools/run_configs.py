#!/usr/bin/env python
"""BASELINE.json configurations 3 and 5 at their STATED size on the GPUs of one box (one process per GPU, z-slabs):

  c3  IISO (27-point) on the synthetic concert hall, 1536 x 1024 x 960 = 1.51e9 voxels, 5 materials, 16 receivers on a
      seating grid, fp32
  c5  fp64 IISO, shoebox 2048 x 1024 x 1920 = 4.03e9 voxels, 20 materials in z-bands, 10 source positions as 10
      separate runs (pfdtd_reset_pressures between them)

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_configs.py --config c3 c5

One JSON line per configuration on stdout (rank 0).  Parity of the same configurations at oracle-sized scale:
tests/test_gpu_configs.py; here the checks are size-independent ones (finite, every receiver reached, ten different
responses, and -- c3 -- the same responses with the NCCL transport instead of peer-mapped stores)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", nargs="+", default=["c3", "c5"])
    ap.add_argument("--steps", type=int, default=300, help="steps per run (c5: per source; BASELINE says 4000, see --full)")
    ap.add_argument("--c3-steps", type=int, default=900, help="c3 runs longer so that the seating grid is reached")
    ap.add_argument("--full", action="store_true", help="c5 with the 4000 steps per source BASELINE.json names")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every dimension (smoke runs)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from parallelfdtd_b200 import capi, slabs, synth

    rank, world, lr = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

    def barrier():
        if world > 1:
            dist.barrier()

    def maxr(v):
        return slabs.max_over_ranks(v) if world > 1 else float(v)

    def dims_of(d):
        x, y, z = (max(64, int(round(v * args.scale / 32)) * 32) for v in d)
        return x, y, (z // world) * world

    def run_block(ss, first, n):
        barrier()
        ss.solver.enqueue_steps(first, n)
        ss.solver.sync()
        barrier()
        return maxr(ss.solver.last_timing()[0])

    for cfg in args.config:
        if cfg == "c3":
            dims, n_mat, ut, dtype, npdt = dims_of((1536, 1024, 960)), 5, 3, capi.F32, np.float32
            gen = lambda a, b: synth.hall(dims, n_mat, a, b)
            X, Y, Z = dims
            rec = [[int(X * (0.2 + 0.08 * i)), int(Y * (0.5 + 0.11 * j)), int(Z * (0.08 + 0.11 * i))] for i in range(8) for j in range(2)]
            sources = [[X // 2, int(Y * 0.25), Z // 2]]
            what = "IISO on the synthetic hall, 5 materials, 16 receivers on a seating grid"
        elif cfg == "c5":
            dims, n_mat, ut, dtype, npdt = dims_of((2048, 1024, 1920)), 20, 3, capi.F64, np.float64
            gen = lambda a, b: synth.banded_shoebox(dims, n_mat, a, b)
            X, Y, Z = dims
            rng = np.random.default_rng(5)
            reach = max(8, int(0.3 * args.steps))                    # the wave front travels ~0.87 voxels per step
            sources = [[int(X // 2 + rng.integers(-reach, reach)), int(Y // 2 + rng.integers(-reach, reach)), int(Z // 2 + rng.integers(-reach, reach))]
                       for _ in range(10)]
            rec = [[X // 2, Y // 2, Z // 2], [X // 5, int(Y * 0.8), int(Z * 0.9)]]     # one every run reaches, one far away
            what = "fp64 IISO, 20 materials in z-bands, 10 source positions as separate runs"
        else:
            raise SystemExit(f"unknown config {cfg}")
        steps = 4000 if (cfg == "c5" and args.full) else (args.c3_steps if cfg == "c3" else args.steps)
        lam = float(np.sqrt(3.0) / 2)
        prm = np.array([lam, lam * lam, 1.0 / 3.0, 0.0], dtype=npdt)
        tab = synth.material_table(list(np.linspace(0.99, 0.5, n_mat))).astype(npdt)
        t0 = time.time()
        res = {}
        for transport in (("auto", "nccl") if (cfg == "c3" and world > 1) else ("auto",)):
            ss = slabs.SlabSolver(capi, dims, gen, block=(32, 4, 1), element_type=ut, dtype=dtype, params=prm, materials=tab, rank=rank,
                                  world=world, device=lr,
                                  options=[(capi.OPT_MATIDX_AS_WRITTEN, 0), (capi.OPT_PEER_STORES, 0 if transport == "nccl" else 1)])
            ss.connect()
            t_setup = maxr(time.time() - t0)
            n = np.arange(steps, dtype=np.float64)
            pulse = np.exp(-0.5 * ((n - 40.0) / 6.0) ** 2).astype(npdt)[None, :]
            ss.set_receivers(rec)
            runs, resp_all = [], []
            for si, src in enumerate(sources if transport == "auto" else sources[:1]):
                if si:
                    ss.solver.reset_pressures()
                ss.set_sources([src], [capi.SRC_HARD], pulse)
                ss.solver.reserve_steps(steps)
                warm = min(20, steps // 4)
                run_block(ss, 0, warm)
                ms = run_block(ss, warm, steps - warm)
                runs.append(ms / (steps - warm))
                resp_all.append(ss.responses(steps))
            Xp, Yp, _ = ss.solver.dims()
            res[transport] = dict(ms_per_step=float(np.median(runs)), ms_per_step_runs=runs, kernel=ss.solver.kernel_name(),
                                  halo=ss.solver.halo_transport(), setup_seconds=t_setup, responses=resp_all, padded=(Xp, Yp))
            ss.close()
            t0 = time.time()
        if rank == 0:
            a = res["auto"]
            Xp, Yp = a["padded"]
            nvox = Xp * Yp * dims[2]
            r = a["responses"]
            line = {"config": cfg, "what": what, "dims": list(dims), "voxels": nvox, "n_gpus": world, "steps_per_run": steps, "runs": len(r),
                    "dtype": "f64" if dtype == capi.F64 else "f32", "ms_per_step": a["ms_per_step"], "ms_per_step_runs": a["ms_per_step_runs"],
                    "value": nvox / (a["ms_per_step"] * 1e-3) / 1e6, "unit": "Mvox/s", "kernel": a["kernel"], "halo": a["halo"],
                    "setup_seconds_incl_geometry": a["setup_seconds"],
                    "responses_finite": bool(all(np.isfinite(x).all() for x in r)),
                    "receivers_reached": [int((np.abs(x).max(axis=1) > 0).sum()) for x in r],
                    "distinct_responses": len({x.tobytes() for x in r})}
            if "nccl" in res:
                line["nccl_ms_per_step"] = res["nccl"]["ms_per_step"]
                line["same_responses_with_nccl_transport"] = bool(np.array_equal(res["nccl"]["responses"][0], r[0]))
            print(json.dumps(line), flush=True)
        barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
