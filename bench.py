#!/usr/bin/env python
"""bench.py -- Mvox-updates/s of the time-stepping hot path on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 (N=1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      the reference's own implementation of the path

A "step" is one leapfrog time step (source injection -> fused pressure/boundary update -> halo
exchange -> receiver capture) over the whole domain.  Workload at N GPUs (weak scaling): the
BASELINE.json config-2 room -- shoebox, 6 wall materials, SRL_FORWARD, fp32 -- of 512x512x512
voxels PER GPU, stacked in z (512 x 512 x 512N), one z-slab per rank exactly as the reference's
getPartitionIndexing would cut it.  Metric as the reference defines it (App.h:218): padded X*Y*Z
voxels (solid included) * steps / seconds / 1e6.

One JSON line on stdout (rank 0); see DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (per-GPU voxel dims x, y, z ; materials ; description)
    "c2": ((512, 512, 512), 6, "shoebox 512x512x512 per GPU, 6 wall materials (BASELINE config 2)"),
    "c4": ((1024, 1024, 960), 6, "shoebox 1024x1024x960 (1.007e9 voxels) per GPU (BASELINE config 4)"),
    "c1": ((64, 64, 64), 1, "shoebox 64x64x64 (BASELINE config 1)"),
    "c2half": ((512, 512, 256), 6, "shoebox 512x512x256 per GPU"),
}
ALGO_BYTES = {"f32": 13, "f64": 25}     # P^n + P^(n-1) + P^(n+1) + node byte per voxel update
UPDATE_NAMES = {0: "SRL_FORWARD", 1: "SHARED", 2: "SRL", 3: "IISO", 4: "IWB"}
COURANT = {0: float(np.sqrt(1.0 / 3.0)), 1: float(np.sqrt(1.0 / 3.0)), 2: float(np.sqrt(1.0 / 3.0)),   # setUpdateType, SimulationParameters.cpp:123-137
           3: float(np.sqrt(3.0) / 2), 4: 1.0}                                                          # IISO / IWB stability limits


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tma-hints", type=int, default=None, help="PFDTD_OPT_TMA_HINTS override (sweeps)")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"],
                    help="reference arm: the reference's own CUDA build (oracle/_ref) or the CPU oracle port")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--update-type", type=int, default=0, choices=[0, 1, 2, 3, 4])
    ap.add_argument("--dif-order", type=int, default=2, choices=[0, 1, 2, 3, 4],
                    help="frequency-dependent boundaries: order of the per-material digital impedance filters (BASELINE config 2 "
                         "asks for them; 0 = the reference's frequency-independent admittance, the only boundary the reference "
                         "arm can run)")
    ap.add_argument("--no-variants", action="store_true",
                    help="N=1 only: skip the short device-resident runs of the other BASELINE config-2 variants (DIF order 2, "
                         "fp64, IISO) that are appended to the JSON line as `variants`")
    ap.add_argument("--kernel", default="auto", choices=["auto", "tma", "plain"])
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=0, help="steps of the CPU-baseline sample (0 = sized for ~15 s)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append((time.time(), parts))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0 - 0.05) and (t1 is None or t <= t1 + 0.1)] or \
               [r for (_, r) in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, dtype, update_type, dif_order=0):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one matches."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        for e in json.load(open(p)):
            if (e.get("workload") == workload and e.get("dtype") == dtype and e.get("update_type") == update_type
                    and e.get("dif_order", 0) == dif_order):
                return e.get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None
    return None


def pinned_u8(shape):
    """uint8 host buffer in page-locked memory (torch is only the allocator here)."""
    try:
        import torch
        t = torch.empty(int(np.prod(shape)), dtype=torch.uint8, pin_memory=True)
        a = t.numpy().reshape(shape)
        a.flags.writeable = True
        return a, t
    except Exception:  # noqa: BLE001
        return np.empty(shape, dtype=np.uint8), None


def receiver_positions(gdims):
    """Four receivers: one 18 voxels from the source (reached within the warm-up at any domain height, so the
    finite-and-nonzero check of the responses means something at every N), three spread over the height so that
    tall multi-slab domains record in several slabs."""
    cx, cy, cz = gdims[0] // 2, gdims[1] // 2, gdims[2] // 2
    far = [[cx + 17, cy + 5, min(gdims[2] - 2, 3 + (i * (gdims[2] - 6)) // 3)] for i in (0, 2, 3)]
    return [[cx + 17, cy + 5, min(gdims[2] - 2, cz + 3)]] + far


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    from parallelfdtd_b200 import capi, synth, slabs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (one process per GPU)")
        args.gpus = world
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")

    (vx, vy, vz1), n_mat, wdesc = WORKLOADS[args.workload]
    gdims = (vx, vy, vz1 * world)
    K, W = args.steps, args.warmup
    double = args.dtype == "f64"
    dt = capi.F64 if double else capi.F32
    npdt = np.float64 if double else np.float32
    lam = COURANT[args.update_type]
    prm = np.array([lam, lam * lam, 1.0 / 3.0, 0.0], dtype=npdt)
    refl = list(np.linspace(0.99, 0.5, n_mat)) if n_mat > 1 else [0.9]
    tab = (synth.filter_material_table(refl, args.dif_order) if args.dif_order else synth.material_table(refl)).astype(npdt)
    plan = slabs.SlabPlan(gdims[2], world)
    z0, nz = plan.slab(rank)
    t_geo = time.time()
    bid_np, mat_np = synth.shoebox(gdims, n_mat, z0, z0 + nz)
    bid, _keep1 = pinned_u8(bid_np.shape)
    mat, _keep2 = pinned_u8(mat_np.shape)
    bid[...] = bid_np
    mat[...] = mat_np
    del bid_np, mat_np
    t_geo = time.time() - t_geo
    total = W + K + 64
    cx, cy, cz = gdims[0] // 2, gdims[1] // 2, gdims[2] // 2
    src_xyz = [[cx, cy, cz]]
    n = np.arange(total, dtype=np.float64)
    src_tab = np.exp(-0.5 * ((n - 40.0) / 6.0) ** 2).astype(npdt)[None, :]      # DATA-type input: a Gaussian pulse
    rec_xyz = receiver_positions(gdims)
    opts = [(capi.OPT_MATIDX_AS_WRITTEN, 0), (capi.OPT_OVERLAP, 0 if args.no_overlap else 1),
            (capi.OPT_KERNEL, {"auto": capi.KERNEL_AUTO, "tma": capi.KERNEL_TMA, "plain": capi.KERNEL_PLAIN}[args.kernel]),
            (capi.OPT_TMA_TILE, args.tile), (capi.OPT_TMA_CHUNK, args.chunk), (capi.OPT_DIF_ORDER, args.dif_order)]
    if args.tma_hints is not None:
        opts.append((capi.OPT_TMA_HINTS, args.tma_hints))

    def make_solver():
        ss = slabs.SlabSolver(capi, gdims, lambda a, b: (bid, mat), block=(32, 4, 1), element_type=args.update_type, dtype=dt,
                              params=prm, materials=tab, rank=rank, world=world, device=local_rank, options=opts)
        return ss

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident measurement (`value`) --------------------------------------------------------
    ss = make_solver()
    uid = ss.connect()
    s = ss.solver
    ss.set_sources(src_xyz, [capi.SRC_HARD], src_tab)
    ss.set_receivers(rec_xyz)
    s.reserve_steps(total)
    X, Y, _ = s.dims()
    nvox_global = X * Y * gdims[2]
    s.enqueue_steps(0, W)
    s.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    barrier()
    l0 = s.launch_count()
    t0 = time.time()
    s.enqueue_steps(W, K)
    s.sync()
    t1 = time.time()
    barrier()
    launches = s.launch_count() - l0
    dev_ms, _, _ = s.last_timing()
    halo_ms_last = s.last_halo_ms() if world > 1 else 0.0
    step_ms = slabs.max_over_ranks(dev_ms) if world > 1 else dev_ms
    wall_ms = slabs.max_over_ranks((t1 - t0) * 1e3) if world > 1 else (t1 - t0) * 1e3
    kname = s.kernel_name()

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream ------------
    s.set_option(capi.OPT_TIME_KERNELS, 1)
    K2 = max(3, min(K, 100))
    s.enqueue_steps(W + K, min(K2, total - W - K))
    s.sync()
    t2 = time.time()
    _, kern_ms, n_k = s.last_timing()
    s.set_option(capi.OPT_TIME_KERNELS, 0)
    clocks = sampler.stop(t0, t2) if rank == 0 else None
    resp = ss.responses(W + K)
    resp_ok = bool(np.isfinite(resp).all() and np.abs(resp).max() > 0)
    steps_timed_k = min(K2, total - W - K)
    slab_updates = X * Y * (nz - 2)                       # voxels one rank updates per step
    kern_ms_per_step = kern_ms / max(steps_timed_k, 1)      # all update launches of one step on this rank
    peak, peak_src = measured_peak()
    achieved = slab_updates * ALGO_BYTES[args.dtype] / (kern_ms_per_step * 1e-3) / 1e9 if kern_ms_per_step > 0 else 0.0
    traffic = ncu_traffic(args.workload, args.dtype, args.update_type, args.dif_order)
    dif_addon = None
    if args.dif_order:   # reported next to, not inside, the 13 B / 25 B per voxel update
        _, _, n_bnd = s.counts()
        pad = 1 if args.dif_order == 1 else (2 if args.dif_order == 2 else 4)
        dif_addon = {"filter_voxels": int(n_bnd), "state_bytes_read_plus_written": int(2 * n_bnd * pad * prm.itemsize),
                     "row_segment_entries_bytes": int(8 * nz * Y * ((X + 127) // 128))}
    ss.close()

    # ---- end-to-end through the C ABI with HOST buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        barrier()
        te0 = time.time()
        se = make_solver()                                   # H2D of both node volumes, pad, translate, partition, alloc
        if world > 1:
            se.connect(uid)                                   # communicator of this job: created once per process, reused
        se.set_sources(src_xyz, [capi.SRC_HARD], src_tab[:, :K])     # H2D of the source table happens in run()
        se.set_receivers(rec_xyz)
        r_e2e, _ = se.solver.run(K)                          # K steps + D2H of the responses
        te1 = time.time()
        se.close()
        t_e2e = slabs.max_over_ranks(te1 - te0) if world > 1 else (te1 - te0)
        h2d = int(bid.size + mat.size + tab.nbytes + prm.nbytes + src_tab[:, :K].nbytes + 12 * (len(src_xyz) + len(rec_xyz)))
        d2h = int(len(rec_xyz) * K * prm.itemsize + 16)
        e2e = {"value": nvox_global * K / t_e2e / 1e6, "unit": "Mvox/s", "h2d_bytes_per_step": h2d * world / K,
               "d2h_bytes_per_step": d2h / K, "seconds": t_e2e,
               "what": "pfdtd_setup_mesh(host bid+mat, pinned) + make_partition + set_sources/receivers + pfdtd_run(K) incl. response D2H"
                       + ("; the job's NCCL communicator already exists (created once per process)" if world > 1 else "")}

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, bid, mat, tab, prm, src_xyz, rec_xyz, src_tab)

    if rank == 0:
        value = nvox_global * K / (step_ms * 1e-3) / 1e6
        line = {
            "metric": "Mvox-updates/s", "value": value, "unit": "Mvox/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{wdesc}; global {gdims[0]}x{gdims[1]}x{gdims[2]} -> padded {X}x{Y}x{gdims[2]}",
                       "update_type": UPDATE_NAMES[args.update_type], "materials": n_mat, "sources": 1, "receivers": len(rec_xyz),
                       "boundaries": (f"frequency-dependent: order-{args.dif_order} digital impedance filter per material, states of the "
                                      "boundary voxels updated in the same kernel pass") if args.dif_order else
                                     "frequency-independent admittance per material (the reference's boundary)",
                       "slabs": world, "halo": "none" if world == 1 else "one plane each way per interface per step, NCCL p2p over NVLink"
                       + ("" if args.no_overlap else ", overlapped with the interior update"),
                       "kernel": kname, "cache": "inputs larger than L2 (fields %.0f MiB per GPU vs 126 MB L2), no flush" %
                       (2 * X * Y * nz * prm.itemsize / 2**20), "wall_ms_per_step": wall_ms / K,
                       "responses_finite_nonzero": resp_ok, "geometry_seconds": round(t_geo, 2)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_voxel_update": ALGO_BYTES[args.dtype], "voxel_updates_per_step_per_gpu": slab_updates,
                         "kernel_ms_per_step": kern_ms_per_step, "update_launches_per_step": n_k / max(steps_timed_k, 1),
                         "how": f"CUDA events around every update launch over {steps_timed_k} steps right after the timed region",
                         "addon_bytes_per_step_not_in_achieved": dif_addon},
            "cpu_baseline": cpu,
        }
        if world == 1 and not args.no_variants:
            line["variants"] = run_variants(args)
        if world > 1:
            line["halo_ms_last_step"] = halo_ms_last
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


VARIANTS = [("f32 SRL_FORWARD, frequency-independent admittance (the reference's boundary; parity pinned bit-exact)", "f32", 0, 0),
            ("f32 SRL_FORWARD + DIF order 2", "f32", 0, 2),
            ("f64 SRL_FORWARD + DIF order 2", "f64", 0, 2),
            ("f64 SRL_FORWARD, frequency-independent", "f64", 0, 0),
            ("f32 SRL (centred boundary) + DIF order 2", "f32", 2, 2),
            ("f32 IISO (27-point), frequency-independent", "f32", 3, 0),
            ("f32 IISO + DIF order 2", "f32", 3, 2),
            ("f64 IISO + DIF order 2", "f64", 3, 2)]


def run_variants(args):
    """The other variants BASELINE config 2 names (fp64, the reference's frequency-independent boundary) and the
    interpolated scheme, each as a short device-resident run of this same script in a child process (same workload,
    200 steps)."""
    out = []
    for name, dtype, ut, order in VARIANTS:
        if (dtype, ut, order) == (args.dtype, args.update_type, args.dif_order):
            continue
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--steps", "200", "--warmup", "10", "--no-e2e",
               "--no-cpu-baseline", "--no-variants", "--dtype", dtype, "--update-type", str(ut), "--dif-order", str(order)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out.append({"variant": name, "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                        "kernel": d["config"]["kernel"], "roofline_achieved_gbs": d["roofline"]["achieved"],
                        "roofline_frac": d["roofline"]["frac"], "bytes_per_voxel_update": d["roofline"]["bytes_per_voxel_update"]})
        except Exception as e:  # noqa: BLE001
            out.append({"variant": name, "error": repr(e)[:200]})
    return out


def cpu_baseline(args, bid, mat, tab, prm, src_xyz, rec_xyz, src_tab):
    """Oracle port (oracle/fdtd_oracle.cpp, OpenMP) on the same workload for a bounded number of steps."""
    from oracle import oracle
    double = args.dtype == "f64"
    t0 = time.time()
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (32, 4, 1), args.update_type, double)
    nvox = pos.size
    threads = oracle.num_threads()
    scheme = 0 if args.update_type in (0, 1) else (2 if args.update_type == 2 else 3)
    if scheme == 3:
        prm = oracle.params_interp(float(prm[0]), 0, oracle.interp_coefficients(args.update_type, float(prm[1])), double)
    if args.dif_order:
        def timed(k):
            t = time.time()
            oracle.run_dif(pos, m, scheme, prm, tab, args.dif_order, src_xyz, [0], np.ascontiguousarray(src_tab[:, :k]), rec_xyz, k, 1)
            return time.time() - t
        steps = args.cpu_steps
        if not steps:   # two short runs separate the per-run setup from the per-step cost; then ~15 s of steps
            t3, t9 = timed(3), timed(9)
            per = max((t9 - t3) / 6, 1e-6)
            steps = int(max(8, min(src_tab.shape[1], 15.0 / per)))
        secs = timed(steps)
        return {"value": nvox * steps / secs / 1e6, "unit": "Mvox/s", "cores": threads, "kind": "port",
                "sample": f"same workload (order-{args.dif_order} filters), {steps} steps of the C++/OpenMP oracle incl. its per-run setup "
                          f"({secs:.1f} s; {time.time() - t0:.1f} s with calibration)"}
    steps = args.cpu_steps
    if not steps:   # calibrate on 4 steps, then size the sample for ~12 s of CPU work
        _, s4 = oracle.run(pos, m, scheme, prm, tab, src_xyz, [0], np.ascontiguousarray(src_tab[:, :5]), rec_xyz, 5, 1, 0, 0, 1)
        steps = int(max(8, min(src_tab.shape[1], 12.0 / max(s4 / 4, 1e-6))))
    smp = np.ascontiguousarray(src_tab[:, :steps])
    _, secs = oracle.run(pos, m, scheme, prm, tab, src_xyz, [0], smp, rec_xyz, steps, 1, 0, 0, 1)
    timed = steps - 1
    return {"value": nvox * timed / secs / 1e6, "unit": "Mvox/s", "cores": threads, "kind": "port",
            "sample": f"same workload, {timed} timed steps (1 warm-up) of the C++/OpenMP oracle, {time.time() - t0:.1f} s total"}


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own implementation of the path: its CUDA kernels + CudaMesh + launchFDTD3d, built
    unmodified into oracle/_ref/ref_fdtd (the reference has no CPU implementation; its N-GPU mode is one
    process driving N devices).  --ref-kind port times the CPU oracle port instead."""
    from oracle import casefile, oracle
    from parallelfdtd_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n = max(args.gpus, world)
    (vx, vy, vz1), n_mat, wdesc = WORKLOADS[args.workload]
    gdims = (vx, vy, vz1 * n)
    K, W = args.steps, args.warmup
    double = args.dtype == "f64"
    kind = args.ref_kind
    if kind == "auto":
        kind = "reference" if casefile.ref_available() else "port"
    nvox_in = gdims[0] * gdims[1] * gdims[2]
    base = {"impl": "reference", "metric": "Mvox-updates/s", "unit": "Mvox/s", "n_gpus": n, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic"}
    if kind == "reference" and nvox_in >= 2**31:
        # the reference indexes voxels with 32-bit ints (cudaMesh.h:142, kernels3d.cu:505-506)
        emit({"impl": "reference", "unavailable": f"{nvox_in} voxels exceed the reference's 32-bit indexing"})
        return
    bid, mat = synth.shoebox(gdims, n_mat)
    tab = synth.material_table(list(np.linspace(0.99, 0.5, n_mat)) if n_mat > 1 else [0.9])
    cx, cy, cz = gdims[0] // 2, gdims[1] // 2, gdims[2] // 2
    rec_xyz = [tuple(r) for r in receiver_positions(gdims)]
    nn = np.arange(K + W + 64, dtype=np.float64)
    pulse = np.exp(-0.5 * ((nn - 40.0) / 6.0) ** 2)
    if kind == "reference":
        case = dict(bid=bid, mat=mat, block=(32, 4, 1), update_type=args.update_type, double=double, steps=K, octave=0,
                    n_parts=n, devices=list(range(n)), materials=tab, sources=[(cx, cy, cz, 0, 3, 0)], receivers=rec_xyz,
                    input_data=[pulse])
        res = casefile.run_reference(case, "/tmp/pfdtd_ref/bench", warmup_steps=W, timeout=3000)
        X, Y, Z = res["dims"]
        nvox = X * Y * Z
        wall, wall_e2e = res["wall_seconds"], res["wall_e2e_seconds"]
        value = nvox * K / wall / 1e6
        line = dict(base, value=value, ms_per_step=wall / K * 1e3,
                    config={"workload": f"{wdesc}; global {gdims[0]}x{gdims[1]}x{gdims[2]} -> padded {X}x{Y}x{Z}",
                            "update_type": UPDATE_NAMES[args.update_type], "materials": n_mat, "slabs": n,
                            "boundaries": "frequency-independent admittance per material: the reference contains no digital impedance "
                                          "filter (SURVEY section 0), so its arm runs the boundary it has on the same room",
                            "what": "reference src/kernels/{kernels3d,cudaMesh,cudaUtils}.cu + host classes compiled unmodified for sm_100 "
                                    "(oracle/Makefile), driven like its own tests: setupMesh -> makePartition -> launchFDTD3d"},
                    cpu_baseline={"value": value, "unit": "Mvox/s", "cores": 1, "kind": "reference",
                                  "sample": f"launchFDTD3d{'Double' if double else ''} for {K} steps after a {W}-step warm-up run; "
                                            "1 host thread + the box's B200(s): the reference's hot path is itself CUDA"},
                    e2e={"value": nvox * K / wall_e2e / 1e6, "unit": "Mvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                         "seconds": wall_e2e,
                         "what": "toDevice(bid,mat) + setupMesh + makePartition + launchFDTD3d incl. per-step source H2D and final response D2H"},
                    gpu_launches=0)
        emit(line)
        return
    # CPU oracle port
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (32, 4, 1), args.update_type, double)
    npdt = np.float64 if double else np.float32
    lam = float(np.sqrt(1.0 / 3.0))
    prm = np.array([lam, lam * lam, 1.0 / 3.0, 0.0], dtype=npdt)
    threads = oracle.num_threads()
    steps = int(max(4, min(K, 60.0 * 0.6e9 * max(threads, 1) / 16 / pos.size)))
    smp = pulse[None, :steps].astype(npdt)
    _, secs = oracle.run(pos, m, 0 if args.update_type in (0, 1) else 2, prm, tab.astype(npdt), [(cx, cy, cz)], [0], smp, rec_xyz,
                         steps, n, 0, 0, 1)
    value = pos.size * (steps - 1) / secs / 1e6
    line = dict(base, value=value, ms_per_step=secs / (steps - 1) * 1e3, steps=steps - 1,
                config={"workload": f"{wdesc}; global {gdims[0]}x{gdims[1]}x{gdims[2]}", "update_type": UPDATE_NAMES[args.update_type],
                        "materials": n_mat, "slabs": n},
                cpu_baseline={"value": value, "unit": "Mvox/s", "cores": threads, "kind": "port",
                              "sample": f"{steps - 1} timed steps of the C++/OpenMP oracle port on the same workload"},
                e2e={"value": value, "unit": "Mvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
    emit(line)


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: everything libraries print there while the job runs (NCCL's
    version banner, torch warnings) is sent to stderr instead, and the result goes to the original descriptor."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    txt = json.dumps(line) + "\n"
    if _RESULT_FD is None:
        sys.stdout.write(txt)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, txt.encode())


if __name__ == "__main__":
    a = parse_args()
    claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
