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
    ap.add_argument("--halo", default="auto", choices=["auto", "nccl"],
                    help="N>1 halo transport: auto = peer-mapped stores from the edge launches (CUDA IPC) where possible; nccl = ncclSend/Recv")
    ap.add_argument("--no-like-for-like", action="store_true", help="skip the frequency-independent block the reference arm is compared with")
    ap.add_argument("--no-invariance", action="store_true", help="N>1: skip the single-GPU re-run that checks the responses bit for bit")
    ap.add_argument("--c4", default="auto", choices=["auto", "on", "off"],
                    help="append the BASELINE config-4 block (1024x1024x960 voxels per GPU); auto = with the default workload")
    ap.add_argument("--c4-steps", type=int, default=100)
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
REPEATS = 5          # the K-step block is timed this many times; the median block is the reported one
E2E_REPEATS = 5


class Ctx:
    """what every measurement of one bench invocation shares: ranks, the room of this rank's slab in pinned host memory"""

    def __init__(self, args, workload):
        from parallelfdtd_b200 import capi, slabs, synth
        self.capi, self.slabs, self.synth = capi, slabs, synth
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.workload = workload
        (vx, vy, vz1), self.n_mat, self.wdesc = WORKLOADS[workload]
        self.gdims = (vx, vy, vz1 * self.world)
        self.plan = slabs.SlabPlan(self.gdims[2], self.world)
        self.z0, self.nz = self.plan.slab(self.rank)
        t = time.time()
        bid_np, mat_np = synth.shoebox(self.gdims, self.n_mat, self.z0, self.z0 + self.nz)
        self.bid, self._k1 = pinned_u8(bid_np.shape)
        self.mat, self._k2 = pinned_u8(mat_np.shape)
        self.bid[...] = bid_np
        self.mat[...] = mat_np
        self.t_geo = time.time() - t
        self.uid = None
        cx, cy, cz = self.gdims[0] // 2, self.gdims[1] // 2, self.gdims[2] // 2
        self.src_xyz = [[cx, cy, cz]]
        self.rec_xyz = receiver_positions(self.gdims)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def maxr(self, v):
        return self.slabs.max_over_ranks(v) if self.world > 1 else float(v)

    def tables(self, dtype, update_type, dif_order):
        npdt = np.float64 if dtype == "f64" else np.float32
        lam = COURANT[update_type]
        prm = np.array([lam, lam * lam, 1.0 / 3.0, 0.0], dtype=npdt)
        refl = list(np.linspace(0.99, 0.5, self.n_mat)) if self.n_mat > 1 else [0.9]
        tab = (self.synth.filter_material_table(refl, dif_order) if dif_order else self.synth.material_table(refl)).astype(npdt)
        return prm, tab

    def options(self, dif_order, overlap=True):
        capi, a = self.capi, self.args
        opts = [(capi.OPT_MATIDX_AS_WRITTEN, 0), (capi.OPT_OVERLAP, 1 if overlap else 0),
                (capi.OPT_KERNEL, {"auto": capi.KERNEL_AUTO, "tma": capi.KERNEL_TMA, "plain": capi.KERNEL_PLAIN}[a.kernel]),
                (capi.OPT_TMA_TILE, a.tile), (capi.OPT_TMA_CHUNK, a.chunk), (capi.OPT_DIF_ORDER, dif_order),
                (capi.OPT_PEER_STORES, 0 if a.halo == "nccl" else 1)]
        if a.tma_hints is not None:
            opts.append((capi.OPT_TMA_HINTS, a.tma_hints))
        return opts

    def solver(self, dtype, update_type, dif_order, overlap=True):
        capi = self.capi
        prm, tab = self.tables(dtype, update_type, dif_order)
        return self.slabs.SlabSolver(capi, self.gdims, lambda a, b: (self.bid, self.mat), block=(32, 4, 1), element_type=update_type,
                                     dtype=capi.F64 if dtype == "f64" else capi.F32, params=prm, materials=tab, rank=self.rank,
                                     world=self.world, device=self.local_rank, options=self.options(dif_order, overlap))

    def connect(self, ss):
        if self.world > 1:
            self.uid = ss.connect(self.uid)        # the job's communicator: made once per process, reused afterwards


def pulse_table(total, npdt):
    n = np.arange(total, dtype=np.float64)
    return np.exp(-0.5 * ((n - 40.0) / 6.0) ** 2).astype(npdt)[None, :]      # DATA-type input: a Gaussian pulse


def measure(ctx, dtype, update_type, dif_order, K, W, with_e2e=True, sampler=None):
    """Device-resident K-step blocks (REPEATS of them, median reported), the roofline of the dominant kernel from per-launch
    CUDA events, and the end-to-end run through the C ABI from host buffers.  Returns a dict; collective over all ranks."""
    capi = ctx.capi
    npdt = np.float64 if dtype == "f64" else np.float32
    K2 = max(3, min(K, 100))
    total = W + REPEATS * K + K2
    src_tab = pulse_table(total, npdt)
    ss = ctx.solver(dtype, update_type, dif_order, overlap=not ctx.args.no_overlap)
    ctx.connect(ss)
    s = ss.solver
    ss.set_sources(ctx.src_xyz, [capi.SRC_HARD], src_tab)
    ss.set_receivers(ctx.rec_xyz)
    s.reserve_steps(total)
    X, Y, _ = s.dims()
    nvox_global = X * Y * ctx.gdims[2]
    s.enqueue_steps(0, W)
    s.sync()
    ctx.barrier()
    if sampler is not None and ctx.rank == 0:
        sampler.start()
        time.sleep(0.15)
    blocks, walls = [], []
    l0 = s.launch_count()
    t_first = time.time()
    for r in range(REPEATS):
        ctx.barrier()                                    # barrier + synchronize on both sides of every K-step block
        t0 = time.time()
        s.enqueue_steps(W + r * K, K)
        s.sync()
        t1 = time.time()
        ctx.barrier()
        blocks.append(ctx.maxr(s.last_timing()[0]))       # device time (CUDA events on the launching stream), max over ranks
        walls.append(ctx.maxr((t1 - t0) * 1e3))
    launches = (s.launch_count() - l0) // REPEATS
    med = int(np.argsort(blocks)[len(blocks) // 2])
    step_ms = blocks[med]
    # ---- roofline of the dominant kernel: CUDA events around every update launch, on the launch's own stream ----
    s.set_option(capi.OPT_TIME_KERNELS, 1)
    s.enqueue_steps(W + REPEATS * K, K2)
    s.sync()
    t_last = time.time()
    bulk_ms, n_bulk, bulk_planes, edge_ms, n_edge = s.last_timing_detail()
    s.set_option(capi.OPT_TIME_KERNELS, 0)
    halo_clean_ms = s.time_halo_exchange(20) if ctx.world > 1 else 0.0
    transport = s.halo_transport()
    resp = ss.responses(total)
    kname = s.kernel_name()
    peak, peak_src = measured_peak()
    bulk_bytes = X * Y * bulk_planes * ALGO_BYTES[dtype]                      # algorithmic bytes of all bulk launches timed
    achieved = bulk_bytes / (bulk_ms * 1e-3) / 1e9 if bulk_ms > 0 else 0.0
    dif_addon = None
    if dif_order:   # reported next to, not inside, the 13 B / 25 B per voxel update
        _, _, n_bnd = s.counts()
        pad = 1 if dif_order == 1 else (2 if dif_order == 2 else 4)
        dif_addon = {"filter_voxels": int(n_bnd), "state_bytes_read_plus_written": int(2 * n_bnd * pad * np.dtype(npdt).itemsize),
                     "row_segment_entries_bytes": int(8 * ctx.nz * Y * ((X + 127) // 128))}
    ss.close()
    clocks = sampler.stop(t_first, t_last) if (sampler is not None and ctx.rank == 0) else None   # the timed region only
    out = {
        "clocks": clocks,
        "value": nvox_global * K / (step_ms * 1e-3) / 1e6, "ms_per_step": step_ms / K, "step_ms_blocks": [b / K for b in blocks],
        "wall_ms_per_step": walls[med] / K, "launches": int(launches), "kernel": kname, "halo": transport,
        "padded": (X, Y), "nvox_global": nvox_global, "responses": resp, "t_window": (t_first, t_last),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(ctx.workload, dtype, update_type, dif_order), "peak_source": peak_src,
                     "bytes_per_voxel_update": ALGO_BYTES[dtype], "voxel_updates_per_launch": X * Y * bulk_planes / max(n_bulk, 1),
                     "kernel_ms_per_launch": bulk_ms / max(n_bulk, 1), "launches_timed": int(n_bulk),
                     "what": "the full-slab launch (N=1) / the interior launch of a slab with neighbours (N>1); the one-plane edge "
                             "launches run concurrently on their own stream and are listed separately",
                     "edge_kernel_ms_per_step": (edge_ms / K2) if n_edge else 0.0, "edge_launches_per_step": n_edge / K2,
                     "how": f"CUDA events around every update launch over {K2} steps right after the timed blocks (per-step launches: the "
                            "dispatch latency of a launch falls inside its bracket; the timed blocks replay a CUDA graph)",
                     # N = 1: one bulk launch per step, so the graph-replayed step time (which also holds the source / receiver
                     # work) bounds the kernel's bandwidth from below without any event in between
                     "achieved_from_step_time": (X * Y * (ctx.nz - 2) * ALGO_BYTES[dtype] / (step_ms / K * 1e-3) / 1e9) if ctx.world == 1 else None,
                     "frac_from_step_time": (X * Y * (ctx.nz - 2) * ALGO_BYTES[dtype] / (step_ms / K * 1e-3) / 1e9 / peak) if ctx.world == 1 else None,
                     "addon_bytes_per_step_not_in_achieved": dif_addon},
        "halo_ms_per_exchange_alone": halo_clean_ms,
    }
    if with_e2e:
        out["e2e"] = measure_e2e(ctx, dtype, update_type, dif_order, K, src_tab, nvox_global)
    return out


def measure_e2e(ctx, dtype, update_type, dif_order, K, src_tab, nvox_global):
    """The same job through the reference-facing calls with HOST buffers: setup_mesh (H2D of both node volumes from pinned
    memory, pad + translate + classes), make_partition, sources / receivers, pfdtd_run(K) including the D2H of the responses."""
    capi = ctx.capi
    prm, tab = ctx.tables(dtype, update_type, dif_order)
    runs = []
    for _ in range(E2E_REPEATS):
        ctx.barrier()
        t0 = time.time()
        se = ctx.solver(dtype, update_type, dif_order, overlap=not ctx.args.no_overlap)     # setup_mesh + make_partition
        t1 = time.time()
        ctx.connect(se)
        t2 = time.time()
        se.set_sources(ctx.src_xyz, [capi.SRC_HARD], src_tab[:, :K])
        se.set_receivers(ctx.rec_xyz)
        t3 = time.time()
        se.solver.run(K)                                                                    # K steps + D2H of the responses
        t4 = time.time()
        ctx.barrier()
        runs.append({"seconds": ctx.maxr(t4 - t0), "setup_seconds": ctx.maxr(t1 - t0), "connect_seconds": ctx.maxr(t2 - t1),
                     "srcrec_seconds": ctx.maxr(t3 - t2), "run_seconds": ctx.maxr(t4 - t3)})
        se.close()
    runs.sort(key=lambda r: r["seconds"])
    mid = runs[len(runs) // 2]
    h2d = int(ctx.bid.size + ctx.mat.size + tab.nbytes + prm.nbytes + src_tab[:, :K].nbytes + 12 * (len(ctx.src_xyz) + len(ctx.rec_xyz)))
    d2h = int(len(ctx.rec_xyz) * K * prm.itemsize + 16)
    return {"value": nvox_global * K / mid["seconds"] / 1e6, "unit": "Mvox/s", "h2d_bytes_per_step": h2d * ctx.world / K,
            "d2h_bytes_per_step": d2h / K, "seconds": mid["seconds"], "setup_seconds": mid["setup_seconds"],
            "run_seconds": mid["run_seconds"], "srcrec_seconds": mid["srcrec_seconds"], "connect_seconds": mid["connect_seconds"],
            "all_seconds": [r["seconds"] for r in runs], "all_setup_seconds": [r["setup_seconds"] for r in runs], "repeats": E2E_REPEATS,
            "what": "median of %d: pfdtd_setup_mesh(host bid+mat, pinned) + make_partition [setup_seconds] + set_sources/receivers "
                    "+ pfdtd_run(K) incl. response D2H [run_seconds]; a repeat builds its solver out of the device blocks the library "
                    "kept from the solvers destroyed before it (pfdtd_release_cached_memory), as every job after the first of a "
                    "session does" % E2E_REPEATS
                    + ("; the job's NCCL communicator already exists (created once per process)" if ctx.world > 1 else "")}


def single_gpu_responses(ctx, dtype, update_type, dif_order, n_steps, src_tab):
    """rank 0 only: the SAME global domain as one slab on one GPU (the reference's invariant, CudaMeshTest.cpp:472-575)"""
    capi = ctx.capi
    bid, mat = ctx.synth.shoebox(ctx.gdims, ctx.n_mat)
    prm, tab = ctx.tables(dtype, update_type, dif_order)
    s = capi.Solver()
    try:
        for k, v in ctx.options(dif_order):
            s.set_option(k, v)
        s.setup_mesh(bid, mat, (32, 4, 1), update_type, capi.F64 if dtype == "f64" else capi.F32, prm, tab)
        del bid, mat
        s.make_partition(1, [ctx.local_rank])
        s.set_sources(ctx.src_xyz, [capi.SRC_HARD], src_tab[:, :n_steps])
        s.set_receivers(ctx.rec_xyz)
        r, _ = s.run(n_steps)
    finally:
        s.close()
    return r


def slab_invariance(ctx, dtype, update_type, dif_order, multi_resp):
    """N>1: rank 0 re-runs the whole domain on ONE GPU for the same number of steps; the receiver responses of the N-slab
    run must equal it bit for bit.  Returns (ok, detail) on rank 0, (None, None) elsewhere; all ranks wait."""
    res = (None, None)
    if ctx.rank == 0:
        n_steps = multi_resp.shape[1]
        npdt = np.float64 if dtype == "f64" else np.float32
        t = time.time()
        one = single_gpu_responses(ctx, dtype, update_type, dif_order, n_steps, pulse_table(n_steps, npdt))
        ok = bool(np.array_equal(one, multi_resp) and np.abs(one).max() > 0)
        res = (ok, {"steps": int(n_steps), "receivers": int(one.shape[0]), "max_abs_diff": float(np.abs(one.astype(np.float64) - multi_resp).max()),
                    "seconds": round(time.time() - t, 1)})
    ctx.barrier()
    return res


def c4_block(args, base_ctx):
    """BASELINE config 4 at its stated size: 1024 x 1024 x 960 voxels (1.007e9) PER GPU, SRL fp32, the reference's
    frequency-independent boundary; step time with the halo exchange overlapped with the interior update and not, the
    halo exchange alone, and (N <= 2) the single-GPU invariance check."""
    sub = argparse.Namespace(**vars(args))
    ctx = Ctx(sub, "c4")
    ctx.dist, ctx.uid = base_ctx.dist, None
    capi = ctx.capi
    K, W = args.c4_steps, 10
    out = {"workload": f"{ctx.wdesc}; global {ctx.gdims[0]}x{ctx.gdims[1]}x{ctx.gdims[2]} = {ctx.gdims[0] * ctx.gdims[1] * ctx.gdims[2] / 1e9:.2f}e9 voxels",
           "update_type": "SRL_FORWARD", "dtype": "f32", "steps": K, "warmup": W, "geometry_seconds": round(ctx.t_geo, 1)}
    for label, no_overlap in (("overlap", False), ("no_overlap", True)):
        if ctx.world == 1 and no_overlap:
            continue
        sub.no_overlap = no_overlap
        m = measure(ctx, "f32", 0, 0, K, W, with_e2e=False)
        out[label] = {"value": m["value"], "unit": "Mvox/s", "ms_per_step": m["ms_per_step"], "step_ms_blocks": m["step_ms_blocks"],
                      "halo": m["halo"], "kernel": m["kernel"], "roofline_frac": m["roofline"]["frac"],
                      "roofline_achieved_gbs": m["roofline"]["achieved"], "interior_kernel_ms": m["roofline"]["kernel_ms_per_launch"],
                      "edge_kernel_ms_per_step": m["roofline"]["edge_kernel_ms_per_step"],
                      "halo_ms_per_exchange_alone": m["halo_ms_per_exchange_alone"]}
        if label == "overlap":
            resp = m["responses"]
            out["responses_finite_nonzero"] = bool(np.isfinite(resp).all() and np.abs(resp).max() > 0)
            if 1 < ctx.world <= 2:
                ok, det = slab_invariance(ctx, "f32", 0, 0, resp)
                if ctx.rank == 0:
                    out["slab_invariance"], out["slab_invariance_detail"] = ok, det
    if "no_overlap" in out and "overlap" in out:
        out["overlap_gain"] = out["no_overlap"]["ms_per_step"] / out["overlap"]["ms_per_step"]
    return out


def run_ours(args):
    from parallelfdtd_b200 import capi

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
    ctx = Ctx(args, args.workload)
    ctx.dist = dist
    K, W = args.steps, args.warmup
    gdims, n_mat, t_geo = ctx.gdims, ctx.n_mat, ctx.t_geo

    sampler = ClockSampler(local_rank)
    head = measure(ctx, args.dtype, args.update_type, args.dif_order, K, W, with_e2e=not args.no_e2e, sampler=sampler)
    clocks = head["clocks"]
    resp = head["responses"]
    resp_ok = bool(np.isfinite(resp).all() and np.abs(resp).max() > 0)

    # ---- like for like with the reference arm: the frequency-independent boundary, the only one the reference has ----
    l4l = None
    if args.dif_order != 0 and not args.no_like_for_like:
        l4l = measure(ctx, args.dtype, args.update_type, 0, K, W, with_e2e=not args.no_e2e)

    # ---- N > 1: the N-slab responses against the same domain on one GPU, bit for bit ----
    inv = None
    if world > 1 and not args.no_invariance:
        ok, det = slab_invariance(ctx, args.dtype, args.update_type, args.dif_order, resp)
        inv = {"ok": ok, "headline": det}
        if l4l is not None:
            ok2, det2 = slab_invariance(ctx, args.dtype, args.update_type, 0, l4l["responses"])
            if rank == 0:
                inv["ok"] = bool(ok and ok2)
                inv["like_for_like"] = det2

    c4 = None
    if args.c4 == "on" or (args.c4 == "auto" and args.workload == "c2" and (args.dtype, args.update_type) == ("f32", 0)):
        del ctx.bid, ctx.mat
        c4 = c4_block(args, ctx)
        ctx = None

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c2 = Ctx(args, args.workload)
        prm, tab = c2.tables(args.dtype, args.update_type, args.dif_order)
        cpu = cpu_baseline(args, c2.bid, c2.mat, tab, prm, c2.src_xyz, c2.rec_xyz, pulse_table(W + K + 64, prm.dtype.type))
        del c2

    if rank == 0:
        X, Y = head["padded"]
        bdesc = {True: (f"frequency-dependent: order-{args.dif_order} digital impedance filter per material, states of the boundary "
                        "voxels updated in the same kernel pass"),
                 False: "frequency-independent admittance per material (the reference's boundary)"}
        halo_desc = "none" if world == 1 else ("one plane each way per interface per step over NVLink: " + head["halo"]
                                               + ("" if args.no_overlap else ", overlapped with the interior update"))
        def cfg(m, dif):
            return {"workload": f"{WORKLOADS[args.workload][2]}; global {gdims[0]}x{gdims[1]}x{gdims[2]} -> padded {X}x{Y}x{gdims[2]}",
                    "update_type": UPDATE_NAMES[args.update_type], "materials": n_mat, "slabs": world, "boundaries": bdesc[bool(dif)],
                    "sources": 1, "receivers": 4, "halo": halo_desc, "kernel": m["kernel"],
                    "cache": "inputs larger than L2 (fields %.0f MiB per GPU vs 126 MB L2), no flush" %
                             (2 * X * Y * WORKLOADS[args.workload][0][2] * (8 if args.dtype == "f64" else 4) / 2**20),
                    "timing": f"{REPEATS} blocks of K steps, each bracketed by barrier + synchronize, device events, max over ranks; "
                              "the median block is reported, all are listed in step_ms_blocks",
                    "wall_ms_per_step": m["wall_ms_per_step"]}
        config = cfg(head, args.dif_order)
        config.update({"responses_finite_nonzero": resp_ok, "geometry_seconds": round(t_geo, 2)})
        line = {
            "metric": "Mvox-updates/s", "value": head["value"], "unit": "Mvox/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config, "step_ms_blocks": head["step_ms_blocks"],
            "clocks": clocks, "e2e": head.get("e2e"), "gpu_launches": head["launches"], "roofline": head["roofline"],
            "cpu_baseline": cpu,
        }
        if l4l is not None:
            line["like_for_like"] = {
                "why": "same room, scheme, dtype, steps and slabs as the headline, with the boundary the reference arm runs "
                       "(bench.py --impl reference): divide THIS block by the reference line for a same-config ratio",
                "metric": "Mvox-updates/s", "value": l4l["value"], "unit": "Mvox/s", "ms_per_step": l4l["ms_per_step"],
                "step_ms_blocks": l4l["step_ms_blocks"], "config": cfg(l4l, 0), "e2e": l4l.get("e2e"), "roofline": l4l["roofline"],
                "gpu_launches": l4l["launches"]}
        if world > 1:
            line["halo_ms_per_exchange_alone"] = head["halo_ms_per_exchange_alone"]
            line["halo_edge_kernel_ms_per_step"] = head["roofline"]["edge_kernel_ms_per_step"]
            if inv is not None:
                line["slab_invariance"] = inv["ok"]
                line["slab_invariance_detail"] = {k: v for k, v in inv.items() if k != "ok"}
        if c4 is not None:
            line["c4"] = c4
        if world == 1 and not args.no_variants:
            line["variants"] = run_variants(args)
        emit(line)
    bad = rank == 0 and ((inv is not None and inv["ok"] is False) or (c4 is not None and c4.get("slab_invariance") is False))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        raise SystemExit("slab invariance FAILED: the N-slab responses differ from the single-GPU run")


VARIANTS = [("f32 SRL_FORWARD, frequency-independent admittance (the reference's boundary; parity pinned bit-exact)", "f32", 0, 0, None),
            ("f32 SRL_FORWARD + DIF order 2", "f32", 0, 2, None),
            ("f64 SRL_FORWARD + DIF order 2", "f64", 0, 2, None),
            ("f64 SRL_FORWARD, frequency-independent", "f64", 0, 0, None),
            ("f32 SRL (centred boundary), frequency-independent", "f32", 2, 0, None),
            ("f32 SRL (centred boundary) + DIF order 2", "f32", 2, 2, None),
            ("f32 IISO (27-point), frequency-independent", "f32", 3, 0, None),
            ("f32 IISO + DIF order 2", "f32", 3, 2, None),
            ("f64 IISO + DIF order 2", "f64", 3, 2, None),
            ("BASELINE config 1: 64^3 shoebox, f32 SRL_FORWARD, frequency-independent (launch-bound regime)", "f32", 0, 0, "c1")]


def run_variants(args):
    """The other variants BASELINE config 2 names (fp64, the reference's frequency-independent boundary), the centred and
    interpolated schemes and config 1, each as a short device-resident run of this same script in a child process (100-step
    blocks)."""
    out = []
    for name, dtype, ut, order, workload in VARIANTS:
        if (dtype, ut, order, workload or args.workload) == (args.dtype, args.update_type, args.dif_order, args.workload):
            continue
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", workload or args.workload, "--steps", "100" if not workload else "500",
               "--warmup", "10", "--no-e2e", "--no-cpu-baseline", "--no-variants", "--no-like-for-like", "--c4", "off", "--dtype", dtype,
               "--update-type", str(ut), "--dif-order", str(order)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out.append({"variant": name, "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                        "kernel": d["config"]["kernel"], "roofline_achieved_gbs": d["roofline"]["achieved"],
                        "roofline_frac": d["roofline"]["frac"], "roofline_frac_from_step_time": d["roofline"].get("frac_from_step_time"),
                        "bytes_per_voxel_update": d["roofline"]["bytes_per_voxel_update"],
                        "gpu_launches": d["gpu_launches"]})
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
                            "boundaries": "frequency-independent admittance per material (the reference's boundary)",
                            "note": "the reference contains no digital impedance filter (SURVEY section 0), so its arm runs the boundary "
                                    "it has on the same room; the same-config line of the other arm is its `like_for_like` block",
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
