"""Case / result files exchanged with oracle/_ref/ref_fdtd (the reference's own CUDA build).

TEST INFRASTRUCTURE ONLY.  A *case* is a plain dict:
  bid, mat        uint8 [vz][vy][vx] voxelizer-style volumes
  block           (bx, by, bz)
  update_type     0 SRL_FORWARD, 1 SHARED, 2 SRL
  double          bool
  steps, octave   ints
  n_parts, devices
  materials       float32 [n_mat][20]
  sources         list of (x, y, z, src_type, input_type, data_idx)   final element coordinates
  receivers       list of (x, y, z)
  input_data      list of float64 arrays (DATA sources)
"""
from __future__ import annotations

import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(_HERE, "_ref", "ref_fdtd")


def ref_available():
    return os.path.exists(REF_BIN)


def write_case(path, case):
    bid = np.ascontiguousarray(case["bid"], dtype=np.uint8)
    mat = np.ascontiguousarray(case["mat"], dtype=np.uint8)
    vz, vy, vx = bid.shape
    n_parts = int(case.get("n_parts", 1))
    devices = list(case.get("devices", [0] * n_parts))
    materials = np.ascontiguousarray(case["materials"], dtype=np.float32)
    sources = np.asarray(case.get("sources", []), dtype=np.int32).reshape(-1, 6)
    receivers = np.asarray(case.get("receivers", []), dtype=np.int32).reshape(-1, 3)
    data = [np.ascontiguousarray(d, dtype=np.float64) for d in case.get("input_data", [])]
    bx, by, bz = case.get("block", (32, 4, 1))
    with open(path, "wb") as f:
        f.write(b"PFDTDCAS")
        f.write(struct.pack("<16I", 1, vx, vy, vz, bx, by, bz, int(case["update_type"]), int(bool(case.get("double", False))),
                            int(case["steps"]), int(case.get("octave", 0)), n_parts, materials.shape[0], sources.shape[0],
                            receivers.shape[0], len(data)))
        f.write(np.asarray(devices, dtype=np.uint32).tobytes())
        f.write(materials.tobytes())
        f.write(sources.tobytes())
        f.write(receivers.tobytes())
        for d in data:
            f.write(struct.pack("<I", d.size))
            f.write(d.tobytes())
        f.write(bid.tobytes())
        f.write(mat.tobytes())


def read_result(path, with_nodes=False):
    with open(path, "rb") as f:
        assert f.read(8) == b"PFDTDOUT"
        X, Y, Z, n_parts, n_rec, steps, is_double, n_air, n_bnd = struct.unpack("<9I", f.read(36))
        t_ret, wall, wall_e2e = struct.unpack("<3d", f.read(24))
        parts = [struct.unpack("<2I", f.read(8)) for _ in range(n_parts)]
        resp = np.frombuffer(f.read(8 * n_rec * steps), dtype=np.float64).reshape(n_rec, steps).copy()
        nodes = []
        if with_nodes:
            for first, size in parts:
                n = size * X * Y
                pos = np.frombuffer(f.read(n), dtype=np.uint8).reshape(size, Y, X).copy()
                mat = np.frombuffer(f.read(n), dtype=np.uint8).reshape(size, Y, X).copy()
                nodes.append((pos, mat))
    if not is_double:
        resp = resp.astype(np.float32)
    return dict(dims=(X, Y, Z), n_parts=n_parts, steps=steps, double=bool(is_double), n_air=n_air, n_boundary=n_bnd,
                time_per_step_returned=t_ret, wall_seconds=wall, wall_e2e_seconds=wall_e2e, partitions=parts, responses=resp, nodes=nodes)


def run_reference(case, workdir, dump_nodes=False, timeout=600, warmup_steps=0):
    """Run the reference's own CUDA hot path (needs a GPU). Returns read_result(...)."""
    os.makedirs(workdir, exist_ok=True)
    cp = os.path.join(workdir, "case.bin")
    op = os.path.join(workdir, "out.bin")
    write_case(cp, case)
    r = subprocess.run([REF_BIN, cp, op, "1" if dump_nodes else "0", str(int(warmup_steps))], cwd=workdir, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_fdtd failed ({r.returncode}): {r.stdout}\n{r.stderr}")
    res = read_result(op, with_nodes=dump_nodes)
    res["stdout"] = r.stdout
    return res
