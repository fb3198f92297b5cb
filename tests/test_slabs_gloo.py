"""CPU, world_size 2 (gloo): the host-side logic of the one-process-per-GPU z-slab decomposition
(parallelfdtd_b200/slabs.py): slab index sets identical to the reference's getPartitionIndexing,
source/receiver ownership (cudaMesh.h:251-266, 321-338), NCCL-id hand-off, response merge, max-over-ranks
timing, and a full emulation of the per-step halo protocol with torch.distributed send/recv whose
receiver responses must be bit-identical to the single-process oracle."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import oracle
from parallelfdtd_b200 import slabs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("Z,N", [(100, 13), (49, 2), (49, 5), (512, 8), (4096, 8), (64, 1)])
def test_slab_plan_equals_reference_partition_indexing(Z, N):
    plan = slabs.SlabPlan(Z, N)
    first, size = oracle.partition_indexing(Z, N)
    assert [plan.slab(r) for r in range(N)] == list(zip(first, size))
    # every slice 1..Z-2 is updated by exactly one rank (kernels3d.cu:112-113)
    owner = np.zeros(Z, dtype=int)
    for r in range(N):
        a, b = plan.updated(r)
        owner[a:b] += 1
    assert (owner[1:Z - 1] == 1).all() and owner[0] == 0 and owner[Z - 1] == 0


def test_slab_plan_ownership_rules():
    plan = slabs.SlabPlan(49, 2)                     # slab0 = z 0..24, slab1 = z 23..48
    assert plan.holders(23) == [0, 1] and plan.holders(24) == [0, 1] and plan.holders(22) == [0] and plan.holders(25) == [1]
    assert plan.owner(23) == 0 and plan.owner(24) == 0 and plan.owner(25) == 1 and plan.owner(60) == -1
    with pytest.raises(ValueError):
        slabs.SlabPlan(4, 8)


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["PFDTD_ROOT"])
    from oracle import oracle
    from parallelfdtd_b200 import slabs, synth

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    # 1. the 128-byte id reaches every rank
    uid = bytes(range(128)) if rank == 0 else None
    got = slabs.broadcast_bytes(uid, 128, 0)
    assert got == bytes(range(128)), got[:8]

    # 2. max over ranks
    assert slabs.max_over_ranks(1.0 + rank) == float(world)

    # 3. emulate the multi-process step loop on the CPU: every rank steps ITS slab with the oracle's
    #    single-slab update and exchanges one plane each way per step with send/recv.
    dims = (24, 20, 33)
    steps = 60
    lam = float(np.sqrt(1.0 / 3.0))
    bid, mat = synth.shoebox(dims, 6)
    pos, m, _, _ = oracle.setup_mesh(bid, mat, (8, 4, 1), 2, True)
    Z, Y, X = pos.shape
    prm = oracle.params(lam, 0, True)
    tab = synth.material_table(list(np.linspace(0.99, 0.5, 6))).astype(np.float64)
    src_xyz = [(5, 6, 16), (12, 9, 3)]
    rec_xyz = [(10, 9, 15), (10, 9, 16), (10, 9, 17), (3, 3, 30), (3, 3, 2)]
    src = np.stack([oracle.source_samples(0, steps, double=True), oracle.source_samples(1, steps, double=True)])
    plan = slabs.SlabPlan(Z, world)
    z0, nz = plan.slab(rank)
    # slab volumes generated per rank equal the slices of the global volume
    sb, sm = synth.shoebox(dims, 6, z0, z0 + nz)
    assert np.array_equal(sb, bid[z0:z0 + nz]) and np.array_equal(sm, mat[z0:z0 + nz])
    lpos, lmat = np.ascontiguousarray(pos[z0:z0 + nz]), np.ascontiguousarray(m[z0:z0 + nz])
    cur = np.zeros((nz, Y, X)); past = np.zeros((nz, Y, X))
    mine_src = [i for i, p in enumerate(src_xyz) if rank in plan.holders(p[2])]
    mine_rec = [i for i, p in enumerate(rec_xyz) if plan.owner(p[2]) == rank]
    local = np.zeros((len(rec_xyz), steps))
    for n in range(steps):
        for i in mine_src:
            x, y, z = src_xyz[i]
            cur[z - z0, y, x] = src[i, n]
        # one step of the oracle on the slab: state in = (cur, past) via a 1-step run is not exposed, so
        # use the slab as its own tiny domain: sources carry the whole state (hard sources on every voxel
        # would be O(n^2)); instead call the exported single-step helper
        new = oracle.step_slab(lpos, lmat, 2, prm, tab, cur, past)
        past, cur = cur, new
        reqs = []
        if rank + 1 < world:
            reqs.append(dist.isend(torch.from_numpy(cur[nz - 2].copy()), rank + 1))
            up = torch.empty((Y, X), dtype=torch.float64); reqs.append(dist.irecv(up, rank + 1))
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(cur[1].copy()), rank - 1))
            dn = torch.empty((Y, X), dtype=torch.float64); reqs.append(dist.irecv(dn, rank - 1))
        for r in reqs:
            r.wait()
        if rank + 1 < world:
            cur[nz - 1] = up.numpy()
        if rank > 0:
            cur[0] = dn.numpy()
        for i in mine_rec:
            x, y, z = rec_xyz[i]
            local[i, n] = cur[z - z0, y, x]
    merged = slabs.merge_responses(local, [p[2] for p in rec_xyz], plan, rank)
    ref, _ = oracle.run(pos, m, 2, prm, tab, src_xyz, [0, 0], src, rec_xyz, steps, 1)
    assert np.abs(ref).max() > 0
    assert np.array_equal(merged, ref), float(np.abs(merged - ref).max())
    dist.barrier()
    dist.destroy_process_group()
    print("WORKER_OK", rank)
''')


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_multi_process_slab_protocol_on_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PFDTD_ROOT=ROOT, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("WORKER_OK") == world
