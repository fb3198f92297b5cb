"""GPU, >= 2 devices: z-slab decomposition over real devices -- in one process (peer copies, like the
reference's makePartition(n, devices)) and one process per GPU (NCCL halo exchange, torchrun).
The invariant is the reference's (tests/CudaMeshTest.cpp:472-575): responses do not depend on the slab count."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {c["name"]: c for c in fc.parity_cases()}


def _need(gpu, n):
    if gpu < n:
        pytest.skip(f"needs {n} GPUs, box has {gpu}")


@pytest.mark.parametrize("name", ["shoebox_48x40x49_ctr_f64_6mat_5parts", "hall_96x128x64_fwd_f32_5mat_oct1"])
def test_slabs_on_two_devices_in_process(capi, gpu, name):
    _need(gpu, 2)
    case = CASES[name]
    base, _, _ = fc.run_ours(capi, case, n_parts=1)
    r_or, _, _ = fc.run_oracle(case)
    assert np.array_equal(base, r_or)
    for devs in ([0, 1], [1, 0], [0, 1, 0, 1, 0], [0, 1, 1]):
        r, nodes, info = fc.run_ours(capi, case, n_parts=len(devs), devices=devs)
        assert np.array_equal(r, base), devs
    r, _, _ = fc.run_ours(capi, case, n_parts=2, devices=[0, 1], opts=[(capi.OPT_OVERLAP, 0)])
    assert np.array_equal(r, base)


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["PFDTD_ROOT"])
    from oracle import oracle
    from parallelfdtd_b200 import capi, slabs, synth
    from tests import fdtd_cases as fc

    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    for name, steps in (("shoebox_48x40x49_ctr_f64_6mat_5parts", 200), ("hall_96x128x64_fwd_f32_5mat_oct1", 150)):
        case = {c["name"]: c for c in fc.parity_cases()}[name]
        double = case["double"]
        Z, Y, X = case["bid"].shape
        prm = oracle.params(fc.LAM, case["octave"], double)
        # halo transports: peer-mapped stores from the edge launches (CUDA IPC, the default), NCCL send/recv, no overlap
        for overlap, peer in ((1, 1), (1, 0), (0, 1)):
            ss = slabs.SlabSolver(capi, (X, Y, Z), lambda a, b: (case["bid"][a:b], case["mat"][a:b]), block=case["block"],
                                  element_type=case["update_type"], dtype=capi.F64 if double else capi.F32, params=prm,
                                  materials=case["materials"], rank=rank, world=world, device=lr,
                                  options=[(capi.OPT_OVERLAP, overlap), (capi.OPT_PEER_STORES, peer)])
            ss.connect()
            transport = ss.solver.halo_transport()
            thin = min(ss.plan.size(r) for r in range(world)) - 2 < 4          # a slab too thin to split keeps NCCL on its interfaces
            if overlap and peer and not thin and os.environ.get("PFDTD_EXPECT_IPC", "1") == "1":
                assert "peer-mapped" in transport, transport
            if not (overlap and peer):
                assert "peer-mapped" not in transport, transport
            src = np.asarray(case["sources"], dtype=np.int32).reshape(-1, 6)
            ss.set_sources(src[:, :3], src[:, 3], fc.source_table(case)[:, :steps])
            ss.set_receivers(case["receivers"])
            ss.solver.reserve_steps(steps)
            ss.solver.enqueue_steps(0, steps // 2)          # two enqueue blocks: exercises the step-counter hand-over
            ss.solver.enqueue_steps(steps // 2, steps - steps // 2)
            ss.solver.sync()
            merged = ss.responses(steps)
            # node bytes of the rank's slab are the slices of the global volume
            pos, m, _, _ = oracle.setup_mesh(case["bid"], case["mat"], case["block"], case["update_type"], double)
            z0, nz = ss.plan.slab(rank)
            lp, lm = ss.solver.export_partition_nodes(0)
            assert np.array_equal(lp, pos[z0:z0 + nz]) and np.array_equal(lm, m[z0:z0 + nz]), "slab node bytes differ"
            ss.close()
            if rank == 0:
                ref, _, _ = fc.run_oracle(case, n_parts=1)
                assert np.abs(ref).max() > 0
                assert np.array_equal(merged, ref[:, :steps]), (name, overlap, peer, transport, float(np.abs(merged - ref[:, :steps]).max()))
            dist.barrier()
    dist.destroy_process_group()
    print("MP_OK", rank)
''')


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_process_per_gpu_nccl_halo(tmp_path, capi, gpu, world):
    _need(gpu, world)
    script = tmp_path / "mp_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PFDTD_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-4000:]
    assert r.stdout.count("MP_OK") == world
