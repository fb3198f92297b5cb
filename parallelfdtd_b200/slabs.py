"""One-process-per-GPU z-slab decomposition (torchrun / torch.distributed plumbing).

The reference splits the domain into z-slabs with one-slice halos inside ONE process
(CudaMesh::getPartitionIndexing / makePartition / switchHalos, reference
src/kernels/cudaMesh.h:280-307, 648-751, 432-463).  Here every rank owns exactly the slab the
reference would give partition `rank` of `world` partitions -- same first slice, same size, same
halo planes -- and the per-step exchange of one plane each way runs over NVLink: the edge launch of a
slab stores its plane straight into the neighbour process's halo plane through a CUDA-IPC mapping (NCCL
point-to-point where that mapping is not available), overlapped with the interior update.  torch.distributed
is used only to hand the 128-byte NCCL id to every rank, for barriers, and to merge the
receiver responses (each receiver is recorded by the first slab that contains its slice,
cudaMesh.h:251-266).

Everything that does not touch the GPU (`SlabPlan`, `merge_responses`, `broadcast_bytes`) works on
the gloo backend and is covered by world_size-2 CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class SlabPlan:
    """Slab index sets of a Z-slice domain over `world` ranks (reference getPartitionIndexing)."""
    dim_z: int
    world: int

    def __post_init__(self):
        if self.world < 1 or self.dim_z // self.world < 1:
            raise ValueError(f"cannot split {self.dim_z} slices over {self.world} ranks")

    @property
    def part_size(self) -> int:
        return self.dim_z // self.world

    def first(self, rank: int) -> int:
        return rank * self.part_size - (1 if rank > 0 else 0)

    def size(self, rank: int) -> int:
        n = self.part_size + (1 if rank > 0 else 0) + (1 if rank < self.world - 1 else 0)
        if rank != 0 and rank == self.world - 1:
            n += self.dim_z - (rank + 1) * self.part_size
        return n

    def slab(self, rank: int) -> Tuple[int, int]:
        """(first global slice, number of slices) held by `rank`, halos included."""
        return self.first(rank), self.size(rank)

    def updated(self, rank: int) -> Tuple[int, int]:
        """global slices [a, b) this rank updates: local 1..size-2 (kernels3d.cu:112-113)."""
        f, n = self.slab(rank)
        return f + 1, f + n - 1

    def holders(self, z: int) -> List[int]:
        """every rank whose slab contains slice z (sources are injected in all of them, cudaMesh.h:321-338)."""
        return [r for r in range(self.world) if self.first(r) <= z <= self.first(r) + self.size(r) - 1]

    def owner(self, z: int) -> int:
        """first rank whose slab contains z (receivers, cudaMesh.h:251-266); -1 if none."""
        h = self.holders(z)
        return h[0] if h else -1


def broadcast_bytes(payload: bytes | None, n: int, src: int = 0) -> bytes:
    """Hand `n` bytes from rank `src` to every rank (the NCCL unique id)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def merge_responses(local: np.ndarray, rec_z: Sequence[int], plan: SlabPlan, rank: int) -> np.ndarray:
    """[n_rec][steps] on every rank: row r is taken from plan.owner(rec_z[r]); rows of receivers a
    rank does not own are ignored (they are zero in the C ABI's output)."""
    import torch
    import torch.distributed as dist
    mine = np.array([plan.owner(int(z)) == rank for z in rec_z], dtype=bool)
    contrib = np.where(mine[:, None], local, 0).astype(local.dtype, copy=False)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(contrib)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)     # exactly one non-zero contribution per row: exact
    return t.cpu().numpy()


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class SlabSolver:
    """The rank-local solver of a multi-process run.

    `make_volumes(z0, z1)` must return the voxelizer-style (bid, mat) volumes of global slices
    [z0, z1) of the UNPADDED domain `dims` = (vx, vy, vz); vz must already be a multiple of
    block[2] (the reference pads the global volume before partitioning, cudaMesh.cu:253-304).
    """

    def __init__(self, capi, dims, make_volumes, *, block=(32, 4, 1), element_type=0, dtype=0, params=None,
                 materials=None, rank=0, world=1, device=0, options=()):
        vx, vy, vz = dims
        if vz % block[2]:
            raise ValueError("global z extent must be a multiple of block z for the multi-process layout")
        self.capi = capi
        self.plan = SlabPlan(vz, world)
        self.rank, self.world = rank, world
        self.z0, n = self.plan.slab(rank)
        self.z1 = self.z0 + n
        bid, mat = make_volumes(self.z0, self.z1)
        assert bid.shape == (n, vy, vx), (bid.shape, (n, vy, vx))
        s = capi.Solver()
        for k, v in options:
            s.set_option(k, v)
        s.set_option(capi.OPT_GLOBAL_Z_FIRST, self.z0)
        s.set_option(capi.OPT_GLOBAL_Z_DIM, vz)
        s.setup_mesh(bid, mat, block, element_type, dtype, params, materials)
        s.make_partition(1, [device])
        self.solver = s
        self._connected = False
        self._rec_z: List[int] = []

    def connect(self, uid: bytes | None = None) -> bytes | None:
        """Attach the NCCL communicator for the halo exchange (collective over all ranks).  Without `uid` a
        fresh 128-byte id is made on rank 0 and handed round; passing the id returned by an earlier connect()
        reuses that communicator (the library caches communicators per id for the life of the process)."""
        if self.world == 1:
            return None
        import torch.distributed as dist
        if uid is None:
            uid = self.capi.comm_unique_id() if self.rank == 0 else None
            uid = broadcast_bytes(uid, 128, 0)
        self.solver.comm_init(uid, self.rank, self.world)
        self._connected = True
        dist.barrier()
        return uid

    def set_sources(self, xyz, types, samples):
        self.solver.set_sources(xyz, types, samples)        # global coordinates; the library keeps those in its slab

    def set_receivers(self, xyz):
        xyz = np.asarray(xyz, dtype=np.int32).reshape(-1, 3)
        self._rec_z = [int(z) for z in xyz[:, 2]]
        self.solver.set_receivers(xyz)

    def responses(self, n_steps):
        local = self.solver.fetch_responses(n_steps)
        if self.world == 1:
            return local
        return merge_responses(local, self._rec_z, self.plan, self.rank)

    def close(self):
        """Collective when world > 1: every rank first unmaps its neighbours' slabs (CUDA IPC), and only after all have
        done so does any rank free its own (include/pfdtd.h, pfdtd_comm_release)."""
        if self.world > 1 and self._connected:
            import torch.distributed as dist
            self.solver.comm_release()
            dist.barrier()
        self.solver.close()
