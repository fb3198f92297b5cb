"""ctypes binding of libpfdtd_b200.so (the C ABI declared in include/pfdtd.h).

This is the only door between Python and the CUDA path: there is no Python or CPU
fallback.  Importing works without a GPU (so the symbol table can be checked), every
compute call raises :class:`PfdtdError` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfdtd_b200.so")
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "include", "pfdtd.h"))

F32, F64 = 0, 1
SRL_FORWARD, SHARED, SRL, IISO, IWB = 0, 1, 2, 3, 4
SRC_HARD, SRC_SOFT, SRC_TRANSPARENT = 0, 1, 2
(OPT_MATIDX_AS_WRITTEN, OPT_SOFT_ACCUMULATE, OPT_KERNEL, OPT_GLOBAL_Z_FIRST, OPT_GLOBAL_Z_DIM,
 OPT_DOUBLE_PAD_AS_WRITTEN, OPT_USE_GRAPH, OPT_OVERLAP, OPT_TMA_CHUNK, OPT_TMA_TILE, OPT_TIME_KERNELS, OPT_TMA_HINTS, OPT_DIF_ORDER,
 OPT_PEER_STORES, OPT_FUSE_SRCREC) = range(1, 16)
KERNEL_AUTO, KERNEL_TMA, KERNEL_PLAIN = 0, 1, 2

INTERRUPT_CB = C.CFUNCTYPE(C.c_int)
PROGRESS_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_float)


class PfdtdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pfdtd error {code}: {msg}")
        self.code = code


def declared_symbols(header_path: str = HEADER_PATH):
    """Names of every function include/pfdtd.h declares."""
    txt = open(header_path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pfdtd_[a-z0-9_]+)\s*\(", txt)))


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise PfdtdError(-1, f"{path} is missing: build it with `python -m parallelfdtd_b200.build` "
                             "(there is no fallback path)")
    lib = C.CDLL(path)
    lib.pfdtd_last_error.restype = C.c_char_p
    lib.pfdtd_version.restype = C.c_char_p
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load_library(os.environ.get("PFDTD_LIB_PATH", LIB_PATH))   # override: A/B builds of the same ABI (tools/)
    return _lib


def _check(rc):
    if rc != 0:
        raise PfdtdError(rc, lib().pfdtd_last_error().decode(errors="replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().pfdtd_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def release_cached_memory(device: int = -1) -> None:
    """Give the device blocks the library keeps for the next solver back to the driver (pfdtd_release_cached_memory)."""
    _check(lib().pfdtd_release_cached_memory(C.c_int(device)))


def voxelize(vertices, indices, dx, triangle_material=None):
    """voxelizeGeometry (reference src/kernels/voxelizationUtils.cu:47-146) on the device: closed triangle mesh ->
    voxelizer-style ``bid`` and material volumes ``[vz][vy][vx]``."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
    m = None if triangle_material is None else np.ascontiguousarray(triangle_material, dtype=np.uint8)
    vx, vy, vz = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(lib().pfdtd_voxelize_dims(_ptr(v), C.c_uint32(len(v)), C.c_float(dx), C.byref(vx), C.byref(vy), C.byref(vz)))
    bid = np.empty((vz.value, vy.value, vx.value), dtype=np.uint8)
    mat = np.empty_like(bid)
    _check(lib().pfdtd_voxelize(_ptr(v), C.c_uint32(len(v)), _ptr(t), C.c_uint32(len(t)), _ptr(m), C.c_float(dx), _ptr(bid), _ptr(mat)))
    return bid, mat


def partition_indexing(dim_z: int, n: int):
    """CudaMesh::getPartitionIndexing (reference src/kernels/cudaMesh.h:280-307)."""
    first = (C.c_uint32 * n)()
    size = (C.c_uint32 * n)()
    _check(lib().pfdtd_partition_indexing(C.c_uint32(dim_z), C.c_uint32(n), first, size))
    return list(first), list(size)


class Solver:
    """Thin object wrapper over a ``pfdtd_solver*``; method names follow the C ABI."""

    def __init__(self):
        self._h = C.c_void_p()
        _check(lib().pfdtd_create(C.byref(self._h)))
        self.dtype = F32
        self._keep = []

    def close(self):
        if self._h:
            lib().pfdtd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def np_dtype(self):
        return np.float32 if self.dtype == F32 else np.float64

    def set_option(self, opt, value):
        _check(lib().pfdtd_set_option(self._h, C.c_int(opt), C.c_int64(int(value))))

    def get_option(self, opt):
        v = C.c_int64(0)
        _check(lib().pfdtd_get_option(self._h, C.c_int(opt), C.byref(v)))
        return v.value

    def set_scheme_coefficients(self, d):
        arr = (C.c_double * 4)(*[float(v) for v in d]) if d is not None else None
        _check(lib().pfdtd_set_scheme_coefficients(self._h, arr))

    def setup_mesh(self, bid, mat, block=(32, 4, 1), element_type=SRL_FORWARD, dtype=F32, params=None, materials=None):
        bid = np.ascontiguousarray(bid, dtype=np.uint8)
        mat = np.ascontiguousarray(mat, dtype=np.uint8)
        assert bid.ndim == 3 and bid.shape == mat.shape
        vz, vy, vx = bid.shape
        self.dtype = dtype
        params = np.ascontiguousarray(params, dtype=self.np_dtype)
        materials = np.ascontiguousarray(materials, dtype=self.np_dtype)
        assert params.size == 4 and materials.ndim == 2 and materials.shape[1] == 20
        _check(lib().pfdtd_setup_mesh(self._h, _ptr(bid), _ptr(mat), C.c_uint32(vx), C.c_uint32(vy), C.c_uint32(vz),
                                      C.c_uint32(block[0]), C.c_uint32(block[1]), C.c_uint32(block[2]),
                                      C.c_uint32(element_type), C.c_int(dtype), _ptr(params), _ptr(materials),
                                      C.c_uint32(materials.shape[0])))

    def make_partition(self, n=1, devices=None):
        dl = None
        if devices is not None:
            assert len(devices) == n
            dl = (C.c_uint32 * n)(*devices)
        _check(lib().pfdtd_make_partition(self._h, C.c_uint32(n), dl))

    def dims(self):
        x, y, z = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().pfdtd_get_dims(self._h, C.byref(x), C.byref(y), C.byref(z)))
        return x.value, y.value, z.value

    def counts(self):
        n, a, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().pfdtd_get_counts(self._h, C.byref(n), C.byref(a), C.byref(b)))
        return n.value, a.value, b.value

    def num_partitions(self):
        n = C.c_uint32()
        _check(lib().pfdtd_get_num_partitions(self._h, C.byref(n)))
        return n.value

    def partition(self, k):
        f, s, d = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().pfdtd_get_partition(self._h, C.c_uint32(k), C.byref(f), C.byref(s), C.byref(d)))
        return f.value, s.value, d.value

    def element_idx_and_partition(self, x, y, z):
        p, e = C.c_int(), C.c_int64()
        _check(lib().pfdtd_get_element_idx_and_partition(self._h, C.c_uint32(x), C.c_uint32(y), C.c_uint32(z),
                                                         C.byref(p), C.byref(e)))
        return p.value, e.value

    def export_partition_nodes(self, k):
        X, Y, _ = self.dims()
        _, nz, _ = self.partition(k)
        pos = np.empty((nz, Y, X), dtype=np.uint8)
        mat = np.empty((nz, Y, X), dtype=np.uint8)
        _check(lib().pfdtd_export_partition_nodes(self._h, C.c_uint32(k), _ptr(pos), _ptr(mat)))
        return pos, mat

    def export_partition_pressure(self, k, which=0):
        X, Y, _ = self.dims()
        _, nz, _ = self.partition(k)
        out = np.empty((nz, Y, X), dtype=self.np_dtype)
        _check(lib().pfdtd_export_partition_pressure(self._h, C.c_uint32(k), C.c_int(which), _ptr(out)))
        return out

    def capture_slice(self, slice_idx, orientation, with_position=False):
        """captureSliceFast (reference visualizationUtils.cu:127-254): [Y][X] / [Z][X] / [Z][Y] for orientation 0 / 1 / 2."""
        X, Y, Z = self.dims()
        shape = {0: (Y, X), 1: (Z, X), 2: (Z, Y)}.get(int(orientation), (0, 0))
        out = np.empty(shape, dtype=self.np_dtype)
        pos = np.empty(shape, dtype=np.uint8) if with_position else None
        _check(lib().pfdtd_capture_slice(self._h, C.c_uint32(slice_idx), C.c_uint32(orientation), _ptr(out), _ptr(pos)))
        return (out, pos) if with_position else out

    def capture_mesh(self):
        """captureMesh (reference visualizationUtils.cu:111-125): the whole current field [Z][Y][X]."""
        X, Y, Z = self.dims()
        out = np.empty((Z, Y, X), dtype=self.np_dtype)
        _check(lib().pfdtd_capture_mesh(self._h, _ptr(out)))
        return out

    def set_sample(self, x, y, z, v):
        _check(lib().pfdtd_set_sample(self._h, C.c_uint32(x), C.c_uint32(y), C.c_uint32(z), C.c_double(v)))

    def add_sample(self, x, y, z, v):
        _check(lib().pfdtd_add_sample(self._h, C.c_uint32(x), C.c_uint32(y), C.c_uint32(z), C.c_double(v)))

    def get_sample(self, x, y, z):
        v = C.c_double()
        _check(lib().pfdtd_get_sample(self._h, C.c_uint32(x), C.c_uint32(y), C.c_uint32(z), C.byref(v)))
        return v.value

    def set_sample_at(self, x, y, z, part, v):
        _check(lib().pfdtd_set_sample_at(self._h, C.c_uint32(x), C.c_uint32(y), C.c_uint32(z), C.c_uint32(part), C.c_double(v)))

    def get_sample_at(self, x, y, z, part):
        v = C.c_double()
        _check(lib().pfdtd_get_sample_at(self._h, C.c_uint32(x), C.c_uint32(y), C.c_uint32(z), C.c_uint32(part), C.byref(v)))
        return v.value

    def switch_halos(self):
        _check(lib().pfdtd_switch_halos(self._h))

    def flip_pressure_pointers(self):
        _check(lib().pfdtd_flip_pressure_pointers(self._h))

    def reset_pressures(self):
        _check(lib().pfdtd_reset_pressures(self._h))

    def set_sources(self, xyz, types, samples):
        xyz = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        types = np.ascontiguousarray(types, dtype=np.int32).reshape(-1)
        samples = np.ascontiguousarray(samples, dtype=self.np_dtype)
        n = xyz.shape[0]
        samples = samples.reshape(n, -1) if n else samples.reshape(0, 0)
        _check(lib().pfdtd_set_sources(self._h, C.c_uint32(n), _ptr(xyz), _ptr(types), _ptr(samples),
                                       C.c_uint32(samples.shape[1] if n else 0)))

    def set_receivers(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
        self._n_rec = xyz.shape[0]
        _check(lib().pfdtd_set_receivers(self._h, C.c_uint32(xyz.shape[0]), _ptr(xyz)))

    def run(self, n_steps, interrupt=None, progress=None):
        """launchFDTD3d[Double]: returns (responses [n_rec][n_steps], seconds_per_step)."""
        n_rec = getattr(self, "_n_rec", 0)
        out = np.zeros((n_rec, n_steps), dtype=self.np_dtype)
        icb = INTERRUPT_CB(interrupt) if interrupt else C.cast(None, INTERRUPT_CB)
        pcb = PROGRESS_CB(progress) if progress else C.cast(None, PROGRESS_CB)
        sps = C.c_float()
        rc = lib().pfdtd_run(self._h, C.c_uint32(n_steps), _ptr(out), icb, pcb, C.byref(sps))
        if rc not in (0, 5):
            _check(rc)
        return out, sps.value

    def step(self, step, direction=1, response=None, n_steps_total=0):
        _check(lib().pfdtd_step(self._h, C.c_uint32(step), C.c_int(direction), _ptr(response), C.c_uint32(n_steps_total)))

    def reserve_steps(self, n):
        _check(lib().pfdtd_reserve_steps(self._h, C.c_uint32(n)))

    def enqueue_steps(self, first, n):
        _check(lib().pfdtd_enqueue_steps(self._h, C.c_uint32(first), C.c_uint32(n)))

    def sync(self):
        _check(lib().pfdtd_sync(self._h))

    def fetch_responses(self, n_steps):
        n_rec = getattr(self, "_n_rec", 0)
        out = np.zeros((n_rec, n_steps), dtype=self.np_dtype)
        _check(lib().pfdtd_fetch_responses(self._h, _ptr(out), C.c_uint32(n_steps)))
        return out

    def last_timing(self):
        t, k, n = C.c_float(), C.c_float(), C.c_uint32()
        _check(lib().pfdtd_last_timing(self._h, C.byref(t), C.byref(k), C.byref(n)))
        return t.value, k.value, n.value

    def last_timing_detail(self):
        """(bulk_kernel_ms, n_bulk_launches, bulk_planes, edge_kernel_ms, n_edge_launches) of the last timed enqueue"""
        b, e = C.c_float(), C.c_float()
        nb, ne, pl = C.c_uint32(), C.c_uint32(), C.c_uint64()
        _check(lib().pfdtd_last_timing_detail(self._h, C.byref(b), C.byref(nb), C.byref(pl), C.byref(e), C.byref(ne)))
        return b.value, nb.value, pl.value, e.value, ne.value

    def time_halo_exchange(self, reps=20):
        ms = C.c_float()
        _check(lib().pfdtd_time_halo_exchange(self._h, C.c_uint32(reps), C.byref(ms)))
        return ms.value

    def comm_release(self):
        if self._h:
            _check(lib().pfdtd_comm_release(self._h))

    def halo_transport(self):
        buf = C.create_string_buffer(256)
        _check(lib().pfdtd_halo_transport(self._h, buf, C.c_size_t(256)))
        return buf.value.decode()

    def last_halo_ms(self):
        h = C.c_float()
        _check(lib().pfdtd_last_halo_ms(self._h, C.byref(h)))
        return h.value

    def comm_init(self, id128: bytes, rank: int, nranks: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(id128)
        _check(lib().pfdtd_comm_init(self._h, buf, C.c_int(rank), C.c_int(nranks)))

    def kernel_name(self):
        buf = C.create_string_buffer(256)
        _check(lib().pfdtd_kernel_name(self._h, buf, C.c_size_t(256)))
        return buf.value.decode()

    def launch_count(self):
        n = C.c_uint64()
        _check(lib().pfdtd_launch_count(self._h, C.byref(n)))
        return n.value


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(lib().pfdtd_comm_unique_id(buf))
    return bytes(buf)
