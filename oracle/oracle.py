"""ctypes loader for oracle/libpfdtd_oracle.so (oracle/fdtd_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- see the header of fdtd_oracle.cpp.  Builds the library with
``make -C oracle oracle`` when it is missing (gcc is part of the image).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfdtd_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "fdtd_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.pfo_run_f32.restype = C.c_double
        _lib.pfo_run_f64.restype = C.c_double
        _lib.pfo_dx.restype = C.c_float
        _lib.pfo_regular_sample_f32.restype = C.c_float
        _lib.pfo_regular_sample_f64.restype = C.c_double
        _lib.pfo_transparent_sample_f32.restype = C.c_float
        _lib.pfo_transparent_sample_f64.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().pfo_num_threads()


def partition_indexing(dim, n):
    first = np.zeros(n, dtype=np.int64)
    size = np.zeros(n, dtype=np.int64)
    lib().pfo_partition_indexing(C.c_int64(dim), C.c_int(n), _p(first), _p(size))
    return first.tolist(), size.tolist()


def padded_dims(dims, block):
    out = np.zeros(3, dtype=np.uint32)
    lib().pfo_padded_dims(*(C.c_uint32(int(v)) for v in dims), *(C.c_uint32(int(v)) for v in block), _p(out))
    return tuple(int(v) for v in out)


def pad_with_zeros(vol, block):
    """vol: [z][y][x] uint8 -> padded [Z][Y][X] (reference padWithZeros semantics)."""
    vol = np.ascontiguousarray(vol, dtype=np.uint8)
    vz, vy, vx = vol.shape
    X, Y, Z = padded_dims((vx, vy, vz), block)
    out = np.zeros((Z, Y, X), dtype=np.uint8)
    lib().pfo_pad_with_zeros(_p(vol), C.c_uint32(vx), C.c_uint32(vy), C.c_uint32(vz),
                             C.c_uint32(block[0]), C.c_uint32(block[1]), C.c_uint32(block[2]), _p(out))
    return out


def translate(pos, mat, centred):
    """in-place toBilbao / toKowalczyk; returns (n_air, n_boundary)."""
    assert pos.flags.c_contiguous and mat.flags.c_contiguous and pos.dtype == np.uint8 and mat.dtype == np.uint8
    counts = np.zeros(2, dtype=np.uint64)
    fn = lib().pfo_to_kowalczyk if centred else lib().pfo_to_bilbao
    fn(_p(pos), _p(mat), C.c_uint64(pos.size), _p(counts))
    return int(counts[0]), int(counts[1])


def setup_mesh(bid, mat, block=(32, 4, 1), element_type=0, dtype_is_double=False, double_pad_as_written=False):
    """CudaMesh::setupMesh[Double] on the host: returns (pos, mat, n_air, n_boundary)."""
    blk = list(block)
    if dtype_is_double and double_pad_as_written:
        blk[1] = blk[0]
    pos = pad_with_zeros(bid, blk)
    m = pad_with_zeros(mat, blk)
    # reference: types 0,1,(3) -> Bilbao / forward byte, 2 -> Kowalczyk (cudaMesh.cu:70-73,134-137); the enum values
    # 3 (IISO) and 4 (IWB) appended by this build use the forward byte in both precisions
    centred = element_type == 2
    air, bnd = translate(pos, m, centred)
    return pos, m, air, bnd


def params(lam, octave, double=False):
    out = np.zeros(4, dtype=np.float64 if double else np.float32)
    (lib().pfo_params_f64 if double else lib().pfo_params_f32)(C.c_double(lam), C.c_uint(octave), _p(out))
    return out


def interp_coefficients(update_type, lam2):
    """d1..d4 of the compact explicit family as the library derives them (pfdtd_api.cu)."""
    l2 = float(lam2)
    if update_type == 3:      # IISO: a = 1/6, b = 0
        return [l2 / 3, l2 / 6, 0.0, 2 - 4 * l2]
    if update_type == 4:      # IWB: a = 1/4, b = 1/16
        return [l2 / 4, l2 / 8, l2 / 16, 2 - 3.5 * l2]
    raise ValueError(update_type)


def interp_lambda(update_type):
    return {3: float(np.sqrt(3.0) / 2), 4: 1.0}[update_type]


def params_interp(lam, octave, d, double=False):
    """8-entry parameter vector of the interpolated schemes: [lam, lam^2, 1/3, octave, d1, d2, d3, d4]."""
    p = params(lam, octave, double)
    return np.concatenate([p, np.asarray(d, dtype=p.dtype)])


def run(pos, mat, scheme, params_v, materials, src_xyz, src_type, src_samples, rec_xyz, steps, n_parts=1,
        matidx_as_written=1, soft_accumulate=0, timed_from_step=0):
    """Full simulation on padded+translated volumes.  Returns (responses [n_rec][steps], seconds)."""
    double = params_v.dtype == np.float64
    dt = np.float64 if double else np.float32
    Z, Y, X = pos.shape
    materials = np.ascontiguousarray(materials, dtype=dt)
    src_xyz = np.ascontiguousarray(src_xyz, dtype=np.int32).reshape(-1, 3)
    src_type = np.ascontiguousarray(src_type, dtype=np.int32).reshape(-1)
    rec_xyz = np.ascontiguousarray(rec_xyz, dtype=np.int32).reshape(-1, 3)
    n_src, n_rec = src_xyz.shape[0], rec_xyz.shape[0]
    src_samples = np.ascontiguousarray(src_samples, dtype=dt).reshape(n_src, steps) if n_src else np.zeros((0, steps), dt)
    out = np.zeros((n_rec, steps), dtype=dt)
    fn = lib().pfo_run_f64 if double else lib().pfo_run_f32
    secs = fn(_p(pos), _p(mat), C.c_int64(X), C.c_int64(Y), C.c_int64(Z), C.c_int(scheme), _p(params_v), _p(materials),
              C.c_int(matidx_as_written), C.c_int(soft_accumulate), C.c_int(n_parts), C.c_int(n_src), _p(src_xyz),
              _p(src_type), _p(src_samples), C.c_int(n_rec), _p(rec_xyz), C.c_int64(steps), _p(out),
              C.c_int(timed_from_step))
    return out, secs


def step_slab(pos, mat, scheme, params_v, materials, cur, past, matidx_as_written=1):
    """One update launch over a slab [nz][Y][X] (planes 0 and nz-1 are halos): returns the next field."""
    double = params_v.dtype == np.float64
    dt = np.float64 if double else np.float32
    nz, Y, X = pos.shape
    materials = np.ascontiguousarray(materials, dtype=dt)
    cur = np.ascontiguousarray(cur, dtype=dt)
    new = np.array(past, dtype=dt, order="C", copy=True)
    fn = lib().pfo_step_slab_f64 if double else lib().pfo_step_slab_f32
    fn(_p(np.ascontiguousarray(pos)), _p(np.ascontiguousarray(mat)), C.c_int64(X), C.c_int64(Y), C.c_int64(nz), C.c_int(scheme),
       _p(params_v), _p(materials), C.c_int(matidx_as_written), _p(cur), _p(new))
    return new


def run_dif(pos, mat, scheme, params_v, materials, dif_order, src_xyz, src_type, src_samples, rec_xyz, steps, n_parts=1):
    """run() with digital-impedance-filter boundaries: materials rows are [b0..bN, a1..aN]."""
    double = params_v.dtype == np.float64
    dt = np.float64 if double else np.float32
    Z, Y, X = pos.shape
    materials = np.ascontiguousarray(materials, dtype=dt)
    assert materials.ndim == 2 and materials.shape[1] == 20 and 2 * dif_order + 1 <= 20
    src_xyz = np.ascontiguousarray(src_xyz, dtype=np.int32).reshape(-1, 3)
    src_type = np.ascontiguousarray(src_type, dtype=np.int32).reshape(-1)
    rec_xyz = np.ascontiguousarray(rec_xyz, dtype=np.int32).reshape(-1, 3)
    n_src, n_rec = src_xyz.shape[0], rec_xyz.shape[0]
    src_samples = np.ascontiguousarray(src_samples, dtype=dt).reshape(n_src, steps) if n_src else np.zeros((0, steps), dt)
    out = np.zeros((n_rec, steps), dtype=dt)
    L = lib()
    fn = L.pfo_run_dif_f64 if double else L.pfo_run_dif_f32
    fn.restype = C.c_double
    secs = fn(_p(pos), _p(mat), C.c_int64(X), C.c_int64(Y), C.c_int64(Z), C.c_int(scheme), _p(params_v), _p(materials),
              C.c_int(dif_order), C.c_int(n_parts), C.c_int(n_src), _p(src_xyz), _p(src_type), _p(src_samples), C.c_int(n_rec),
              _p(rec_xyz), C.c_int64(steps), _p(out))
    return out, secs


def source_samples(input_type, steps, fs=7000, data=None, double=False, transparent=False, grid_ir=None):
    """SimulationParameters::getSourceSample[Double] for steps 0..steps-1."""
    dt = np.float64 if double else np.float32
    data = np.ascontiguousarray(data if data is not None else [], dtype=dt)
    ir = np.ascontiguousarray(grid_ir if grid_ir is not None else [], dtype=np.float32)
    out = np.zeros(steps, dtype=dt)
    L = lib()
    for i in range(steps):
        if transparent:
            f = L.pfo_transparent_sample_f64 if double else L.pfo_transparent_sample_f32
            out[i] = f(C.c_int(input_type), C.c_uint(i), C.c_uint(fs), _p(data), C.c_uint(data.size), _p(ir), C.c_uint(ir.size))
        else:
            f = L.pfo_regular_sample_f64 if double else L.pfo_regular_sample_f32
            out[i] = f(C.c_int(input_type), C.c_uint(i), C.c_uint(fs), _p(data), C.c_uint(data.size))
    return out


def element_idx(p, fs=7000, c=344.0, lam=None, add_padding=True):
    lam = np.float32(1.0 / np.sqrt(3.0)) if lam is None else np.float32(lam)
    out = np.zeros(3, dtype=np.int32)
    lib().pfo_element_idx(C.c_float(p[0]), C.c_float(p[1]), C.c_float(p[2]), C.c_uint(fs), C.c_float(c), C.c_float(lam),
                          C.c_int(1 if add_padding else 0), _p(out))
    return tuple(int(v) for v in out)
