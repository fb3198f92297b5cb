"""GPU: the library's cache of device blocks (pfdtd_internal.h dev_alloc / dev_free).  A solver built out of blocks that
an earlier solver left behind gives the same answers as one built out of fresh memory; the cached blocks count as free
memory, go back to the driver on request, and PFDTD_CACHE_MB=0 turns the whole thing off."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle
from parallelfdtd_b200 import synth
from tests import fdtd_cases as fc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIMS = (256, 128, 96)          # 3.1e6 voxels: node volumes and fields are above the cache's 1 MiB floor


def _free_mb(capi, raw=False):
    if raw:
        import torch
        return torch.cuda.mem_get_info(0)[0] >> 20
    tot, free = C.c_int(0), C.c_int(0)
    assert capi.lib().pfdtd_device_mem_mb(0, C.byref(tot), C.byref(free)) == 0
    return free.value


def _job(capi, dif, steps=40, dirty=None, double=False, n_parts=1):
    bid, mat = synth.shoebox(DIMS, 4)
    npdt = np.float64 if double else np.float32
    refl = list(np.linspace(0.95, 0.6, 4))
    tab = (synth.filter_material_table(refl, dif) if dif else synth.material_table(refl)).astype(npdt)
    s = capi.Solver()
    s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
    s.set_option(capi.OPT_DIF_ORDER, dif)
    s.setup_mesh(bid, mat, (32, 4, 1), 0, capi.F64 if double else capi.F32, oracle.params(fc.LAM, 0, double), tab)
    s.make_partition(n_parts, [0] * n_parts)
    src = oracle.source_samples(1, steps, double=double)[None, :]
    s.set_sources([[3, 3, 3]], [capi.SRC_HARD], src)          # next to three walls: the filter states matter from the start
    s.set_receivers([[5, 4, 6], [200, 100, 80], [2, 2, 2]])
    r, _ = s.run(steps)
    if dirty is not None:                                     # leave something other than zeros in every block
        for _ in range(3):
            s.set_sample(dirty[0], dirty[1], dirty[2], 1e3)
            s.run(8)
    s.close()
    return r


@pytest.mark.parametrize("dif,double,n_parts", [(0, False, 1), (2, False, 1), (2, True, 1), (2, False, 3)])
def test_recycled_blocks_give_the_answers_of_fresh_memory(capi, gpu, dif, double, n_parts):
    capi.release_cached_memory(-1)
    fresh = _job(capi, dif, double=double, n_parts=n_parts, dirty=(100, 60, 40))
    assert np.abs(fresh).max() > 0
    assert _free_mb(capi) - _free_mb(capi, raw=True) >= 20    # the dead solver's blocks are kept ...
    for _ in range(2):
        again = _job(capi, dif, double=double, n_parts=n_parts, dirty=(30, 90, 70))
        assert np.array_equal(again, fresh)                   # ... and the next solver is built out of them
    held = _free_mb(capi) - _free_mb(capi, raw=True)
    capi.release_cached_memory(0)
    assert _free_mb(capi) - _free_mb(capi, raw=True) <= 4 and held >= 20


def test_cached_blocks_count_as_free_memory_and_repeats_do_not_grow(capi, gpu):
    _job(capi, 2)                                             # kernels loaded, streams and events seen once
    capi.release_cached_memory(-1)
    before = _free_mb(capi)
    _job(capi, 2)
    mid_raw = _free_mb(capi, raw=True)
    assert abs(_free_mb(capi) - before) <= 64                 # what pfdtd_device_mem_mb reports does not shrink
    for _ in range(3):
        _job(capi, 2)
    assert abs(_free_mb(capi, raw=True) - mid_raw) <= 64      # the same blocks every time
    capi.release_cached_memory(-1)
    assert abs(_free_mb(capi, raw=True) - before) <= 64


def test_device_alloc_free_go_through_the_cache_and_adopted_foreign_pointers_do_not(capi, gpu):
    import torch
    capi.release_cached_memory(-1)
    lib = capi.lib()
    p = C.c_void_p()
    assert lib.pfdtd_device_alloc(0, C.c_size_t(8 << 20), C.byref(p)) == 0
    one = np.full(8 << 20, 7, dtype=np.uint8)
    assert lib.pfdtd_device_upload(0, p, one.ctypes.data_as(C.c_void_p), C.c_size_t(one.size)) == 0
    first = p.value
    assert lib.pfdtd_device_free(0, p) == 0
    q = C.c_void_p()
    assert lib.pfdtd_device_alloc(0, C.c_size_t(8 << 20), C.byref(q)) == 0
    assert q.value == first                                   # the same block, handed out zero-filled
    back = np.empty(8 << 20, dtype=np.uint8)
    assert lib.pfdtd_device_download(0, back.ctypes.data_as(C.c_void_p), q, C.c_size_t(back.size)) == 0
    assert not back.any()
    assert lib.pfdtd_device_free(0, q) == 0
    small = C.c_void_p()
    assert lib.pfdtd_device_alloc(0, C.c_size_t(4096), C.byref(small)) == 0 and lib.pfdtd_device_free(0, small) == 0
    capi.release_cached_memory(-1)
    torch.cuda.synchronize()


def test_cache_can_be_switched_off(gpu):
    code = ("import ctypes as C, numpy as np, torch\n"
            "from parallelfdtd_b200 import capi\n"
            "lib = capi.lib(); p = C.c_void_p()\n"
            "torch.cuda.init(); torch.cuda.synchronize()\n"
            "assert lib.pfdtd_device_alloc(0, C.c_size_t(64 << 20), C.byref(p)) == 0\n"
            "a = torch.cuda.mem_get_info(0)[0]\n"
            "assert lib.pfdtd_device_free(0, p) == 0\n"
            "b = torch.cuda.mem_get_info(0)[0]\n"
            "print((b - a) >> 20)\n")
    for mb, lo, hi in (("0", 56, 72), ("1024", -8, 8)):
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env={**os.environ, "PFDTD_CACHE_MB": mb}, capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        assert lo <= int(r.stdout.strip().splitlines()[-1]) <= hi, (mb, r.stdout)
