"""CPU: the synthetic-room generators (test / bench inputs).  The separable fast shoebox must equal the general
predicate path byte for byte, for whole volumes and z-ranges (every z-slab owner builds only its own part)."""
import numpy as np
import pytest

from parallelfdtd_b200 import synth


@pytest.mark.parametrize("dims,n_mat,zr,shell", [((64, 64, 64), 1, None, 1), ((48, 40, 49), 6, None, 1), ((48, 40, 49), 6, (10, 30), 1),
                                                 ((33, 17, 29), 4, (0, 5), 1), ((20, 22, 24), 6, (20, 24), 2), ((16, 12, 9), 20, None, 1),
                                                 ((5, 5, 5), 6, None, 1), ((4, 8, 8), 6, None, 1)])
def test_fast_shoebox_equals_the_general_path(dims, n_mat, zr, shell):
    z0, z1 = zr if zr else (0, None)
    bid, mat = synth.shoebox(dims, n_mat, z0, z1, shell)
    gbid, gmat = synth.shoebox_generic(dims, n_mat, z0, z1, shell)
    assert np.array_equal(bid, gbid) and np.array_equal(mat, gmat)
    assert bid.dtype == np.uint8 and mat.dtype == np.uint8 and bid.flags.c_contiguous


def test_slab_ranges_tile_the_whole_volume():
    dims = (40, 36, 50)
    for gen in (synth.shoebox, synth.hall, synth.banded_shoebox):
        whole = gen(dims, 5)
        parts = [gen(dims, 5, a, b) for a, b in ((0, 17), (17, 18), (18, 50))]
        for i in range(2):
            assert np.array_equal(np.concatenate([p[i] for p in parts]), whole[i]), gen.__name__
