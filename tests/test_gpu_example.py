"""GPU: examples/test_bench.py -- the reference's python/testBench.py workflow (JSON geometry, per-layer materials,
libPyFDTD.App, post-filter) end to end, in every mode the script offers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_example_runs_in_every_mode(capi, gpu):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import test_bench as tb
    base = tb.run(steps=200, fs=10000, verbose=False)
    r = base["responses"]
    assert r.shape == (200, 2) and np.isfinite(r).all() and np.abs(r).max() > 0
    assert base["filtered"].shape == r.shape and np.abs(base["filtered"]).max() > 0
    d = tb.run(double=True, steps=200, fs=10000, verbose=False)["responses"]
    assert d.dtype == np.float64 and np.linalg.norm(d - r) / np.linalg.norm(d) < 1e-4
    c = tb.run(captures=True, steps=200, fs=10000, verbose=False)
    assert np.array_equal(c["responses"], r) and len(c["slices"]) == 1 and np.abs(c["slices"][0]).max() > 0
    f = tb.run(filters=True, steps=200, fs=10000, verbose=False)["responses"]
    assert np.isfinite(f).all() and not np.array_equal(f, r)
    for scheme in (2, 3, 4):
        s = tb.run(scheme=scheme, steps=200, fs=10000, verbose=False)["responses"]
        assert np.isfinite(s).all() and np.abs(s).max() > 0
