"""CPU: the response post-filters and the JSON geometry loader (SURVEY f-3 / f-4) against golden vectors made by
the reference's own python/FDTDfilter.py (tools/make_postfilter_golden.py) and against scipy / closed forms."""
import json
import os

import numpy as np
import pytest

from parallelfdtd_b200 import postfilter as pfl

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postfilter_fdtdfilter.npz")


def test_fdtdfilter_matches_the_reference_golden_vectors():
    g = np.load(GOLDEN)
    for k in range(3):
        sfs, cut = g[f"arg{k}"]
        y = pfl.FDTDfilter(g[f"x{k}"], sfs, 0, cut)
        assert y.shape == g[f"y{k}"].shape
        assert np.max(np.abs(y - g[f"y{k}"])) <= 1e-12 * max(1.0, np.max(np.abs(g[f"y{k}"])))


def test_fdtdfilter_is_a_unit_gain_lowpass():
    n = np.arange(4000)
    sfs = 100000.0
    lo = pfl.FDTDfilter(np.sin(2 * np.pi * 2000.0 / sfs * n), sfs, 0, 0.2)      # pass band (cut-off 20 kHz)
    hi = pfl.FDTDfilter(np.sin(2 * np.pi * 30000.0 / sfs * n), sfs, 0, 0.2)     # stop band
    assert abs(np.max(np.abs(lo[1000:])) - 1.0) < 1e-3
    assert np.max(np.abs(hi[1000:])) < 10 ** (-60 / 20)
    with pytest.raises(ValueError):
        pfl.FDTDfilter(np.zeros(10), sfs, 0, 0.5)                                # cut-off at Nyquist


def test_postfilter_matches_scipy_restatement_of_the_matlab_function():
    from scipy.signal import firwin, lfilter
    rng = np.random.default_rng(7)
    ir = rng.standard_normal((2, 1500)) + 0.3                                     # two responses as rows, DC offset
    fs, frac = 44100.0, 0.4
    y = pfl.FDTDpostFilter(ir, fs, frac)
    assert y.shape == (1500, 2)                                                   # columns out (FDTDpostFilter.m:27)
    b = firwin(201, frac)                                                         # fir1(200, frac): Hamming, unit DC gain
    a = pfl.dcblock_pole(5.0, fs)
    ref = lfilter([1.0, -1.0], [1.0, -a], lfilter(b, 1.0, ir, axis=1), axis=1).T
    assert np.max(np.abs(y - ref)) < 1e-10
    assert np.array_equal(pfl.FDTDpostFilter(ir.T, fs, frac), y)                  # column input is transposed first
    assert abs(y[-400:].mean()) < 0.3                                            # the offset is being removed (5 Hz corner: slowly)


def test_dcblock_pole_closed_form():
    assert pfl.dcblock_pole(0.0, 48000.0) == pytest.approx(1.0)
    a = pfl.dcblock_pole(5.0, 44100.0)
    assert 0.998 < a < 1.0
    # at the cut-on frequency the normalised response (1+a)/2 |H| is one half (the formula's definition of cut-on)
    w = 2 * np.pi * 5.0 / 44100.0
    h = (1 + a) / 2 * abs(1 - np.exp(-1j * w)) / abs(1 - a * np.exp(-1j * w))
    assert h == pytest.approx(0.5, rel=1e-9)
    with pytest.raises(ValueError):
        pfl.dcblock_pole(10000.0, 44100.0)


def test_json_geometry_loader(tmp_path):
    box = {"vertices": [1, 0, 0, 0, 0, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1],
           "indices": [3, 1, 2, 1, 3, 0, 5, 7, 4, 7, 5, 6, 5, 1, 0, 1, 5, 4, 7, 1, 4, 1, 7, 2, 7, 3, 2, 3, 7, 6, 3, 5, 0, 5, 3, 6],
           "layers_of_triangles": ["floor"] * 2 + ["ceiling"] * 2 + ["walls"] * 8,
           "layer_names": ["floor", "ceiling", "walls"]}
    p = tmp_path / "box.json"
    p.write_text(json.dumps(box))
    v, t, layers = pfl.load_json_geometry(str(p))
    assert v.shape == (8, 3) and v.dtype == np.float32 and t.shape == (12, 3) and t.dtype == np.uint32
    assert layers == {"floor": [0, 1], "ceiling": [2, 3], "walls": list(range(4, 12))}
    box["indices"][0] = 8
    p.write_text(json.dumps(box))
    with pytest.raises(ValueError):
        pfl.load_json_geometry(str(p))
