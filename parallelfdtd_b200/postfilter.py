"""Post-filters for simulated responses and the JSON geometry format of the reference's front ends (SURVEY f-3/f-4).

FDTDfilter       -- reference python/FDTDfilter.py:13-19: 200-tap linear-phase FIR low-pass, Dolph-Chebyshev window
                    with 70 dB ripple, cut-off `normalized_cutoff * sfs`, applied causally along axis 0.
FDTDpostFilter   -- reference matlab/functions/FDTDpostFilter.m:1-30: 201-tap Hamming-window low-pass followed by a
                    one-pole / one-zero DC blocker with a 5 Hz corner, responses as rows in, columns out.
load_json_geometry -- the geometry files of python/testBench.py:44-69 / matlab/testBench.m (`vertices`, `indices`,
                    `layers_of_triangles`, `layer_names`).

These run on the host on nRec x nSteps responses (kilobytes); they are not part of the device path.
"""
from __future__ import annotations

import json

import numpy as np


def _windowed_sinc_lowpass(n_taps: int, cutoff: float, window: np.ndarray) -> np.ndarray:
    """Type-I/II linear-phase low-pass by the window method; `cutoff` is relative to Nyquist (0..1); unit DC gain."""
    if not 0.0 < cutoff < 1.0:
        raise ValueError("cut-off must lie strictly between 0 and the Nyquist frequency")
    m = np.arange(n_taps, dtype=np.float64) - (n_taps - 1) / 2.0
    h = cutoff * np.sinc(cutoff * m) * window
    return h / h.sum()


def _fir_apply(taps: np.ndarray, x: np.ndarray, axis: int) -> np.ndarray:
    """Causal FIR filtering (zero initial state) along `axis`, output as long as the input."""
    x = np.asarray(x, dtype=np.float64)
    moved = np.moveaxis(x, axis, -1)
    flat = moved.reshape(-1, moved.shape[-1])
    out = np.empty_like(flat)
    for i, row in enumerate(flat):
        out[i] = np.convolve(row, taps)[:row.size]
    return np.moveaxis(out.reshape(moved.shape), -1, axis)


def FDTDfilter(x, sfs, fs, normalized_cutoff):
    """Low-pass a response (or a matrix of responses, time along axis 0) simulated at sampling rate `sfs`.
    `fs` is accepted and unused, as in the reference (python/FDTDfilter.py:13)."""
    from scipy.signal.windows import chebwin
    taps = _windowed_sinc_lowpass(200, (sfs * normalized_cutoff) / (sfs / 2.0), chebwin(200, 70.0))
    return _fir_apply(taps, x, 0)


def dcblock_pole(cutoff_hz: float, fs: float) -> float:
    """Pole a of the DC blocker H(z) = (1 - z^-1) / (1 - a z^-1) for a cut-on frequency `cutoff_hz`: the closed form
    of J. de Freitas, "The DC Blocking Filter" (2007), which the reference obtains from the MATLAB File Exchange
    function `dcblock(fc, fs)` (matlab/functions/FileExchange/dcblock.m:188-190, third party):
    a = (sqrt 3 - 2 sin(pi Fc)) / (sin(pi Fc) + sqrt 3 cos(pi Fc)), Fc = fc / (fs / 2), valid for 0 <= Fc <= 1/3."""
    fc = 2.0 * cutoff_hz / fs
    if not 0.0 <= fc <= 1.0 / 3.0:
        raise ValueError("cut-on frequency must lie between 0 and fs/6")
    s3 = np.sqrt(3.0)
    return float((s3 - 2.0 * np.sin(np.pi * fc)) / (np.sin(np.pi * fc) + s3 * np.cos(np.pi * fc)))


def FDTDpostFilter(ir, fs, frac):
    """Low-pass at `frac` (relative to Nyquist) + 5 Hz DC blocker; `ir` holds one response per row (transposed
    first if it has more rows than columns) and the result one response per column (FDTDpostFilter.m:15-28)."""
    ir = np.atleast_2d(np.asarray(ir, dtype=np.float64))
    if ir.shape[0] > ir.shape[1]:
        ir = ir.T
    y = _fir_apply(_windowed_sinc_lowpass(201, frac, np.hamming(201)), ir, 1)
    p = dcblock_pole(5.0, fs)
    out = np.empty_like(y)
    for r in range(y.shape[0]):
        prev_x = 0.0
        prev_y = 0.0
        row, o = y[r], out[r]
        for n in range(row.size):
            prev_y = row[n] - prev_x + p * prev_y
            prev_x = row[n]
            o[n] = prev_y
    return out.T


def load_json_geometry(path):
    """-> (vertices [nV][3] float32, indices [nT][3] uint32, layers {name: [triangle indices]})"""
    with open(path) as f:
        m = json.load(f)
    v = np.asarray(m["vertices"], dtype=np.float32).reshape(-1, 3)
    t = np.asarray(m["indices"], dtype=np.uint32).reshape(-1, 3)
    if t.size and int(t.max()) >= len(v):
        raise ValueError(f"{path}: triangle index {int(t.max())} beyond {len(v)} vertices")
    names = m.get("layer_names", [])
    of_tri = m.get("layers_of_triangles", [])
    if of_tri and len(of_tri) != len(t):
        raise ValueError(f"{path}: {len(of_tri)} layer entries for {len(t)} triangles")
    layers = {name: [i for i, l in enumerate(of_tri) if l == name] for name in names}
    return v, t, layers
