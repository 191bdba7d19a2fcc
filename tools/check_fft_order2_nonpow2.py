"""Order-2 periodic derivative at lengths that are not powers of two: error against numpy.fft of (a) the dense ordered
contraction (the default for orders >= 2) and (b) the zero-padded transform, and of numpy.fft itself against a
long-double DFT (what 'exact' is).  Field: exp(sin x) * cos(2 x)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
import importlib
from numgrids_b200 import _tables  # noqa: E402
D = importlib.import_module("numgrids_b200.diff")

A = ng.AxisType
for n in (96, 100, 360, 500, 1000, 1536, 3000):
    ax = ng.create_axis(A.EQUIDISTANT_PERIODIC, n, 0.0, 2 * np.pi)
    g = ng.Grid(ng.create_axis(A.EQUIDISTANT, 64, 0.0, 1.0), ax)
    X, P = g.meshed_coords
    f = np.exp(np.sin(P)) * np.cos(2 * P) * (1 + X)
    W = _tables.fft_multiplier(n, ax.spacing, 2)
    ref = np.real(np.fft.ifft(W * np.fft.fft(f, axis=1), axis=1))
    # long-double DFT of one pencil: the exact answer to ~1e-19
    k = np.arange(n, dtype=np.longdouble)
    E = np.exp(-2j * np.pi * np.outer(k, k).astype(np.longdouble) / n)
    Fk = E @ f[3].astype(np.longdouble)
    exact = np.real((np.conj(E) @ (W.astype(np.clongdouble) * Fk)) / n).astype(np.float64)
    den = np.max(np.abs(exact))
    e_numpy = np.max(np.abs(ref[3] - exact)) / den
    res = {}
    for name, order1_only in (("dense", True), ("padded", False)):
        D.FFTDiff.PADDED_ORDERS = (1,) if order1_only else (1, 2)
        out = ng.Diff(g, 2, 1)(torch.from_numpy(f).cuda()).cpu().numpy()
        res[name] = (np.max(np.abs(out - ref)) / np.max(np.abs(ref)), np.max(np.abs(out[3] - exact)) / den)
    print(f"n={n:5d}: numpy vs exact {e_numpy:.2e} | dense vs numpy {res['dense'][0]:.2e} (vs exact {res['dense'][1]:.2e}) | "
          f"padded vs numpy {res['padded'][0]:.2e} (vs exact {res['padded'][1]:.2e})", flush=True)
