"""NumPy restatement of the reference's ``Integral`` -- TEST INFRASTRUCTURE ONLY.

Follows reference numgrids/integration.py: per axis (``__init__`` :33-40) the first-derivative matrix
of the 1-D grid with its first row and column removed is inverted with ``scipy.sparse.linalg.inv``;
``__call__`` (:58-72) walks the axes in order -- a periodic equidistant axis contributes
``spacing * np.sum(f, axis=0)``; any other axis drops the first layer, multiplies by the
Kronecker-lifted inverse and keeps the LAST row, i.e. ``sum_j Dinv[-1, j-1] * f[j]`` for j = 1..n-1,
accumulated in ascending j from 0.0 (the order scipy's sparse matvec uses; probe-verified bit for
bit against the real reference by tests/golden/make_golden_int.py).
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import csc_matrix
from scipy.sparse.linalg import inv


def axis_weights(D_1d, periodic: bool, spacing: float | None, n: int) -> np.ndarray:
    """Weight vector of one axis; ``D_1d`` is the 1-D first-derivative matrix (sparse), unused if periodic."""
    if periodic:
        return np.full(n, spacing)
    Dinv = inv(csc_matrix(D_1d[1:, 1:]))
    w = np.zeros(n)
    w[1:] = np.asarray(Dinv.toarray())[-1]
    return w


def integral(f: np.ndarray, weights: list[np.ndarray], periodic: list[bool]) -> np.float64:
    f_ = f
    for w, per in zip(weights, periodic):
        if per:
            f_ = w[0] * np.sum(f_, axis=0)
        else:
            acc = np.zeros(f_.shape[1:])
            for j in range(1, len(w)):
                acc = acc + w[j] * f_[j]
            f_ = acc
    return np.float64(f_)
