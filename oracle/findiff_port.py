"""Restatement of ``findiff.FinDiff`` on a uniform grid -- ORACLE / TEST INFRASTRUCTURE ONLY.

The reference calls the third-party package ``findiff`` (pinned only as
``findiff>=0.10``, reference pyproject.toml:23) at numgrids/diff.py:10, :88
(``FinDiff(axis_index, spacing, order, acc=acc)``), :91 (``.matrix(shape)``),
:259 and :273.  The package is not vendored under /root/reference and cannot be
installed offline, so its published algorithm is restated here:

coefficients (findiff ``coefficients(deriv, acc)`` / ``calc_coefs``)
    central offsets ``-p..p`` with ``p = (2*floor((deriv+1)/2) - 1 + acc)//2``;
    forward offsets ``0..m-1`` and backward offsets ``0,-1,..,-(m-1)`` with
    ``m = 2p+1`` (+1 when ``deriv`` is even); weights solve the Taylor system
    ``A[i,k] = offset_k**i``, ``b[deriv] = deriv!`` with ``numpy.linalg.solve``.
apply (findiff ``Diff.diff`` / ``_apply_to_array``)
    ``yd = zeros_like(y)``; the central scheme fills ``[nb, n-nb)``, forward
    ``[0, nb)``, backward ``[n-nb, n)`` with ``nb = len(center)//2``; for each
    tap in stored order ``yd[ref] += w*y[ref+off]`` (plain ``+= y[...]`` when
    ``|1-w| < 1e-14``); finally ``yd * (1/h**deriv)``.

PARITY NOTE: because real findiff is unavailable, FD parity is pinned to this
restatement (it passes the reference's own test-suite, see
tests/golden/make_golden.py), i.e. "parity unpinned" with respect to the real
third-party package.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

__version__ = "0.10.0-port"


def _solve_taylor(deriv: int, offsets: list[int]) -> dict:
    m = len(offsets)
    A = np.array([[1.0 for _ in offsets]] +
                 [[float(o) ** i for o in offsets] for i in range(1, m)], dtype="float")
    b = np.zeros(m, dtype="float")
    b[deriv] = math.factorial(deriv)
    return {"coefficients": np.linalg.solve(A, b), "offsets": list(offsets)}


def coefficients(deriv: int, acc: int) -> dict:
    """Weights/offsets of the three schemes, like ``findiff.coefficients(deriv, acc)``."""
    if acc % 2 == 1 or acc <= 0:
        raise ValueError("Accuracy order acc must be positive EVEN integer")
    if deriv < 0:
        raise ValueError("Derive degree must be positive integer")
    n_central = 2 * math.floor((deriv + 1) / 2) - 1 + acc
    p = n_central // 2
    m = n_central + 1 if deriv % 2 == 0 else n_central
    return {
        "center": _solve_taylor(deriv, list(range(-p, p + 1))),
        "forward": _solve_taylor(deriv, list(range(m))),
        "backward": _solve_taylor(deriv, [-o for o in range(m)]),
    }


class FinDiff:
    """``FinDiff(axis, h, order, acc=2)`` -- uniform-grid partial derivative."""

    def __init__(self, axis: int, h: float, order: int = 1, acc: int = 2) -> None:
        self.axis = int(axis)
        self.h = h
        self.order = int(order)
        self.acc = int(acc)
        self.coefs = coefficients(self.order, self.acc)

    # -- array apply ---------------------------------------------------
    def __call__(self, y: np.ndarray) -> np.ndarray:
        y = np.asarray(y)
        dim = self.axis
        n = y.shape[dim]
        nb = len(self.coefs["center"]["coefficients"]) // 2
        yd = np.zeros_like(y)
        regions = (("center", nb, n - nb), ("forward", 0, nb), ("backward", n - nb, n))
        for scheme, lo, hi in regions:
            w = self.coefs[scheme]["coefficients"]
            offs = self.coefs[scheme]["offsets"]
            dst = [slice(None)] * y.ndim
            dst[dim] = slice(lo, hi, 1)
            dst = tuple(dst)
            for wk, ok in zip(w, offs):
                src = [slice(None)] * y.ndim
                src[dim] = slice(lo + ok, hi + ok, 1)
                src = tuple(src)
                if abs(1 - wk) < 1.0e-14:
                    yd[dst] += y[src]
                else:
                    yd[dst] += wk * y[src]
        h_inv = 1.0 / self.h ** self.order
        return yd * h_inv

    # -- sparse matrix -------------------------------------------------
    def matrix(self, shape) -> sp.spmatrix:
        shape = tuple(int(s) for s in shape)
        n = shape[self.axis]
        nb = len(self.coefs["center"]["coefficients"]) // 2
        rows, cols, vals = [], [], []
        for i in range(n):
            scheme = "forward" if i < nb else ("backward" if i >= n - nb else "center")
            for wk, ok in zip(self.coefs[scheme]["coefficients"], self.coefs[scheme]["offsets"]):
                rows.append(i)
                cols.append(i + ok)
                vals.append(wk / self.h ** self.order)
        D1 = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
        before = int(np.prod(shape[: self.axis], dtype=np.int64))
        after = int(np.prod(shape[self.axis + 1:], dtype=np.int64))
        return sp.csr_matrix(sp.kron(sp.kron(sp.identity(before), D1), sp.identity(after)))
