"""Pencil oracle -- bit-faithful NumPy restatement of the reference hot path.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function
cites the reference lines it restates (paths relative to /root/reference).
Unlike the reference it never builds a Kronecker matrix (numgrids/diff.py:145-171,
:222-239) or a meshgrid (numgrids/grids.py:27-28), so it runs at config sizes.

Pinning: ``tests/golden/make_golden.py`` (build container) checks every function
here with ``np.array_equal`` against the real reference classes imported from
/root/reference (findiff supplied by oracle/findiff_port.py -- FD parity
therefore "unpinned" w.r.t. the real third-party findiff) and stores the
resulting vectors under tests/golden/.
"""
from __future__ import annotations

import numpy as np

from oracle.findiff_port import FinDiff, coefficients  # noqa: F401  (re-export)

# --------------------------------------------------------------------------
# axis coordinates                                     numgrids/axes.py
# --------------------------------------------------------------------------

def equidistant_coords(n: int, low: float, high: float, periodic: bool = False):
    """numgrids/axes.py:238-240 (internal linspace) and :122 (linear map)."""
    ci = np.linspace(0, 1, n, endpoint=not periodic)
    return ci, ci * (high - low) + low


def chebyshev_coords(n: int, low: float, high: float):
    """numgrids/axes.py:318-324: descending cos nodes; reversed, mapped to [low, high]."""
    ci = np.cos(np.arange(n) * np.pi / (n - 1))
    co = (ci[::-1] + 1) / 2
    return ci, co * (high - low) + low


def log_coords(n: int, low: float, high: float):
    """numgrids/axes.py:368-374: linspace in ln x (internal), logspace (external)."""
    return (np.linspace(np.log(low), np.log(high), n),
            np.logspace(np.log10(low), np.log10(high), n))


# --------------------------------------------------------------------------
# finite differences                      numgrids/diff.py:82-91 -> findiff
# --------------------------------------------------------------------------

def fd_apply(f: np.ndarray, axis: int, h: float, order: int, acc: int = 4) -> np.ndarray:
    """``FiniteDifferenceDiff.operator(f)`` = ``FinDiff(axis, h, order, acc)(f)``."""
    return FinDiff(axis, h, order, acc=acc)(f)


def log_apply(f: np.ndarray, axis: int, h_internal: float, x: np.ndarray,
              order: int, acc: int = 6) -> np.ndarray:
    """``LogDiff.operator`` (numgrids/diff.py:258-262): FD on ln x, then ``/ x`` ONCE for any order."""
    shape = [1] * f.ndim
    shape[axis] = len(x)
    return FinDiff(axis, h_internal, order, acc=acc)(f) / np.asarray(x).reshape(shape)


# --------------------------------------------------------------------------
# FFT spectral derivative                          numgrids/diff.py:109-143
# --------------------------------------------------------------------------

def fft_multiplier(n: int, spacing: float, order: int) -> np.ndarray:
    """numgrids/diff.py:117-125: W = (i k 2pi/period)**order, Nyquist zeroed for even n / odd order."""
    period = n * spacing
    scale = 2 * np.pi / period
    k = np.fft.fftfreq(n) * n
    W = (1j * k * scale) ** order
    if n % 2 == 0 and order % 2 == 1:
        W[n // 2] = 0
    return W


def fft_apply(f: np.ndarray, axis: int, W: np.ndarray) -> np.ndarray:
    """numgrids/diff.py:133-143: real(ifft(W * fft(f, axis), axis))."""
    shape = [1] * f.ndim
    shape[axis] = len(W)
    F = np.fft.fft(f, axis=axis)
    dF = W.reshape(shape) * F
    return np.real(np.fft.ifft(dF, axis=axis))


# --------------------------------------------------------------------------
# Chebyshev                                          numgrids/diff.py:186-239
# --------------------------------------------------------------------------

def _seq_matmul(A: np.ndarray, B: np.ndarray, descending: bool = False) -> np.ndarray:
    """C = A B accumulated from 0.0 over k in one direction, mul and add rounded separately.

    This is the arithmetic of scipy's SMMP ``csr_matmat`` / BSR block gemm behind
    ``M *= M`` (numgrids/diff.py:190-191).
    """
    C = np.zeros((A.shape[0], B.shape[1]))
    ks = range(A.shape[1] - 1, -1, -1) if descending else range(A.shape[1])
    for k in ks:
        C += A[:, k:k + 1] * B[k:k + 1, :]
    return C


def chebyshev_raw_matrix(n: int, low: float, high: float):
    """numgrids/diff.py:204-222: Trefethen matrix on the descending nodes, negated; and ``scale``."""
    x, coords = chebyshev_coords(n, low, high)
    N = n - 1
    D = np.zeros((n, n))
    D[0, 0] = (2 * N ** 2 + 1) / 6
    D[N, N] = -D[0, 0]
    for j in range(1, N):
        D[j, j] = -0.5 * x[j] / (1 - x[j] ** 2)
    for i in range(n):
        ci = 2 if i in (0, N) else 1
        for j in range(n):
            if i != j:
                cj = 2 if j in (0, N) else 1
                D[i, j] = ci / cj * (-1) ** (i + j) / (x[i] - x[j])
    return -D, coords[-1] - coords[0]


def chebyshev_matrix(n: int, low: float, high: float, order: int,
                     last_axis_of_nd: bool = False) -> np.ndarray:
    """Dense 1-D matrix whose entries equal those of ``ChebyshevDiff._diff_matrix``.

    numgrids/diff.py:188-192: ``order-1`` SELF-products (so order 2 -> D^2, order
    3 -> D^4, order 4 -> D^8), then ``*(2/scale)**order``.  Summation direction
    of self-product t (t = 1..order-1), established against the real reference
    with ``np.array_equal`` (scipy 1.18.1):

    * last axis of an N-D grid (``kron(I, D)`` is BSR, block gemm): k ascending;
    * 1-D grid (CSC) or non-last axis (COO -> CSR, SMMP): SMMP emits each
      row/column in reverse insertion order, so the storage order flips after
      every product -> k ascending for odd t, descending for even t.
    """
    M, scale = chebyshev_raw_matrix(n, low, high)
    for t in range(1, order):
        M = _seq_matmul(M, M, descending=(not last_axis_of_nd) and t % 2 == 0)
    return M * (2 / scale) ** order


def chebyshev_descending(order: int, axis: int, ndim: int) -> bool:
    """Summation direction of the reference's sparse matvec (extended to order >= 3).

    ``kron`` yields COO (non-last axis) or BSR (last axis), both summing j
    ascending for order 1.  Each SMMP self-product of the CSR matrix reverses
    the per-row column order, so on a NON-last axis of an N-D grid the matvec
    sums j descending for even order and ascending for odd order.  Last axis
    (BSR) and 1-D (CSC matvec walks columns in order) are always ascending.
    """
    return ndim > 1 and axis != ndim - 1 and order % 2 == 0


def chebyshev_apply(f: np.ndarray, axis: int, M: np.ndarray, descending: bool) -> np.ndarray:
    """numgrids/diff.py:194-197 as a per-pencil sequential sum y_i = sum_j M_ij f_j.

    Accumulates from 0.0, one j at a time (ascending or descending), with the
    product and the sum rounded separately -- the order scipy's
    coo/csr/bsr/csc_matvec use on the kron'd matrix.
    """
    n = M.shape[0]
    F = np.moveaxis(f, axis, 0)
    shp = F.shape
    F = np.ascontiguousarray(F).reshape(n, -1)
    Y = np.zeros_like(F)
    js = range(n - 1, -1, -1) if descending else range(n)
    for j in js:
        Y += M[:, j:j + 1] * F[j:j + 1, :]
    return np.ascontiguousarray(np.moveaxis(Y.reshape(shp), 0, axis))


# --------------------------------------------------------------------------
# curvilinear composites                     numgrids/curvilinear.py:126-317
# --------------------------------------------------------------------------

def jacobian(h):
    """numgrids/curvilinear.py:133-140: J = ((1*h0)*h1)*..."""
    J = np.ones(np.broadcast_shapes(*[np.shape(hi) for hi in h]))
    for hi in h:
        J = J * hi
    return J


def _finite(a):
    return np.where(np.isfinite(a), a, 0.0)


def gradient(f, h, D):
    """numgrids/curvilinear.py:169-177.  ``D[i]`` is the first-derivative apply along axis i."""
    out = []
    for i in range(len(D)):
        with np.errstate(divide="ignore", invalid="ignore"):
            comp = (1.0 / h[i]) * D[i](f)
        out.append(_finite(comp))
    return tuple(out)


def divergence(v, h, D):
    """numgrids/curvilinear.py:206-215."""
    J = jacobian(h)
    result = np.zeros(np.shape(v[0]))
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(len(D)):
            coeff = J / h[i]
            result = result + D[i](coeff * v[i])
        result = result / J
    return _finite(result)


def laplacian(f, h, D):
    """numgrids/curvilinear.py:235-244: (1/J) sum_i D_i[(J/h_i^2) D_i f]."""
    J = jacobian(h)
    result = np.zeros(np.shape(f))
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(len(D)):
            inner = (J / h[i] ** 2) * D[i](f)
            result = result + D[i](inner)
        result = result / J
    return _finite(result)


def curl(v, h, D):
    """numgrids/curvilinear.py:289-317 (2-D scalar and 3-D vector curl)."""
    nd = len(D)
    if nd == 2:
        with np.errstate(divide="ignore", invalid="ignore"):
            res = (1.0 / (h[0] * h[1])) * (D[0](h[1] * v[1]) - D[1](h[0] * v[0]))
        return _finite(res)
    if nd != 3:
        raise ValueError("Curl is only defined for 2D and 3D grids.")
    out = []
    for i, j, k in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
        with np.errstate(divide="ignore", invalid="ignore"):
            comp = (1.0 / (h[j] * h[k])) * (D[j](h[k] * v[k]) - D[k](h[j] * v[j]))
        out.append(_finite(comp))
    return tuple(out)


def cartesian_laplacian(f, hs, acc: int = 4):
    """Reference idiom Diff(g,2,0)(f)+Diff(g,2,1)(f)+... (tests/test_diff.py:35-37)."""
    res = None
    for a, h in enumerate(hs):
        d = fd_apply(f, a, h, 2, acc)
        res = d if res is None else res + d
    return res


# --------------------------------------------------------------------------
# axis-spec driven operator factory (mirrors Axis.create_diff_operator,
# numgrids/axes.py:242-251, :326-333, :376-379)
# --------------------------------------------------------------------------

def make_axis(kind: str, n: int, low: float, high: float) -> dict:
    kind = kind.lower()
    if kind == "equidistant":
        ci, c = equidistant_coords(n, low, high, False)
    elif kind in ("periodic", "equidistant_periodic"):
        kind = "periodic"
        ci, c = equidistant_coords(n, low, high, True)
    elif kind == "chebyshev":
        ci, c = chebyshev_coords(n, low, high)
    elif kind in ("log", "logarithmic"):
        kind = "log"
        ci, c = log_coords(n, low, high)
    else:
        raise NotImplementedError(kind)
    return {"kind": kind, "n": n, "low": low, "high": high, "coords": c, "internal": ci}


def make_diff(axes: list[dict], order: int, axis: int, acc: int = 4):
    """Return ``f -> d^order f / dx_axis^order`` following the reference's dispatch."""
    ax = axes[axis]
    nd = len(axes)
    if ax["kind"] == "equidistant":
        h = ax["coords"][1] - ax["coords"][0]
        return lambda f: fd_apply(f, axis, h, order, acc)
    if ax["kind"] == "periodic":
        W = fft_multiplier(ax["n"], ax["coords"][1] - ax["coords"][0], order)
        return lambda f: fft_apply(f, axis, W)
    if ax["kind"] == "chebyshev":
        M = chebyshev_matrix(ax["n"], ax["low"], ax["high"], order,
                             last_axis_of_nd=(nd > 1 and axis == nd - 1))
        desc = chebyshev_descending(order, axis, nd)
        return lambda f: chebyshev_apply(f, axis, M, desc)
    if ax["kind"] == "log":
        hi = ax["internal"][1] - ax["internal"][0]
        return lambda f: log_apply(f, axis, hi, ax["coords"], order, acc)
    raise NotImplementedError(ax["kind"])
