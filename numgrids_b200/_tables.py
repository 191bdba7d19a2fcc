"""Host-side coefficient tables handed to the CUDA library (product code; never imports oracle/).

The kernels reproduce the reference bit for bit only if they are fed the reference's own numbers,
so every table is produced by the same floating-point recipe the reference (or the third-party
code it calls) uses, in this process:

* finite-difference weights: the Taylor system solved with ``numpy.linalg.solve`` -- findiff's
  ``coefficients(deriv, acc)`` (called at reference numgrids/diff.py:88, :259).  If the real
  ``findiff`` package is importable its own ``coefficients`` is used instead.
* Chebyshev matrices: reference numgrids/diff.py:188-192, :204-222 including the self-product
  quirk (order k -> D**(2**(k-1))) and scipy's summation directions (see ``chebyshev_matrix``).
* FFT multipliers: reference numgrids/diff.py:117-125.
"""
from __future__ import annotations

import math

import numpy as np

# ------------------------------------------------------------------ finite differences


def _taylor_weights(deriv: int, offsets) -> np.ndarray:
    m = len(offsets)
    rows = [[1.0] * m] + [[float(o) ** i for o in offsets] for i in range(1, m)]
    rhs = np.zeros(m)
    rhs[deriv] = math.factorial(deriv)
    return np.linalg.solve(np.array(rows, dtype="float"), rhs)


def fd_schemes(order: int, acc: int) -> dict:
    """Offsets and weights of the central / forward / backward schemes (findiff layout)."""
    if not (isinstance(acc, (int, np.integer)) and acc > 0 and acc % 2 == 0):
        raise ValueError("Accuracy order acc must be positive EVEN integer")
    try:                                     # prefer the real third-party package when present
        import findiff  # type: ignore
        if not str(getattr(findiff, "__version__", "")).endswith("-port"):
            c = findiff.coefficients(order, acc)
            return {k: (list(c[k]["offsets"]), np.asarray(c[k]["coefficients"], dtype=float))
                    for k in ("center", "forward", "backward")}
    except ImportError:
        pass
    n_central = 2 * ((order + 1) // 2) - 1 + acc
    p = n_central // 2
    m = n_central + (1 if order % 2 == 0 else 0)
    sets = {"center": list(range(-p, p + 1)), "forward": list(range(m)),
            "backward": [-k for k in range(m)]}
    return {k: (offs, _taylor_weights(order, offs)) for k, offs in sets.items()}


def snap_unit_weights(w: np.ndarray) -> np.ndarray:
    """findiff adds ``y`` instead of ``w*y`` when ``|1-w| < 1e-14``: identical to an exact 1.0 weight."""
    w = np.array(w, dtype=float, copy=True)
    w[np.abs(1 - w) < 1.0e-14] = 1.0
    return w


def fd_min_points(order: int, acc: int) -> int:
    """Smallest axis length for which every scheme stays inside the axis."""
    s = fd_schemes(order, acc)
    nb = len(s["center"][0]) // 2
    return max(len(s["center"][0]), nb - 1 + len(s["forward"][0]), 2 * nb)


# ------------------------------------------------------------------ Chebyshev


def _ordered_matmul(A: np.ndarray, B: np.ndarray, descending: bool) -> np.ndarray:
    """A @ B accumulated from zero one k at a time (separately rounded products and sums)."""
    out = np.zeros((A.shape[0], B.shape[1]))
    ks = range(A.shape[1] - 1, -1, -1) if descending else range(A.shape[1])
    for k in ks:
        out += np.multiply.outer(A[:, k], B[k, :])
    return out


def chebyshev_first_matrix(x_internal: np.ndarray) -> np.ndarray:
    """Negated Trefethen differentiation matrix on the descending internal nodes.

    Scalar-by-scalar evaluation in the reference's operation order (numgrids/diff.py:204-222) so
    that every entry has the reference's bits.
    """
    x = np.asarray(x_internal, dtype=float)
    n = len(x)
    N = n - 1
    T = np.zeros((n, n))
    corner = (2 * N ** 2 + 1) / 6
    T[0, 0] = corner
    T[N, N] = -corner
    for j in range(1, N):
        T[j, j] = -0.5 * x[j] / (1 - x[j] ** 2)
    edge = (0, N)
    for i in range(n):
        ci = 2 if i in edge else 1
        for j in range(n):
            if j == i:
                continue
            cj = 2 if j in edge else 1
            T[i, j] = ci / cj * (-1) ** (i + j) / (x[i] - x[j])
    return -T


def chebyshev_matrix(x_internal: np.ndarray, coords: np.ndarray, order: int,
                     last_axis_of_nd: bool) -> np.ndarray:
    """1-D matrix with the entries of the reference's Kronecker-embedded ``_diff_matrix``.

    ``order-1`` self-products (reference quirk, numgrids/diff.py:190-191: order 3 -> D^4).  The
    k-direction of self-product t follows scipy: ascending on the last axis of an N-D grid (BSR
    block gemm); elsewhere SMMP flips the storage order after each product, so ascending for odd
    t, descending for even t.  Then one multiply by ``(2/scale)**order`` (diff.py:192).
    """
    M = chebyshev_first_matrix(x_internal)
    for t in range(1, order):
        M = _ordered_matmul(M, M, descending=(not last_axis_of_nd) and t % 2 == 0)
    scale = coords[-1] - coords[0]
    return M * (2 / scale) ** order


def chebyshev_descending(order: int, axis_index: int, ndims: int) -> bool:
    """j-direction of the reference's sparse matvec: CSR rows come out of SMMP reversed once per
    self-product, so a non-last axis of an N-D grid sums descending for even orders."""
    return ndims > 1 and axis_index != ndims - 1 and order % 2 == 0


# ------------------------------------------------------------------ FFT


def fft_multiplier(n: int, spacing: float, order: int) -> np.ndarray:
    """(i k 2 pi / period)**order with the Nyquist bin zeroed for even n and odd order."""
    period = n * spacing
    scale = 2 * np.pi / period
    k = np.fft.fftfreq(n) * n
    W = (1j * k * scale) ** order
    if n % 2 == 0 and order % 2 == 1:
        W[n // 2] = 0
    return W


def fft_dense_matrix(W: np.ndarray) -> np.ndarray:
    """Real matrix of F^-1 diag(W) F (what reference numgrids/diff.py:145-157 builds)."""
    n = len(W)
    return np.ascontiguousarray(
        np.real(np.fft.ifft(W[:, None] * np.fft.fft(np.eye(n), axis=0), axis=0)))


def is_pow2(n: int) -> bool:
    return n >= 2 and (n & (n - 1)) == 0


MIXED_RADIX_PRIMES = (2, 3, 5, 7, 11, 13)
MIXED_RADIX_MAX = 8192


def fft_mixed_supported(n: int) -> bool:
    """True if the library has a native n-point mixed-radix transform for this non-power-of-two length
    (csrc/fft_mixed.cu: prime factors up to 13, n <= 8192) -- mirrors ng_fft_mixed_supported()."""
    if n < 3 or n > MIXED_RADIX_MAX or is_pow2(n):
        return False
    for p in MIXED_RADIX_PRIMES:
        while n % p == 0:
            n //= p
    return n == 1


# ------------------------------------------------------------------ broadcast-compact coefficients


def compact_coefficient(a: np.ndarray, shape: tuple[int, ...]):
    """Reduce a full-shape (or broadcastable) array to the axes it actually varies along.

    Returns ``(values, strides)``: a contiguous 1-D table and per-axis element strides (0 where the
    array is constant along the axis).  Constancy is tested with bitwise equality (NaNs equal), so
    indexing the table reproduces the full array exactly.
    """
    a = np.broadcast_to(np.asarray(a, dtype=np.float64), shape)
    keep = []
    cur = a
    for ax in range(len(shape)):
        first = np.take(cur, [0], axis=ax)
        same = np.array_equal(cur, np.broadcast_to(first, cur.shape), equal_nan=True)
        if same:
            cur = first
        else:
            keep.append(ax)
    values = np.ascontiguousarray(cur)
    strides = [0] * len(shape)
    s = 1
    for ax in reversed(range(len(shape))):
        if values.shape[ax] > 1:
            strides[ax] = s
            s *= values.shape[ax]
    return values.reshape(-1), strides


def fft_padded_multiplier(W: np.ndarray):
    """Zero-padded power-of-two form of the periodic derivative for an axis of ANY length n.

    ``real(ifft(W * fft(f)))`` (numgrids/diff.py:133-143) is the circular convolution of f with the real kernel
    ``d = ifft(W)``.  Laid out on a circle of ``m >= 2n-1`` points (``d[k]`` at ``k``, ``d[n-k]`` at ``m-k``, zeros
    between), ``ifft_m(fft_m(d_circle) * fft_m(f zero-padded))`` reproduces it on the first n outputs.  Returns
    ``(m, fft_m(d_circle))`` with m the next power of two, or ``None`` when m would exceed the kernels' 8192."""
    n = len(W)
    m = 1
    while m < 2 * n - 1:
        m *= 2
    if m > 8192:
        return None
    d = np.real(np.fft.ifft(W))
    circle = np.zeros(m)
    circle[:n] = d
    circle[m - (n - 1):] = d[1:]
    return m, np.fft.fft(circle)
