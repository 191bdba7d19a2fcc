"""NumPy restatement of the reference's array-level boundary conditions -- TEST INFRASTRUCTURE ONLY.

Follows reference numgrids/boundary.py: ``DirichletBC.apply`` (:157-166) sets ``u = g`` on the
face; ``NeumannBC.apply`` (:216-245) sets the face layer to ``u[adjacent] + h * normal_sign * g``
(first order; low face: h = x[1]-x[0], sign -1; high face: h = x[-1]-x[-2], sign +1);
``apply_to_system`` (:168-187, :247-261, :289-304) replaces the face rows of a sparse system by
identity / ``sign * D`` / ``a I + b sign D`` rows and the right-hand side by g.
Pinned bit for bit against the real reference by tests/golden/make_golden_bc.py.
"""
from __future__ import annotations

import numpy as np


def face_index(ndim: int, axis: int, side: str, adjacent: bool = False):
    idx = [slice(None)] * ndim
    if side == "low":
        idx[axis] = 1 if adjacent else 0
    else:
        idx[axis] = -2 if adjacent else -1
    return tuple(idx)


def face_values(value, coords, axis: int, side: str):
    """boundary.py:117-132: scalar -> broadcast, ndarray -> as is, callable(face coordinates)."""
    shape = [len(c) for c in coords]
    fshape = shape[:axis] + shape[axis + 1:]
    if callable(value):
        mesh = np.meshgrid(*coords, indexing="ij")
        sel = face_index(len(coords), axis, side)
        return np.asarray(value(tuple(m[sel] for m in mesh))).ravel()
    if isinstance(value, np.ndarray):
        return value.ravel()
    return np.full(int(np.prod(fshape)) if fshape else 1, float(value))


def dirichlet_apply(u: np.ndarray, axis: int, side: str, g: np.ndarray) -> np.ndarray:
    out = u.copy()
    sel = face_index(u.ndim, axis, side)
    out[sel] = g.reshape(out[sel].shape)
    return out


def neumann_apply(u: np.ndarray, axis: int, side: str, g: np.ndarray, coords_axis: np.ndarray) -> np.ndarray:
    out = u.copy()
    dst, src = face_index(u.ndim, axis, side), face_index(u.ndim, axis, side, adjacent=True)
    if side == "low":
        h, sign = coords_axis[1] - coords_axis[0], -1
    else:
        h, sign = coords_axis[-1] - coords_axis[-2], 1
    out[dst] = out[src] + h * sign * g.reshape(out[dst].shape)
    return out


def flat_face_indices(shape, axis: int, side: str) -> np.ndarray:
    m = np.zeros(shape, dtype=bool)
    m[face_index(len(shape), axis, side)] = True
    return np.flatnonzero(m)


def system_rows(L_dense: np.ndarray, rhs: np.ndarray, rows: np.ndarray, new_rows: np.ndarray, g: np.ndarray):
    """Dense form of the row replacement all three apply_to_system methods perform."""
    L, r = L_dense.copy(), rhs.copy()
    L[rows, :] = new_rows
    r[rows] = g
    return L, r
