"""NumPy restatement of the linear interpolation the reference delegates to SciPy -- TEST INFRASTRUCTURE ONLY.

Reference numgrids/interpol.py:34-46 builds ``RegularGridInterpolator(grid.coords, f, method)`` (N-D) or
``interp1d(grid.coords, f, kind=method)`` (1-D); ``__call__`` (:60-75) evaluates it at all points of a
target grid or at a list of points.  For ``method="linear"`` SciPy 1.18 computes
  1-D  numpy.interp:        slope = (v[j+1]-v[j])/(x[j+1]-x[j]); slope*(xq-x[j]) + v[j]; knots exactly
  2-D  evaluate_linear_2d:  v00*(1-y0)*(1-y1) + v01*(1-y0)*y1 + v10*y0*(1-y1) + v11*y0*y1
  N-D  _evaluate_linear:    0 + sum over hypercube corners (first axis slowest) of v * (((1*w0)*w1)*...)
Pinned bit for bit against the real reference classes by tests/golden/make_golden_interp.py.
"""
from __future__ import annotations

import itertools

import numpy as np


def axis_table(src, dst):
    i = np.clip(np.searchsorted(src, dst, side="right") - 1, 0, len(src) - 2)
    return i, (dst - src[i]) / (src[i + 1] - src[i])


def linear_points(values: np.ndarray, src_coords, pts: np.ndarray) -> np.ndarray:
    """Interpolate at points pts[npts, ndim] (each point separately, like the reference's loop)."""
    nd = values.ndim
    if nd == 1:
        s, d = np.asarray(src_coords[0]), pts[:, 0]
        j = np.clip(np.searchsorted(s, d, side="right") - 1, 0, len(s) - 2)
        slope = (values[j + 1] - values[j]) / (s[j + 1] - s[j])
        res = slope * (d - s[j]) + values[j]
        res = np.where(d == s[j], values[j], res)
        return np.where(d == s[-1], values[-1], res)
    tabs = [axis_table(np.asarray(s), pts[:, a]) for a, s in enumerate(src_coords)]
    if nd == 2:
        (i0, y0), (i1, y1) = tabs
        return (values[i0, i1] * (1 - y0) * (1 - y1) + values[i0, i1 + 1] * (1 - y0) * y1
                + values[i0 + 1, i1] * y0 * (1 - y1) + values[i0 + 1, i1 + 1] * y0 * y1)
    value = np.zeros(len(pts))
    for corner in itertools.product((0, 1), repeat=nd):
        weight = np.ones(len(pts))
        idx = []
        for (i, y), b in zip(tabs, corner):
            weight = weight * ((1 - y) if b == 0 else y)
            idx.append(i + b)
        value = value + values[tuple(idx)] * weight
    return value


def linear_on_grid(values: np.ndarray, src_coords, dst_coords) -> np.ndarray:
    mesh = np.meshgrid(*dst_coords, indexing="ij")
    pts = np.array([m.reshape(-1) for m in mesh]).T
    return linear_points(values, src_coords, pts).reshape([len(d) for d in dst_coords])


def bspline_from_tables(coef: np.ndarray, first, basis, scattered: bool = False) -> np.ndarray:
    """SciPy's B-spline evaluation order (NdBSpline / BSpline): for every target, the (k+1)^d terms with the last
    axis fastest, factor = ((1*b_0)*b_1)*..., value = value + c*factor.  `first[a]` / `basis[a]` are the per-axis
    tables (first coefficient index, basis values) per target coordinate -- tensor-product targets unless
    `scattered`, where every table is indexed by the point number."""
    nd = coef.ndim
    out_shape = [len(first[0])] if scattered else [len(f) for f in first]
    res = np.zeros(out_shape)
    for combo in itertools.product(*[range(b.shape[1]) for b in basis]):
        factor = np.ones(out_shape)
        idx = []
        for a, j in enumerate(combo):
            shp = [-1] if scattered else [1] * nd
            if not scattered:
                shp[a] = -1
            factor = factor * basis[a][:, j].reshape(shp)
            idx.append((first[a].astype(np.int64) + j).reshape(shp))
        res = res + coef[tuple(idx)] * factor
    return res
