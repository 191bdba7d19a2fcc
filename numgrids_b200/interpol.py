"""Interpolation of meshed data -- the reference's numgrids/interpol.py ``Interpolator``.

The reference wraps ``scipy.interpolate.RegularGridInterpolator`` (N-D) / ``interp1d`` (1-D).

* ``method="linear"`` (what ``MultiGrid.transfer`` and the refinement loop use): per-axis tables (interval index,
  normalised distance) and a fixed sum over the 2^d corners; the tables are built on the host with the same NumPy
  expressions and the sum runs in ``ng_interp_linear`` (csrc/interp.cu) in SciPy's operation order -- bit-identical.
* ``method="cubic"`` (the reference's default) and, in 1-D, ``"quadratic"``: SciPy fits a tensor-product B-spline
  (N-D: ``make_ndbspl`` with the iterative ``gcrotmk`` solver; 1-D: ``make_interp_spline``, not-a-knot).  The fit is
  a global solve and stays SciPy's, on the host, exactly as in the reference (same call, same coefficients); the
  evaluation -- the data-parallel part -- runs in ``ng_interp_bspline`` from the coefficients and per-axis basis
  tables (SciPy's de Boor recursion), summing the (k+1)^d terms in SciPy's order: bit-identical to SciPy's evaluation.

The data may live on the device; results follow the container of the data.  No CPU fallback for the evaluation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _device, _lib
from .grids import Grid


def axis_table(src: np.ndarray, dst: np.ndarray, dim: int):
    """Interval index and normalised distance per target coordinate (scipy's ``find_indices``)."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    if dst.size and (np.any(dst < src[0]) or np.any(dst > src[-1])):
        raise ValueError(f"One of the requested xi is out of bounds in dimension {dim}")
    i = np.clip(np.searchsorted(src, dst, side="right") - 1, 0, len(src) - 2)
    return i, src[i], src[i + 1]


def spline_fit(src_coords, values: np.ndarray, k: int):
    """SciPy's own B-spline fit, as the reference runs it: (knot vectors, coefficient array)."""
    if len(src_coords) > 1:
        import scipy.sparse.linalg as ssl
        try:
            from scipy.interpolate import make_ndbspl
        except ImportError:  # pragma: no cover
            from scipy.interpolate._ndbspline import make_ndbspl
        spl = make_ndbspl(tuple(src_coords), values, k, solver=ssl.gcrotmk)     # RegularGridInterpolator's default solver
        return [np.asarray(t, dtype=np.float64) for t in spl.t], np.ascontiguousarray(spl.c, dtype=np.float64)
    from scipy.interpolate import make_interp_spline
    spl = make_interp_spline(src_coords[0], values, k=k)                        # interp1d's spline (not-a-knot)
    return [np.asarray(spl.t, dtype=np.float64)], np.ascontiguousarray(spl.c, dtype=np.float64).reshape(-1)


def spline_tables(knots, k: int, src_coords, targets):
    """Per axis: first coefficient index and the k+1 non-zero basis values per target coordinate (SciPy's de Boor)."""
    from scipy.interpolate import BSpline
    first, basis = [], []
    for a, (src, dst) in enumerate(zip(src_coords, targets)):
        dst = np.asarray(dst, dtype=np.float64)
        if dst.size and (np.any(dst < src[0]) or np.any(dst > src[-1])):
            raise ValueError(f"One of the requested xi is out of bounds in dimension {a}")
        if dst.size:
            dm = BSpline.design_matrix(dst, knots[a], k).tocsr()
            first.append(np.ascontiguousarray(dm.indices[dm.indptr[:-1]], dtype=np.int32))
            basis.append(np.ascontiguousarray(dm.data.reshape(len(dst), k + 1), dtype=np.float64))
        else:
            first.append(np.zeros(0, dtype=np.int32))
            basis.append(np.zeros((0, k + 1)))
    return first, basis


class Interpolator:
    """Interpolant of array data meshed on a grid; call it with a Grid, a point or a list of points."""

    def __init__(self, grid, f, method: str = "cubic") -> None:
        degree = {"linear": 1, "quadratic": 2, "cubic": 3}
        if method not in degree or (method == "quadratic" and grid.ndims > 1):
            # (RegularGridInterpolator knows no "quadratic": the reference raises the same error there)
            raise ValueError(f"Method '{method}' is not defined")
        self.grid = grid
        self.method = method
        shape = grid.shape
        k = degree[method]
        for i, n in enumerate(shape):
            if n <= k and method != "linear":
                raise ValueError(f"There are {n} points in dimension {i}, but method {method} requires at least "
                                 f" {k + 1} points per dimension.")
        if any(n < 2 for n in shape):
            raise ValueError("linear interpolation needs at least 2 points per dimension")
        self._device_in = _device.is_device_array(f)
        self._torch_in = _device.is_torch_tensor(f)
        self._src = [np.asarray(a.coords, dtype=np.float64) for a in grid.axes]
        self._spline = None
        if method == "linear":
            if self._device_in:
                self._values = _device.as_device_tensor(f, shape)
            else:
                self._values = _device.to_device(_device.as_host_array(f, shape))
            return
        # spline methods: SciPy's own fit on the host (what the reference runs), evaluation on the device
        knots, coef = spline_fit(self._src, _device.as_host_array(f, shape), k)
        self._spline = (knots, k)
        self._values = _device.to_device(np.ascontiguousarray(coef))

    # -- tables ----------------------------------------------------------
    def _tables(self, targets):
        """Device tables (idx, y, 1-y) per axis for the target coordinates of each axis."""
        t = _device.torch()
        nd = len(self._src)
        idx, ys, omys = [], [], []
        for a, (src, dst) in enumerate(zip(self._src, targets)):
            i, x0, x1 = axis_table(src, dst, a)
            if nd == 1:                                   # numpy.interp's form: slope * (xq - x[j]) + v[j]
                y = np.asarray(dst, dtype=np.float64) - x0
                omy = x1 - x0
                omy = np.where(np.asarray(dst) == src[-1], 0.0, omy)      # x == last knot -> v[-1] exactly
            else:
                y = (np.asarray(dst, dtype=np.float64) - x0) / (x1 - x0)
                omy = 1 - y
            idx.append(t.from_numpy(np.ascontiguousarray(i, dtype=np.int32)).cuda())
            ys.append(t.from_numpy(np.ascontiguousarray(y)).cuda())
            omys.append(t.from_numpy(np.ascontiguousarray(omy)).cuda())
        return idx, ys, omys

    def _run_spline(self, targets, out_shape, scattered: bool):
        """B-spline evaluation: per-axis (first coefficient index, k+1 basis values) tables from SciPy's design matrix."""
        _lib.require_gpu()
        lib = _lib.load()
        t = _device.torch()
        knots, k = self._spline
        nd = len(knots)
        fi, bv = spline_tables(knots, k, self._src, targets)
        first = [t.from_numpy(x).cuda() for x in fi]
        basis = [t.from_numpy(x).cuda() for x in bv]
        out = t.empty(tuple(out_shape), dtype=t.float64, device="cuda")
        ptrs = lambda arrs: (C.c_void_p * nd)(*[x.data_ptr() for x in arrs])
        _k1, cshape_p = _lib.as_i64(list(self._values.shape))
        dshape = [len(targets[0])] * nd if scattered else [len(x) for x in targets]
        _k2, dst_p = _lib.as_i64(dshape)
        _k3, nb_p = _lib.as_i32([k + 1] * nd)
        _lib.check(lib.ng_interp_bspline(nd, cshape_p, dst_p, int(scattered), nb_p, ptrs(first), ptrs(basis),
                                         self._values.data_ptr(), out.data_ptr(), _device.current_stream_ptr()),
                   "ng_interp_bspline")
        return out

    def _run(self, targets, out_shape, scattered: bool):
        if self._spline is not None:
            out = self._run_spline(targets, out_shape, scattered)
            if self._device_in:
                return out
            res = out.cpu().numpy()
            return _device.torch().from_numpy(res) if self._torch_in else res
        _lib.require_gpu()
        lib = _lib.load()
        t = _device.torch()
        nd = len(self._src)
        idx, ys, omys = self._tables(targets)
        out = t.empty(tuple(out_shape), dtype=t.float64, device="cuda")
        ptrs = lambda arrs: (C.c_void_p * nd)(*[x.data_ptr() for x in arrs])
        _, src_p = _lib.as_i64(list(self.grid.shape))
        dshape = [len(targets[0])] * nd if scattered else [len(x) for x in targets]
        _keep, dst_p = _lib.as_i64(dshape)
        _lib.check(lib.ng_interp_linear(nd, src_p, dst_p, int(scattered), ptrs(idx), ptrs(ys), ptrs(omys),
                                        self._values.data_ptr(), out.data_ptr(), _device.current_stream_ptr()),
                   "ng_interp_linear")
        if self._device_in:
            return out
        res = out.cpu().numpy()
        return _device.torch().from_numpy(res) if self._torch_in else res

    # -- user level ------------------------------------------------------
    def __call__(self, locations):
        """Interpolate at a Grid (all its points), one point (tuple) or several points (list of tuples / zip)."""
        nd = self.grid.ndims
        if isinstance(locations, Grid):
            if locations.ndims != nd:
                raise ValueError(f"The requested sample points xi have dimension {locations.ndims}, "
                                 f"but this Interpolator has dimension {nd}")
            targets = [np.asarray(a.coords, dtype=np.float64) for a in locations.axes]
            return self._run(targets, locations.shape, scattered=False)
        if not hasattr(locations, "__iter__") and not hasattr(locations, "__len__"):
            locations = [locations]
        if isinstance(locations, zip):
            locations = list(locations)
        locations = np.array(locations, dtype=np.float64)
        single = locations.ndim == 1
        pts = locations.reshape(1, -1) if single and nd > 1 else (locations.reshape(-1, 1) if nd == 1 else locations)
        if pts.ndim != 2 or pts.shape[1] != nd:
            raise ValueError(f"The requested sample points xi have dimension {pts.shape[-1]}, "
                             f"but this Interpolator has dimension {nd}")
        res = self._run([np.ascontiguousarray(pts[:, a]) for a in range(nd)], (pts.shape[0],), scattered=True)
        if single:                  # the reference returns `self._inter(locations)[0]` for any 1-D input
            return res[0]           # (interpol.py:72-73), i.e. the first value even on a 1-D grid
        return res
