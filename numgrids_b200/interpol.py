"""Interpolation of meshed data -- the reference's numgrids/interpol.py ``Interpolator``,
for ``method="linear"``, the method ``MultiGrid.transfer`` uses by default (numgrids/grids.py:291-327).

The reference wraps ``scipy.interpolate.RegularGridInterpolator`` (N-D) / ``interp1d`` (1-D).  For linear
interpolation both reduce to per-axis tables (interval index, normalised distance) and a fixed sum over
the 2^d corners; here the tables are built on the host with the same NumPy expressions and the sum runs
in ``ng_interp_linear`` (csrc/interp.cu) in SciPy's operation order, so results are bit-identical
(tests/test_interp.py).  The data may live on the device and stays there.  "quadratic" and "cubic" are
global spline fits inside SciPy (the cubic one solves for tensor-product B-spline coefficients
iteratively); they are not restated here and raise ``NotImplementedError``.  No CPU fallback.
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


class Interpolator:
    """Interpolant of array data meshed on a grid; call it with a Grid, a point or a list of points."""

    def __init__(self, grid, f, method: str = "cubic") -> None:
        if method not in ("linear", "quadratic", "cubic"):
            raise ValueError(f"Method '{method}' is not defined")
        if method != "linear":
            raise NotImplementedError(
                f"method={method!r} is a global spline fit inside SciPy and is not part of numgrids_b200; "
                "use method='linear' (device kernel) or the reference package")
        self.grid = grid
        self.method = method
        shape = grid.shape
        if any(n < 2 for n in shape):
            raise ValueError("linear interpolation needs at least 2 points per dimension")
        self._device_in = _device.is_device_array(f)
        self._torch_in = _device.is_torch_tensor(f)
        if self._device_in:
            self._values = _device.as_device_tensor(f, shape)
        else:
            self._values = _device.to_device(_device.as_host_array(f, shape))
        self._src = [np.asarray(a.coords, dtype=np.float64) for a in grid.axes]

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

    def _run(self, targets, out_shape, scattered: bool):
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
