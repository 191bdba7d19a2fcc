"""User-facing entry points -- drop-in for the hot-path part of the reference's numgrids/api.py:
``Diff`` (validation + dispatch, reference numgrids/api.py:31-72), ``AxisType`` / ``create_axis``
(:75-123), the built-in curvilinear grids (:126-251), the cached ``diff`` helper (:254-284), plus
the fused Cartesian Laplacian that has no single-call counterpart in the reference.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _device, _lib
from .axes import Axis, ChebyshevAxis, EquidistantAxis, LogAxis
from .curvilinear import CurvilinearGrid
from .grids import Grid


class Diff:
    """Partial derivative ``d^order / d x_axis^order`` on a grid; strategy chosen by the axis type."""

    def __init__(self, grid, order, axis_index=0, acc=4):
        if order <= 0:
            raise ValueError("Derivative order must be positive integer.")
        if not isinstance(grid, Grid):
            raise TypeError("Parameter 'grid' must be of type Grid.")
        if axis_index < 0:
            raise ValueError("axis must be nonnegative integer.")
        if axis_index > grid.ndims - 1:
            raise ValueError("No such axis index in this grid!")
        self.operator = grid.get_axis(axis_index).create_diff_operator(grid, order, axis_index, acc=acc)

    def __call__(self, f):
        return self.operator(f)

    def as_matrix(self):
        return self.operator.as_matrix()

    def as_linear_operator(self):
        """Matrix-free scipy ``LinearOperator`` (extension; the reference only offers ``as_matrix``)."""
        return self.operator.as_linear_operator()

    def apply_batch(self, fs):
        """K stacked device fields of shape (K, *grid.shape) in one launch (extension for launch-bound sizes)."""
        return self.operator.operator.apply_batch(fs)


class AxisType(str, enum.Enum):
    EQUIDISTANT = "equidistant"
    EQUIDISTANT_PERIODIC = "equidistant_periodic"
    CHEBYSHEV = "chebyshev"
    LOGARITHMIC = "log"


def create_axis(axis_type, num_points, low, high, **kwargs) -> Axis:
    if axis_type == AxisType.EQUIDISTANT:
        return EquidistantAxis(num_points, low, high, **kwargs)
    if axis_type == AxisType.EQUIDISTANT_PERIODIC:
        return EquidistantAxis(num_points, low, high, periodic=True, **kwargs)
    if axis_type == AxisType.CHEBYSHEV:
        return ChebyshevAxis(num_points, low, high, **kwargs)
    if axis_type == AxisType.LOGARITHMIC:
        return LogAxis(num_points, low, high, **kwargs)
    raise NotImplementedError(f"No such axis type: {axis_type}")


def _unit(c):
    return np.ones_like(c[0])


class SphericalGrid(CurvilinearGrid):
    """(r, theta, phi) with h = (1, r, r sin theta)."""

    def __init__(self, raxis, theta_axis, phi_axis):
        super().__init__(raxis, theta_axis, phi_axis,
                         scale_factors=(_unit, lambda c: c[0], lambda c: c[0] * np.sin(c[1])))


class CylindricalGrid(CurvilinearGrid):
    """(r, phi, z) with h = (1, r, 1)."""

    def __init__(self, raxis, phi_axis, zaxis):
        super().__init__(raxis, phi_axis, zaxis, scale_factors=(_unit, lambda c: c[0], _unit))


class PolarGrid(CurvilinearGrid):
    """(r, phi) with h = (1, r)."""

    def __init__(self, raxis, phi_axis):
        super().__init__(raxis, phi_axis, scale_factors=(_unit, lambda c: c[0]))


def diff(grid, f, order=1, axis_index=0, acc=4):
    """Differentiate with the operator cached on ``grid.cache['diffs'][(order, axis_index, acc)]``."""
    key = (order, axis_index, acc)
    ops = grid.cache.get("diffs")
    if key not in ops:
        ops[key] = Diff(grid, order, axis_index, acc=acc)
    return ops[key](f)


def interpolate(grid, f, locations, method="cubic"):
    """Interpolate a meshed function at a Grid, a point or a list of points (numgrids/api.py:286-306).

    The reference always builds a cubic interpolant here; ``method`` is an extension (``"linear"``, and
    ``"quadratic"`` on 1-D grids, see interpol.py)."""
    from .interpol import Interpolator
    return Interpolator(grid, f, method=method)(locations)


def integrate(grid, f):
    """Integrate a meshed function over the whole grid; the operator is cached on the grid
    under the reference's key (numgrids/api.py:309-335)."""
    from .integration import Integral
    if grid.cache.get("integral"):
        I = grid.cache["integral"]
    else:
        I = Integral(grid)
        grid.cache["integral"] = I
    return I(f)


class CartesianLaplacian:
    """Fused ``Diff(g,2,0)(f) + Diff(g,2,1)(f) + ...`` on equidistant non-periodic axes.

    The reference has no Cartesian ``laplacian``; users write the sum above (reference
    tests/test_diff.py:29-39).  This operator produces the same bits in one 16 B/point pass
    (csrc/fdlap.cu, csrc/fdlap3d.cu).
    """

    def __init__(self, grid, acc=4):
        if not isinstance(grid, Grid):
            raise TypeError("Parameter 'grid' must be of type Grid.")
        self.grid = grid
        self.acc = acc
        self._axis_ops = []
        for i in range(grid.ndims):
            d = Diff(grid, 2, i, acc=acc).operator
            if type(d).__name__ != "FiniteDifferenceDiff":
                raise TypeError("CartesianLaplacian needs equidistant, non-periodic axes")
            self._axis_ops.append(d)
        self._plan = None

    @property
    def plan(self) -> _lib.Plan:
        if self._plan is None:
            lib = _lib.load()
            _lib.require_gpu()
            children = [d.operator.plan for d in self._axis_ops]
            arr = _lib.ptr_array([p.handle.value for p in children])
            shp, shp_p = _lib.as_i64(list(self.grid.shape))
            out = C.c_void_p()
            _lib.check(lib.ng_fdlap_plan_create(self.grid.ndims, shp_p, arr, C.byref(out)),
                       "ng_fdlap_plan_create")
            self._plan = _lib.Plan(out.value, keepalive=children)
        return self._plan

    def apply_ptr(self, d_in, d_out, stream=None, lo=None, hi=None):
        st = _device.current_stream_ptr() if stream is None else stream
        _lib.check(_lib.load().ng_fdlap_apply_slab(self.plan.handle, d_in, d_out, lo, hi, st),
                   "ng_fdlap_apply")

    def __call__(self, f):
        shape = self.grid.shape
        if _device.is_complex(f):            # real linear operator: componentwise, like the three-call idiom
            re, im = _device.split_complex(f)
            return _device.join_complex(self(re), self(im))
        if _device.is_device_array(f):
            x = _device.as_device_tensor(f, shape)
            with _device.device_of(x):
                out = _device.empty_like_device(x)
                self.apply_ptr(x.data_ptr(), out.data_ptr())
            return out
        a = _device.as_host_array(f, shape)
        _lib.require_gpu()
        out = np.empty_like(a)
        _lib.check(_lib.load().ng_fdlap_apply_host(self.plan.handle, a.ctypes.data, out.ctypes.data),
                   "ng_fdlap_apply_host")
        if a.nbytes >= (1 << 30):                       # big arrays: hand the two device staging arrays back at once
            _lib.check(_lib.load().ng_plan_release_staging(self.plan.handle), "ng_plan_release_staging")
        return out


def cartesian_laplacian(grid, f, acc=4):
    """Cached helper: ``grid.cache['diffs'][('laplacian', acc)]`` holds the fused operator."""
    key = ("laplacian", acc)
    ops = grid.cache.get("diffs")
    if key not in ops:
        ops[key] = CartesianLaplacian(grid, acc=acc)
    return ops[key](f)
