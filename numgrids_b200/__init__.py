"""numgrids_b200 -- B200-native (sm_100a) implementation of the numgrids differentiation hot path.

Drop-in for the reference's hot-path API: ``Grid``, ``CurvilinearGrid``, ``SphericalGrid``,
``CylindricalGrid``, ``PolarGrid``, ``create_axis`` / ``AxisType`` / ``Axis``, ``Diff``, ``diff``,
plus the step that follows a derivative apply in a PDE loop: ``BoundaryFace``, ``DirichletBC``,
``NeumannBC``, ``RobinBC``, ``apply_bcs`` (in place on device arrays) and
``Integral`` / ``integrate`` (one streaming quadrature pass, N3), ``Interpolator`` / ``MultiGrid.transfer``
with ``method="linear"`` and the ``adapt`` / ``estimate_error`` refinement loop built on them (N4).
``Diff.__call__`` and ``grid.laplacian/gradient/divergence/curl`` run hand-written CUDA kernels
through the C ABI in ``include/numgrids_b200.h``; there is no CPU fallback.

Reference components outside the hot path (spline interpolation,
I/O, plotting -- SURVEY.md section 2) are intentionally not provided; accessing their
names raises ``NotImplementedError`` that says so.
"""
__version__ = "0.1.0"

from . import diff as _diff_module   # load the submodule first so that the `diff` helper below is
                                     # not shadowed when the operators import it lazily
from .grids import Grid, MultiGrid
from .curvilinear import CurvilinearGrid
from .api import (Diff, AxisType, create_axis, SphericalGrid, CylindricalGrid, PolarGrid, diff,
                  CartesianLaplacian, cartesian_laplacian, integrate)
from .integration import Integral
from .interpol import Interpolator
from .amr import ErrorEstimator, AdaptationResult, adapt, estimate_error
from .boundary import BoundaryFace, DirichletBC, NeumannBC, RobinBC, apply_bcs
from ._lib import NumgridsB200Error

# Backward compatibility alias kept by the reference (numgrids/__init__.py:13)
Axis = create_axis

_OUT_OF_SCOPE = {
    "save_grid", "load_grid", "interpolate",
}


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f"numgrids.{name} is outside the accelerated hot path (SURVEY.md section 2) and is not "
            "part of numgrids_b200; use the reference package for it.")
    raise AttributeError(f"module 'numgrids_b200' has no attribute {name!r}")
