"""numgrids_b200 -- B200-native (sm_100a) implementation of the numgrids differentiation hot path.

Drop-in for the reference's package surface (same names as numgrids/__init__.py): ``Grid``, ``MultiGrid``,
``CurvilinearGrid``, ``SphericalGrid``, ``CylindricalGrid``, ``PolarGrid``, ``create_axis`` / ``AxisType`` /
``Axis``, ``Diff``, ``diff``, ``integrate`` / ``Integral``, ``interpolate`` / ``Interpolator``,
``BoundaryFace`` / ``DirichletBC`` / ``NeumannBC`` / ``RobinBC`` / ``apply_bcs``, ``ErrorEstimator`` /
``AdaptationResult`` / ``adapt`` / ``estimate_error``, ``save_grid`` / ``load_grid``.

``Diff.__call__`` and ``grid.laplacian/gradient/divergence/curl`` -- the hot path -- and the operators around it
(boundary applies, quadrature, linear interpolation / multigrid transfer, refinement estimates) run hand-written
CUDA kernels through the C ABI in ``include/numgrids_b200.h``; there is no CPU fallback.  Sparse-matrix assembly
(``as_matrix``, ``apply_to_system``), the spline fit of the cubic ``Interpolator`` (SciPy's, as in the reference)
and file I/O are host code.  Not provided: plotting.
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
from .io import save_grid, load_grid
from .api import interpolate
from .boundary import BoundaryFace, DirichletBC, NeumannBC, RobinBC, apply_bcs
from ._lib import NumgridsB200Error

# Backward compatibility alias kept by the reference (numgrids/__init__.py:13)
Axis = create_axis

_OUT_OF_SCOPE = set()


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f"numgrids.{name} is outside the accelerated hot path (SURVEY.md section 2) and is not "
            "part of numgrids_b200; use the reference package for it.")
    raise AttributeError(f"module 'numgrids_b200' has no attribute {name!r}")
