"""Integral: the reference's numgrids/integration.py semantics.

CPU part: the NumPy restatement (oracle/integral_port.py) and the host-side weight vectors against
the golden vectors of the REAL reference (tests/golden/int_cases.npz, made by make_golden_int.py).
GPU part: ``Integral.__call__`` / ``integrate`` through the C ABI (``ng_quad_apply``).

Tolerance: the kernel sums in a different order from the reference's axis-by-axis sequential sums, so
the result is not bit-identical.  The bar is |I_gpu - I_ref| <= 1e-13 * sum|w0 w1 .. f| (the
condition-free scale of the sum), which is <= 1e-12 relative wherever the integral is not a
cancellation of much larger terms.
"""
import numpy as np
import pytest

import numgrids_b200 as ng
from conftest import load_golden
from oracle import integral_port as ip
from tests_support import build_axes

TOL_SCALE = 1e-13


def _scale(f, weights):
    a = np.abs(f)
    for w in weights:
        a = np.tensordot(np.abs(w), a, axes=(0, 0))
    return float(a)


def test_oracle_and_host_weights_match_reference_golden():
    z, meta = load_golden("int_cases.npz")
    assert len(meta["cases"]) >= 18
    for case in meta["cases"]:
        gi = case["grid"]
        grid = ng.Grid(*build_axes(ng, meta["grids"][gi]))
        weights = [z[f"g{gi}_w{i}"] for i in range(grid.ndims)]
        periodic = [bool(a.periodic) for a in grid.axes]
        f = z[f"g{gi}_{case['field']}_f"]
        assert np.array_equal(ip.integral(f, weights, periodic), z[f"g{gi}_{case['field']}_I"]), case
        host = ng.Integral(grid).weights                       # product-side tables, same bits
        for w_host, w_ref in zip(host, weights):
            # Same bits, except on FD axes of <= 9 points: there scipy.sparse.kron inside the findiff
            # restatement switches to BSR and stores the 1-D matrix with explicit zeros, which changes
            # SuperLU's elimination order inside inv() (4e-14 relative).  What the real findiff stores
            # at that size is unpinned.
            assert np.array_equal(w_host, w_ref) or (
                len(w_ref) <= 9 and np.max(np.abs(w_host - w_ref)) <= 1e-13 * np.max(np.abs(w_ref))), case


def test_weights_integrate_polynomials():
    ax = ng.create_axis(ng.AxisType.CHEBYSHEV, 20, -1.0, 2.0)
    w = ng.integration.axis_weights(ax)
    x = ax.coords
    for k in range(8):
        assert abs(w @ x ** k - (2.0 ** (k + 1) - (-1.0) ** (k + 1)) / (k + 1)) < 1e-10 * 2.0 ** (k + 1)
    per = ng.create_axis(ng.AxisType.EQUIDISTANT_PERIODIC, 16, 0.0, 2 * np.pi)
    assert np.array_equal(ng.integration.axis_weights(per), np.full(16, per.spacing))


@pytest.mark.gpu
def test_integral_golden(gpu_ready):
    import torch
    z, meta = load_golden("int_cases.npz")
    for case in meta["cases"]:
        gi = case["grid"]
        grid = ng.Grid(*build_axes(ng, meta["grids"][gi]))
        f = z[f"g{gi}_{case['field']}_f"]
        ref = float(z[f"g{gi}_{case['field']}_I"])
        I = ng.Integral(grid)
        tol = TOL_SCALE * _scale(f, I.weights)
        out = I(f)
        assert isinstance(out, np.float64) and abs(out - ref) <= tol, (case, out, ref, tol)
        out_d = I(torch.from_numpy(f).cuda())
        assert isinstance(out_d, torch.Tensor) and out_d.is_cuda and out_d.ndim == 0
        assert out_d.item() == out                              # deterministic, same path
        assert ng.integrate(grid, f) == out and grid.cache["integral"] is not None


@pytest.mark.gpu
def test_integral_large_and_odd_shapes(gpu_ready):
    """Config-size sum against the oracle's sequential order, plus odd last axes (scalar-load path)."""
    import torch
    for specs in ([("chebyshev", 128, 0.5, 2.0), ("chebyshev", 128, 0.3, 2.8), ("periodic", 256, 0.0, 2 * np.pi)],
                  [("equidistant", 37, 0.0, 1.0), ("equidistant", 1531, 0.0, 3.0)],
                  [("equidistant", 300, 0.0, 1.0), ("chebyshev", 2051, 0.0, 1.0)]):
        grid = ng.Grid(*build_axes(ng, specs))
        X = grid.meshed_coords
        f = np.exp(-X[0]) * np.cos(X[-1]) + 0.3 * X[0] * X[-1]
        I = ng.Integral(grid)
        ref = float(ip.integral(f, I.weights, [bool(a.periodic) for a in grid.axes]))
        tol = TOL_SCALE * _scale(f, I.weights)
        out = I(torch.from_numpy(f).cuda()).item()
        assert abs(out - ref) <= tol, (specs, out, ref, tol)
        assert abs(out - ref) <= 1e-12 * abs(ref)
