"""Pin of oracle/findiff_port.py against the REAL third-party findiff, whenever that package is importable.

The reference differentiates equidistant and logarithmic axes with ``findiff.FinDiff`` (reference
numgrids/diff.py:10, :88, :259; pinned only as ``findiff>=0.10`` in pyproject.toml:23).  findiff is not
vendored under /root/reference and is not installed in the build image, so the finite-difference oracle is
a restatement of its published algorithm ("parity unpinned").  These tests close that gap
on any machine where the real package is present -- a developer box, or a GPU box provisioned with
``baseline/_ref`` -- and are skipped elsewhere.
"""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _real_findiff():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)                     # a driver-provisioned reference install, if any
    try:
        mod = importlib.import_module("findiff")
    except ImportError:
        return None
    return None if getattr(mod, "__version__", "").endswith("-port") else mod


findiff = _real_findiff()
needs_findiff = pytest.mark.skipif(findiff is None, reason="the real findiff package is not installed here")


@needs_findiff
@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("acc", [2, 4, 6])
def test_port_coefficients_equal_findiff(order, acc):
    from oracle import findiff_port
    want = findiff.coefficients(deriv=order, acc=acc)
    got = findiff_port.coefficients(order, acc)
    for scheme in ("center", "forward", "backward"):
        assert list(got[scheme]["offsets"]) == list(want[scheme]["offsets"]), scheme
        assert np.array_equal(np.asarray(got[scheme]["coefficients"], dtype=float),
                              np.asarray(want[scheme]["coefficients"], dtype=float)), scheme


@needs_findiff
@pytest.mark.parametrize("order,acc", [(1, 4), (2, 4), (1, 6), (2, 6), (3, 4)])
def test_port_apply_and_matrix_equal_findiff(order, acc):
    from oracle import findiff_port
    rng = np.random.default_rng(11)
    f = rng.standard_normal((17, 23, 19))
    for axis, h in ((0, 0.037), (1, 1.0 / 22), (2, 2 * np.pi / 19)):
        want = findiff.FinDiff(axis, h, order, acc=acc)(f)
        got = findiff_port.FinDiff(axis, h, order, acc=acc)(f)
        assert np.array_equal(got, want), (axis, float(np.max(np.abs(got - want))))
    want = findiff.FinDiff(1, 0.05, order, acc=acc).matrix((5, 16)).toarray()
    got = findiff_port.FinDiff(1, 0.05, order, acc=acc).matrix((5, 16)).toarray()
    assert np.array_equal(got, want)


@needs_findiff
@pytest.mark.gpu
def test_device_fd_equals_real_findiff(gpu_ready):
    """The CUDA finite-difference path against the real package directly (not through the port)."""
    import numgrids_b200 as ng
    A = ng.AxisType
    axes = [ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0 + i) for i, n in enumerate((40, 36, 64))]
    g = ng.Grid(*axes)
    rng = np.random.default_rng(12)
    f = rng.standard_normal(g.shape)
    for order in (1, 2):
        for axis in range(3):
            want = findiff.FinDiff(axis, axes[axis].spacing, order, acc=4)(f)
            assert np.array_equal(ng.Diff(g, order, axis)(f), want), (order, axis)


def test_port_is_marked_as_a_port():
    """The loader must never mistake the restatement for the real package."""
    from oracle import findiff_port, reference_loader
    assert findiff_port.__version__.endswith("-port")
    if findiff is None:
        assert not reference_loader.real_findiff_available()
