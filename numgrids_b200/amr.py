"""Per-axis adaptive refinement -- the reference's numgrids/amr.py surface.

Same classes, arguments and control flow as the reference (``ErrorEstimator`` :47-190,
``AdaptationResult`` :193-221, ``adapt`` :224-350, ``estimate_error`` :353-394); what changes is where
the arrays live.  The comparison step of every estimate -- interpolate the refined field back onto the
current grid (``Interpolator(..., method="linear")``) and take a norm of the difference -- runs on the
device (``ng_interp_linear`` + ``ng_diff_norm``), so a ``func`` that builds its field with device
operators never brings it to the host.  ``func(grid)`` may return a NumPy array or a device array.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _device, _lib
from .interpol import Interpolator

_NORMS = {"max": 0, "l2": 1, "mean": 2}


def _to_device(f, shape):
    if _device.is_device_array(f):
        return _device.as_device_tensor(f, shape)
    return _device.to_device(_device.as_host_array(f, shape))


def _diff_norm(a, b, kind: int) -> float:
    """Norm of ``a - b`` for two device arrays of one shape (one pass, ng_diff_norm)."""
    lib = _lib.load()
    t = _device.torch()
    work = t.empty(int(lib.ng_quad_workspace_doubles()), dtype=t.float64, device="cuda")
    out = t.empty((), dtype=t.float64, device="cuda")
    _lib.check(lib.ng_diff_norm(a.numel(), a.data_ptr(), b.data_ptr(), kind, work.data_ptr(), out.data_ptr(),
                                _device.current_stream_ptr()), "ng_diff_norm")
    return float(out.item())


class ErrorEstimator:
    """Discretisation-error estimate from re-evaluating ``func`` on refined grids."""

    def __init__(self, grid, func, norm: str = "max") -> None:
        self.grid = grid
        self.func = func
        if norm not in _NORMS:
            raise ValueError(
                f"Unknown norm {norm!r}. Use 'max', 'l2', or 'mean'."
            )
        self.norm = norm
        self._f_current = None
        self._f_current_dev = None

    def _compute_norm(self, diff) -> float:
        """Norm of an already formed difference array (kept for API parity; device path: _diff_norm)."""
        d = _to_device(diff, tuple(diff.shape))
        return _diff_norm(d, _device.torch().zeros_like(d), _NORMS[self.norm])

    @property
    def f_current(self):
        if self._f_current is None:
            self._f_current = self.func(self.grid)
        return self._f_current

    def _current_on_device(self):
        if self._f_current_dev is None:
            self._f_current_dev = _to_device(self.f_current, self.grid.shape)
        return self._f_current_dev

    def _error_against(self, refined_grid) -> float:
        f_refined = _to_device(self.func(refined_grid), refined_grid.shape)
        back = Interpolator(refined_grid, f_refined, method="linear")(self.grid)
        return _diff_norm(self._current_on_device(), back, _NORMS[self.norm])

    def global_error(self, refinement_factor: float = 2.0) -> float:
        """Refine every axis, re-evaluate, interpolate back, norm of the difference."""
        fine = self.grid
        for i in range(self.grid.ndims):
            fine = fine.refine_axis(i, refinement_factor)
        return self._error_against(fine)

    def per_axis_errors(self, refinement_factor: float = 2.0) -> dict:
        """The same with one axis refined at a time: {axis index: error}."""
        return {i: self._error_against(self.grid.refine_axis(i, refinement_factor))
                for i in range(self.grid.ndims)}

    def axis_needing_refinement(self, refinement_factor: float = 2.0):
        errors = self.per_axis_errors(refinement_factor)
        if not errors or max(errors.values()) == 0.0:
            return None
        return max(errors, key=errors.get)


@dataclass
class AdaptationResult:
    """Adapted grid, the field on it, the final error estimates and the per-iteration history."""
    grid: object
    f: object
    errors: dict
    global_error: float
    iterations: int
    converged: bool
    history: list = field(default_factory=list)


def adapt(grid, func, tol: float = 1e-6, norm: str = "max", max_iterations: int = 20,
          max_points_per_axis: int = 1024, refinement_factor: float = 2.0,
          refine_all: bool = False) -> AdaptationResult:
    """Refine the axis that contributes most error (or, with ``refine_all``, every axis above
    ``tol / ndims``) until the global estimate drops below ``tol``, ``max_iterations`` is reached or no
    axis may grow further (``max_points_per_axis``)."""
    history = []
    iterations = 0
    for iteration in range(max_iterations):
        est = ErrorEstimator(grid, func, norm=norm)
        axis_errors = est.per_axis_errors(refinement_factor)
        g_error = est.global_error(refinement_factor)
        record = {"iteration": iteration, "shape": grid.shape, "global_error": g_error,
                  "axis_errors": dict(axis_errors)}
        history.append(record)
        iterations = iteration + 1
        if g_error < tol:
            return AdaptationResult(grid, est.f_current, axis_errors, g_error, iterations, True, history)
        if refine_all:
            chosen = [i for i, err in axis_errors.items()
                      if err > tol / grid.ndims and len(grid.axes[i]) < max_points_per_axis]
        else:
            worst = max(axis_errors, key=axis_errors.get)
            chosen = [worst] if len(grid.axes[worst]) < max_points_per_axis else []
        if not chosen:
            break
        for i in chosen:
            grid = grid.refine_axis(i, refinement_factor)
            record[f"refined_axis_{i}"] = True
    est = ErrorEstimator(grid, func, norm=norm)
    axis_errors = est.per_axis_errors(refinement_factor)
    g_error = est.global_error(refinement_factor)
    return AdaptationResult(grid, est.f_current, axis_errors, g_error, iterations, g_error < tol, history)


def estimate_error(grid, func, norm: str = "max", refinement_factor: float = 2.0) -> dict:
    """One-shot report: ``{"global": ..., "per_axis": {...}}`` (numgrids/amr.py:353-394)."""
    est = ErrorEstimator(grid, func, norm=norm)
    return {"global": est.global_error(refinement_factor),
            "per_axis": est.per_axis_errors(refinement_factor)}
