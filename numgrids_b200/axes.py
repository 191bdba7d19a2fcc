"""Axis types -- same public surface as the reference's numgrids/axes.py (drop-in).

Mirrors (behaviour, not code): ``Axis`` base validation (reference numgrids/axes.py:55-58),
``EquidistantAxis`` (:238-256), ``ChebyshevAxis`` (:318-333), ``LogAxis`` (:364-379).  Each class
picks its differentiation strategy in ``create_diff_operator`` -- the seam the CUDA operators
plug into.
"""
from __future__ import annotations

import numpy as np


class Axis:
    """One coordinate axis: internal (canonical) and external (user) coordinates."""

    def __init__(self, num_points, low, high, periodic, **kwargs):
        if num_points <= 0:
            raise ValueError(f"num_points must be positive, not {num_points}")
        if high <= low:
            raise ValueError(f"high ({high}) must be greater than low ({low})")
        self._num_points = num_points
        self.low = low
        self.high = high
        self.periodic = bool(periodic)
        self.name = kwargs.get("name")
        self._coords_internal = self._setup_internal_coords(low, high)
        self._coords = self._setup_external_coords(low, high)

    # -- hooks --------------------------------------------------------
    def _setup_internal_coords(self, low, high):
        raise NotImplementedError("Must be implemented by child class.")

    def _setup_external_coords(self, low, high):
        return self._coords_internal * (high - low) + low

    def create_diff_operator(self, grid, order, axis_index, acc=4):
        raise NotImplementedError("Must be implemented by child class.")

    # -- public surface -----------------------------------------------
    def resized(self, num_points):
        extra = {}
        if self.periodic:
            extra["periodic"] = True
        if self.name is not None:
            extra["name"] = self.name
        return type(self)(num_points, self.low, self.high, **extra)

    @property
    def boundary(self):
        return slice(None, None) if self.periodic else slice(1, -1)

    @property
    def coords(self):
        return self._coords

    @property
    def coords_internal(self):
        return self._coords_internal

    def __len__(self):
        return self._num_points

    def __getitem__(self, idx):
        if self.periodic:
            return self._coords[idx % self._num_points]
        return self._coords[idx]

    def __str__(self):
        return f"{type(self).__name__}({len(self)} points from {self.coords[0]} to {self.coords[-1]})"

    def __repr__(self):
        return (f"{type(self).__name__}(num_points={len(self)}, low={self.coords[0]}, "
                f"high={self.coords[-1]}, periodic={self.periodic})")

    def plot(self):  # pragma: no cover - visualisation is out of scope
        raise NotImplementedError("plotting is not part of numgrids_b200")


class EquidistantAxis(Axis):
    """Uniform spacing; periodic axes drop the duplicate end point and differentiate spectrally."""

    def __init__(self, num_points, low=0, high=1, periodic=False, **kwargs):
        super().__init__(num_points, low, high, periodic, **kwargs)

    def _setup_internal_coords(self, *_):
        return np.linspace(0, 1, len(self), endpoint=not self.periodic)

    @property
    def spacing(self):
        return self._coords[1] - self._coords[0]

    def create_diff_operator(self, grid, order, axis_index, acc=4):
        from numgrids_b200.diff import FFTDiff, FiniteDifferenceDiff
        cls = FFTDiff if self.periodic else FiniteDifferenceDiff
        return cls(grid, order, axis_index, acc=acc)


class ChebyshevAxis(Axis):
    """Chebyshev extreme points cos(k pi/(n-1)) (descending internally), mapped ascending to [low, high]."""

    def __init__(self, num_points, low=0, high=1, **kwargs):
        super().__init__(num_points, low, high, periodic=False, **kwargs)

    def _setup_internal_coords(self, *_):
        n = len(self)
        return np.cos(np.arange(n) * np.pi / (n - 1))

    def _setup_external_coords(self, low, high):
        unit = (self._coords_internal[::-1] + 1) / 2
        return unit * (high - low) + low

    def create_diff_operator(self, grid, order, axis_index, acc=4):
        from numgrids_b200.diff import ChebyshevDiff
        return ChebyshevDiff(grid, order, axis_index)      # acc is irrelevant for the spectral matrix


class LogAxis(Axis):
    """Log-spaced points; differentiation is FD on ln x followed by the 1/x chain-rule factor."""

    def __init__(self, num_points, low, high, **kwargs):
        if low <= 0:
            raise ValueError("LogAxis requires positive lower boundary.")
        super().__init__(num_points, low, high, periodic=False, **kwargs)

    def _setup_internal_coords(self, low, high):
        return np.linspace(np.log(low), np.log(high), len(self))

    def _setup_external_coords(self, low, high):
        return np.logspace(np.log10(low), np.log10(high), len(self))

    def create_diff_operator(self, grid, order, axis_index, acc=6):
        from numgrids_b200.diff import LogDiff
        return LogDiff(grid, order, axis_index, acc=acc)
