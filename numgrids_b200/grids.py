"""Tensor-product grid -- the API surface of the reference's numgrids/grids.py ``Grid`` (drop-in).

Differences that matter at config sizes: the reference materialises
``np.meshgrid`` and the boundary mask in ``__init__`` (numgrids/grids.py:27-33; 26 GB at 1024^3);
here both are built on first access, so a 1024^3 grid costs three 8 KB coordinate vectors until
somebody asks for meshed coordinates.  ``MultiGrid`` (multigrid hierarchy) is out of scope.
"""
from __future__ import annotations

import numpy as np


class Grid:
    """N-dimensional grid = ordered tuple of axes (axis 0 first)."""

    def __init__(self, *axes):
        self._axes = axes
        self.ndims = len(axes)
        self._meshed_coords = None
        self._boundary = None
        # shared scratch for the api helpers (same keys as the reference, numgrids/grids.py:36-41)
        self.cache = {"diffs": {}, "integrals": None, "interpolators": {}, "faces": None}

    # -- structure ----------------------------------------------------
    @property
    def axes(self):
        return self._axes

    def get_axis(self, idx=0):
        return self.axes[idx]

    @property
    def shape(self):
        return tuple(len(a) for a in self.axes)

    @property
    def size(self):
        return np.prod(self.shape)

    def __repr__(self):
        return f"Grid(shape={self.shape}, ndims={self.ndims})"

    # -- coordinates --------------------------------------------------
    @property
    def coords(self):
        if self.ndims == 1:
            return self.axes[0].coords
        return tuple(a.coords for a in self.axes)

    @property
    def meshed_coords(self):
        if self._meshed_coords is None:
            self._meshed_coords = np.meshgrid(*(a.coords for a in self.axes), indexing="ij")
        return self._meshed_coords

    def __getitem__(self, inds):
        if self.ndims == 1:
            return self.axes[0][inds]
        return np.array(tuple(self.axes[k][inds[k]] for k in range(self.ndims)))

    @property
    def boundary(self):
        if self._boundary is None:
            mask = np.ones(self.shape, dtype=bool)
            mask[tuple(a.boundary for a in self.axes)] = False
            self._boundary = mask
        return self._boundary

    @property
    def faces(self):
        """All boundary faces keyed '{axis_index}_{side}'; periodic axes are skipped (numgrids/grids.py:124-139)."""
        if self.cache["faces"] is None:
            from .boundary import BoundaryFace
            result = {}
            for i, axis in enumerate(self.axes):
                if not axis.periodic:
                    result[f"{i}_low"] = BoundaryFace(self, i, "low")
                    result[f"{i}_high"] = BoundaryFace(self, i, "high")
            self.cache["faces"] = result
        return self.cache["faces"]

    @property
    def meshed_indices(self):
        return np.meshgrid(*(np.arange(len(a)) for a in self.axes), indexing="ij")

    @property
    def index_tuples(self):
        return self._tuple_field(*self.meshed_indices)

    @property
    def coord_tuples(self):
        return self._tuple_field(*self.meshed_coords)

    def _tuple_field(self, *arrs):
        flat = np.vstack([a.reshape(-1) for a in arrs]).T
        return flat.reshape(*self.shape, self.ndims)

    # -- resizing (new grids; operators are re-planned lazily) ---------
    def refine(self):
        return Grid(*(a.resized(len(a) * 2) for a in self.axes))

    def coarsen(self):
        return Grid(*(a.resized(len(a) // 2) for a in self.axes))

    def refine_axis(self, axis_index, factor=2.0):
        if factor <= 0:
            raise ValueError(f"factor must be positive, got {factor}")
        if axis_index < 0 or axis_index >= self.ndims:
            raise ValueError(f"axis_index {axis_index} out of range for {self.ndims}-dimensional grid")
        axes = list(self.axes)
        axes[axis_index] = axes[axis_index].resized(max(2, round(len(axes[axis_index]) * factor)))
        return Grid(*axes)


class MultiGrid:
    """Hierarchy of grids of decreasing resolution on one domain (reference numgrids/grids.py:243-327).

    ``levels[0]`` is the finest grid; every further level has about half the points per axis
    (``n -> n // 2 + n % 2``) until an axis would drop below ``min_size``.
    """

    def __init__(self, *axes, min_size: int = 2) -> None:
        grid = Grid(*axes)
        self._levels = [grid]
        while True:
            sizes = np.array(grid.shape)
            sizes = sizes // 2 + sizes % 2
            if min(sizes) < min_size:
                break
            axes = [axis.resized(int(size)) for size, axis in zip(sizes, axes)]
            grid = Grid(*axes)
            self._levels.append(grid)

    @property
    def levels(self):
        return self._levels

    def transfer(self, f, level_from: int, level_to: int, method: str = "linear"):
        """Coarsen or refine an array by one level (interpolation onto the neighbouring level's grid);
        the array may live on the device and the result stays there."""
        from .interpol import Interpolator
        if tuple(f.shape) != self.levels[level_from].shape:
            raise ValueError(
                f"Array shape {tuple(f.shape)} does not match grid shape "
                f"{self.levels[level_from].shape} at level {level_from}."
            )
        if abs(level_from - level_to) != 1:
            raise ValueError("Can only transfer between adjacent levels.")
        return Interpolator(self.levels[level_from], f, method=method)(self.levels[level_to])
