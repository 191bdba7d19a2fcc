"""Saving and loading grids with meshed data -- the reference's numgrids/io.py file format (host I/O).

``save_grid`` / ``load_grid`` keep the reference's on-disk layout (numgrids/io.py:54-140): one ``.npz`` whose
``__meta__`` entry is the UTF-8 JSON description of the grid (grid class name; per axis: class name, number of
points, low, high, periodic flag, name) and whose ``data_<name>`` entries are the arrays, so files written by
either package load in the other.  Arrays that live on the device are brought to the host for writing; loaded
arrays are NumPy arrays (upload them with ``torch.from_numpy(a).cuda()`` to keep working on the device).
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from . import _device


def _classes():
    from .api import CylindricalGrid, PolarGrid, SphericalGrid
    from .axes import ChebyshevAxis, EquidistantAxis, LogAxis
    from .grids import Grid
    axis_classes = {c.__name__: c for c in (EquidistantAxis, ChebyshevAxis, LogAxis)}
    grid_classes = {c.__name__: c for c in (Grid, SphericalGrid, CylindricalGrid, PolarGrid)}
    return axis_classes, grid_classes


def _describe(grid) -> dict:
    axis_classes, grid_classes = _classes()
    grid_type = type(grid).__name__
    if grid_type not in grid_classes:
        raise TypeError(f"Unsupported grid type: {grid_type}")
    axes = []
    for axis in grid.axes:
        axis_type = type(axis).__name__
        if axis_type not in axis_classes:
            raise TypeError(f"Unsupported axis type: {axis_type}")
        coords = axis.coords
        high = float(coords[-1])
        if axis.periodic and len(axis) > 1:          # the endpoint is not a grid point: one spacing further
            high += float(coords[1] - coords[0])
        axes.append({"type": axis_type, "num_points": len(axis), "low": float(coords[0]), "high": high,
                     "periodic": axis.periodic, "name": axis.name})
    return {"grid_type": grid_type, "axes": axes}


def _rebuild(meta: dict):
    axis_classes, grid_classes = _classes()
    axes = []
    for ax in meta["axes"]:
        extra = {}
        if ax.get("periodic"):
            extra["periodic"] = True
        if ax.get("name") is not None:
            extra["name"] = ax["name"]
        axes.append(axis_classes[ax["type"]](ax["num_points"], ax["low"], ax["high"], **extra))
    return grid_classes[meta["grid_type"]](*axes)


def save_grid(path, grid, **data) -> None:
    """Write the grid description and the named meshed arrays (each of shape ``grid.shape``) to ``path`` (.npz)."""
    arrays = {}
    for name, arr in data.items():
        if tuple(arr.shape) != tuple(grid.shape):
            raise ValueError(
                f"Array '{name}' has shape {tuple(arr.shape)}, "
                f"expected grid shape {grid.shape}"
            )
        host = _device.as_host_array(arr) if (_device.is_device_array(arr) or _device.is_torch_tensor(arr)) else arr
        arrays[f"data_{name}"] = host
    meta = np.array(np.void(json.dumps(_describe(grid)).encode("utf-8")))
    np.savez(str(path), __meta__=meta, **arrays)


def load_grid(path):
    """Return ``(grid, {name: array})`` from a file written by ``save_grid`` (of this package or the reference)."""
    path = Path(path)
    if not path.suffix:
        path = path.with_suffix(".npz")
    with np.load(str(path), allow_pickle=False) as npz:
        raw = npz["__meta__"].item()
        meta = json.loads(raw.decode("utf-8") if isinstance(raw, bytes) else str(raw))
        data = {key[len("data_"):]: npz[key] for key in npz.files if key.startswith("data_")}
    return _rebuild(meta), data
