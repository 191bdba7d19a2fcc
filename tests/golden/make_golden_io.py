"""A grid file written by the REAL reference's save_grid (build container only): tests/golden/io_ref.npz."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import reference_loader  # noqa: E402

ref = reference_loader.load()
from numgrids import AxisType, SphericalGrid, create_axis, save_grid  # noqa: E402

if __name__ == "__main__":
    g = SphericalGrid(create_axis(AxisType.CHEBYSHEV, 6, 0.5, 2.0, name="r"),
                      create_axis(AxisType.CHEBYSHEV, 5, 0.3, 2.8),
                      create_axis(AxisType.EQUIDISTANT_PERIODIC, 8, 0.0, 2 * np.pi, name="phi"))
    R, T, P = g.meshed_coords
    save_grid(os.path.join(HERE, "io_ref.npz"), g, density=R * np.sin(T), phase=np.cos(P))
    print("io_ref.npz written")
