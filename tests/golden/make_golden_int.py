"""Golden vectors of ``Integral`` from the REAL reference (build container only; conventions as in
make_golden.py).  Asserts that oracle/integral_port.py reproduces the reference bit for bit and that
the per-axis weight vectors equal the ones the reference's own inverse matrices hold, then writes
int_cases.npz."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))

from oracle import integral_port as ip, reference_loader  # noqa: E402
from tests_support import smooth_field  # noqa: E402

ref = reference_loader.load()
from numgrids import AxisType, Diff, Grid, create_axis  # noqa: E402
from numgrids.integration import Integral  # noqa: E402

KIND = {"equidistant": AxisType.EQUIDISTANT, "periodic": AxisType.EQUIDISTANT_PERIODIC,
        "chebyshev": AxisType.CHEBYSHEV, "log": AxisType.LOGARITHMIC}
GRIDS = [
    [("equidistant", 33, 0.0, 1.0)],
    [("chebyshev", 24, -1.0, 2.0)],
    [("periodic", 16, 0.0, 2 * np.pi)],
    [("log", 40, 0.1, 10.0)],
    [("equidistant", 21, 0.0, 2.0), ("chebyshev", 18, -1.0, 1.0)],
    [("log", 14, 0.1, 10.0), ("equidistant", 13, 0.0, 1.0)],
    [("chebyshev", 12, 0.5, 2.0), ("chebyshev", 11, 0.3, 2.8), ("periodic", 16, 0.0, 2 * np.pi)],
    [("periodic", 8, 0.0, 1.0), ("equidistant", 17, 0.0, 1.0), ("chebyshev", 10, 0.0, 1.0)],
    [("equidistant", 9, 0.0, 1.0), ("periodic", 6, 0.0, 1.0), ("chebyshev", 7, 0.0, 1.0), ("equidistant", 12, -1.0, 1.0)],
]


def main():
    arrays, meta = {}, []
    rng = np.random.default_rng(11)
    for gi, specs in enumerate(GRIDS):
        axes = [create_axis(KIND[k], n, lo, hi) for k, n, lo, hi in specs]
        grid = Grid(*axes)
        I = Integral(grid)
        weights, periodic = [], []
        for i, ax in enumerate(axes):
            per = bool(ax.periodic)
            D = None if per else Diff(Grid(ax), 1, 0).as_matrix()
            w = ip.axis_weights(D, per, ax.spacing if per else None, len(ax))
            if not per:      # the reference's own stored inverse holds the same last row
                assert np.array_equal(w[1:], np.asarray(I.D_invs[i].toarray())[-1]), (gi, i)
            weights.append(w)
            periodic.append(per)
            arrays[f"g{gi}_w{i}"] = w
        coords = [a.coords for a in axes]
        for name, f in (("smooth", smooth_field(coords)), ("noise", rng.standard_normal(grid.shape))):
            r = I(f)
            assert np.array_equal(np.float64(r), ip.integral(f, weights, periodic)), (gi, name)
            arrays[f"g{gi}_{name}_f"] = f
            arrays[f"g{gi}_{name}_I"] = np.float64(r)
            meta.append({"grid": gi, "field": name})
    arrays["__meta__"] = np.frombuffer(json.dumps({"grids": GRIDS, "cases": meta}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "int_cases.npz"), **arrays)
    return len(meta)


if __name__ == "__main__":
    print("integral cases:", main())
    print(f"  int_cases.npz: {os.path.getsize(os.path.join(HERE, 'int_cases.npz')) / 1024:.0f} KiB")
