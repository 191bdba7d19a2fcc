"""Golden vectors of the spline Interpolator methods (cubic; quadratic in 1-D) from the REAL reference (build
container only).  Asserts that the host tables of numgrids_b200.interpol fed through oracle/interp_port.py's
restatement of SciPy's evaluation order reproduce the reference bit for bit, then writes spline_cases.npz."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))

from numgrids_b200 import interpol  # noqa: E402  (host tables only: no GPU needed)
from oracle import interp_port as ip, reference_loader  # noqa: E402
from tests_support import smooth_field  # noqa: E402

ref = reference_loader.load()
from numgrids import AxisType, Grid, Interpolator, create_axis  # noqa: E402

KIND = {"equidistant": AxisType.EQUIDISTANT, "chebyshev": AxisType.CHEBYSHEV, "log": AxisType.LOGARITHMIC}
CASES = [
    ("cubic", [("equidistant", 17, 0.0, 1.0)], [("equidistant", 40, 0.0, 1.0)]),
    ("quadratic", [("chebyshev", 14, -1.0, 2.0)], [("equidistant", 23, -1.0, 2.0)]),
    ("cubic", [("equidistant", 12, 0.0, 2.0), ("chebyshev", 9, -1.0, 1.0)], [("equidistant", 23, 0.0, 2.0), ("chebyshev", 17, -1.0, 1.0)]),
    ("cubic", [("log", 10, 0.5, 4.0), ("equidistant", 8, 0.0, 1.0), ("chebyshev", 7, 0.0, 1.0)],
     [("log", 13, 0.5, 4.0), ("equidistant", 11, 0.0, 1.0), ("chebyshev", 9, 0.0, 1.0)]),
]


def axes_of(specs):
    return [create_axis(KIND[k], n, lo, hi) for k, n, lo, hi in specs]


def main():
    arrays, rng = {}, np.random.default_rng(4)
    for ci, (method, src, dst) in enumerate(CASES):
        gs, gd = Grid(*axes_of(src)), Grid(*axes_of(dst))
        cs, cd = [a.coords for a in gs.axes], [a.coords for a in gd.axes]
        k = {"cubic": 3, "quadratic": 2}[method]
        for name, f in (("smooth", smooth_field(cs)), ("noise", rng.standard_normal(gs.shape))):
            r = Interpolator(gs, f, method=method)(gd)
            knots, coef = interpol.spline_fit(cs, f, k)
            fi, bv = interpol.spline_tables(knots, k, cs, cd)
            assert np.array_equal(r, ip.bspline_from_tables(coef, fi, bv)), (ci, name)
            arrays[f"c{ci}_{name}_f"] = f
            arrays[f"c{ci}_{name}_out"] = r
        pts = np.stack([rng.uniform(c[0], c[-1], 7) for c in cs], axis=1)
        pts[0] = [c[0] for c in cs]
        pts[1] = [c[-1] for c in cs]
        f = arrays[f"c{ci}_noise_f"]
        rp = np.asarray(Interpolator(gs, f, method=method)([tuple(p) for p in pts]))
        knots, coef = interpol.spline_fit(cs, f, k)
        fi, bv = interpol.spline_tables(knots, k, cs, [pts[:, a] for a in range(len(cs))])
        assert np.array_equal(rp, ip.bspline_from_tables(coef, fi, bv, scattered=True)), ci
        arrays[f"c{ci}_pts"] = pts
        arrays[f"c{ci}_pts_out"] = rp
    arrays["__meta__"] = np.frombuffer(json.dumps({"cases": CASES}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "spline_cases.npz"), **arrays)
    return len(CASES)


if __name__ == "__main__":
    print("spline cases:", main())
    print(f"  spline_cases.npz: {os.path.getsize(os.path.join(HERE, 'spline_cases.npz')) / 1024:.0f} KiB")
