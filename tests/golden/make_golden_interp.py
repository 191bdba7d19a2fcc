"""Golden vectors of linear interpolation / multigrid transfer from the REAL reference (build container
only; conventions as in make_golden.py).  Asserts that oracle/interp_port.py reproduces the reference
(= SciPy's arithmetic) bit for bit, then writes interp_cases.npz."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))

from oracle import interp_port as ip, reference_loader  # noqa: E402
from tests_support import smooth_field  # noqa: E402

ref = reference_loader.load()
from numgrids import AxisType, Grid, Interpolator, MultiGrid, create_axis  # noqa: E402

KIND = {"equidistant": AxisType.EQUIDISTANT, "periodic": AxisType.EQUIDISTANT_PERIODIC,
        "chebyshev": AxisType.CHEBYSHEV, "log": AxisType.LOGARITHMIC}
# (source grid, target grid): same domains, different resolutions / axis types
PAIRS = [
    ([("equidistant", 17, 0.0, 1.0)], [("equidistant", 33, 0.0, 1.0)]),
    ([("chebyshev", 20, -1.0, 2.0)], [("equidistant", 11, -1.0, 2.0)]),
    ([("equidistant", 12, 0.0, 2.0), ("chebyshev", 9, -1.0, 1.0)], [("equidistant", 23, 0.0, 2.0), ("chebyshev", 17, -1.0, 1.0)]),
    ([("log", 14, 0.1, 10.0), ("equidistant", 13, 0.0, 1.0)], [("log", 7, 0.1, 10.0), ("equidistant", 7, 0.0, 1.0)]),
    ([("chebyshev", 9, 0.5, 2.0), ("equidistant", 10, 0.0, 1.0), ("equidistant", 8, -1.0, 1.0)],
     [("chebyshev", 17, 0.5, 2.0), ("equidistant", 19, 0.0, 1.0), ("equidistant", 15, -1.0, 1.0)]),
    ([("equidistant", 6, 0.0, 1.0), ("equidistant", 5, 0.0, 1.0), ("chebyshev", 7, 0.0, 1.0), ("equidistant", 4, 0.0, 2.0)],
     [("equidistant", 5, 0.0, 1.0), ("equidistant", 9, 0.0, 1.0), ("chebyshev", 4, 0.0, 1.0), ("equidistant", 7, 0.0, 2.0)]),
]
MULTIGRIDS = [
    [("equidistant", 33, 0.0, 1.0)],
    [("chebyshev", 17, 0.0, 1.0), ("equidistant", 20, -1.0, 1.0)],
    [("equidistant", 12, 0.0, 1.0), ("log", 11, 0.5, 4.0), ("chebyshev", 9, 0.0, 2.0)],
]


def axes_of(specs):
    return [create_axis(KIND[k], n, lo, hi) for k, n, lo, hi in specs]


def main():
    arrays, meta = {}, {"pairs": PAIRS, "multigrids": MULTIGRIDS, "points": [], "transfers": []}
    rng = np.random.default_rng(3)
    for pi, (src, dst) in enumerate(PAIRS):
        gs, gd = Grid(*axes_of(src)), Grid(*axes_of(dst))
        cs, cd = [a.coords for a in gs.axes], [a.coords for a in gd.axes]
        for name, f in (("smooth", smooth_field(cs)), ("noise", rng.standard_normal(gs.shape))):
            r = Interpolator(gs, f, method="linear")(gd)
            assert np.array_equal(r, ip.linear_on_grid(f, cs, cd)), (pi, name)
            arrays[f"p{pi}_{name}_f"] = f
            arrays[f"p{pi}_{name}_out"] = r
        # scattered points (the reference evaluates them one at a time)
        pts = np.stack([rng.uniform(c[0], c[-1], 9) for c in cs], axis=1)
        pts[0] = [c[0] for c in cs]
        pts[1] = [c[-1] for c in cs]
        pts[2] = [c[len(c) // 2] for c in cs]
        f = arrays[f"p{pi}_noise_f"]
        inter = Interpolator(gs, f, method="linear")
        r = inter([tuple(p) for p in pts])
        assert np.array_equal(r, ip.linear_points(f, cs, pts)), pi
        assert np.array_equal(inter(tuple(pts[3])), r[3])
        arrays[f"p{pi}_pts"] = pts
        arrays[f"p{pi}_pts_out"] = np.asarray(r)
        meta["points"].append(pi)
    for mi, specs in enumerate(MULTIGRIDS):
        mg = MultiGrid(*axes_of(specs), min_size=3)
        meta["transfers"].append({"mg": mi, "shapes": [list(g.shape) for g in mg.levels]})
        f = smooth_field([a.coords for a in mg.levels[0].axes])
        arrays[f"m{mi}_l0"] = f
        for lv in range(len(mg.levels) - 1):                       # coarsen all the way down ...
            f2 = mg.transfer(f, lv, lv + 1)
            assert np.array_equal(f2, ip.linear_on_grid(f, [a.coords for a in mg.levels[lv].axes],
                                                        [a.coords for a in mg.levels[lv + 1].axes]))
            arrays[f"m{mi}_down{lv + 1}"] = f2
            f = f2
        for lv in range(len(mg.levels) - 1, 0, -1):                # ... and refine back up
            f2 = mg.transfer(f, lv, lv - 1)
            arrays[f"m{mi}_up{lv - 1}"] = f2
            f = f2
    arrays["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "interp_cases.npz"), **arrays)
    return len(PAIRS) * 3 + sum(2 * (len(t["shapes"]) - 1) for t in meta["transfers"])


if __name__ == "__main__":
    print("interpolation cases:", main())
    print(f"  interp_cases.npz: {os.path.getsize(os.path.join(HERE, 'interp_cases.npz')) / 1024:.0f} KiB")
