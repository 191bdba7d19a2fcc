"""Golden histories of the reference's adaptive refinement loop (build container only; conventions as in
make_golden.py): ErrorEstimator / adapt / estimate_error of the REAL reference on analytic fields, stored
as JSON inside amr_cases.npz (shapes, per-axis and global errors per iteration)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))

from oracle import reference_loader  # noqa: E402
from tests_support import amr_field  # noqa: E402

ref = reference_loader.load()
from numgrids import AxisType, Grid, adapt, create_axis, estimate_error  # noqa: E402

KIND = {"equidistant": AxisType.EQUIDISTANT, "chebyshev": AxisType.CHEBYSHEV, "log": AxisType.LOGARITHMIC}
CASES = [
    {"axes": [("equidistant", 10, 0.0, 1.0)], "field": "poly5", "norm": "max", "tol": 1e-4, "kw": {}},
    {"axes": [("equidistant", 10, 0.0, 1.0), ("equidistant", 10, 0.0, 1.0)], "field": "aniso", "norm": "max", "tol": 1e-4, "kw": {}},
    {"axes": [("chebyshev", 8, 0.0, 1.0), ("equidistant", 12, 0.0, 2.0)], "field": "aniso", "norm": "l2", "tol": 1e-5, "kw": {"refine_all": True}},
    {"axes": [("equidistant", 9, 0.0, 1.0), ("log", 8, 0.5, 4.0), ("chebyshev", 7, -1.0, 1.0)], "field": "wave3", "norm": "mean", "tol": 1e-4,
     "kw": {"max_iterations": 6, "max_points_per_axis": 40, "refinement_factor": 1.5}},
    {"axes": [("equidistant", 16, 0.0, 1.0), ("equidistant", 16, 0.0, 1.0)], "field": "aniso", "norm": "max", "tol": 1e-12, "kw": {"max_iterations": 3}},
]


def main():
    out = []
    for case in CASES:
        grid = Grid(*[create_axis(KIND[k], n, lo, hi) for k, n, lo, hi in case["axes"]])
        func = lambda g, name=case["field"]: amr_field(name, g.meshed_coords)
        res = adapt(grid, func, tol=case["tol"], norm=case["norm"], **case["kw"])
        est = estimate_error(grid, func, norm=case["norm"])
        out.append({
            "case": case,
            "final_shape": list(res.grid.shape), "iterations": res.iterations, "converged": bool(res.converged),
            "global_error": res.global_error, "errors": {str(k): v for k, v in res.errors.items()},
            "history": [{"shape": list(h["shape"]), "global_error": h["global_error"],
                         "axis_errors": {str(k): v for k, v in h["axis_errors"].items()},
                         "refined": sorted(int(k.rsplit("_", 1)[1]) for k in h if k.startswith("refined_axis_"))}
                        for h in res.history],
            "estimate": {"global": est["global"], "per_axis": {str(k): v for k, v in est["per_axis"].items()}},
            "f_checksum": float(np.sum(res.f)),
        })
    np.savez_compressed(os.path.join(HERE, "amr_cases.npz"),
                        __meta__=np.frombuffer(json.dumps(out).encode(), dtype=np.uint8))
    return out


if __name__ == "__main__":
    for o in main():
        print(o["case"]["field"], o["case"]["norm"], "->", o["final_shape"], o["iterations"], o["converged"], f"{o['global_error']:.3e}")
