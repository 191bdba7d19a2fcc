"""Golden vectors of the array-level / system-level boundary conditions, from the REAL reference
(build container only; see make_golden.py for the conventions).  Asserts that
oracle/boundary_port.py reproduces the reference bit for bit, then writes bc_cases.npz."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import boundary_port as bp, reference_loader  # noqa: E402

ref = reference_loader.load()
from numgrids import AxisType, BoundaryFace, Diff, DirichletBC, Grid, NeumannBC, RobinBC, create_axis  # noqa: E402

KIND = {"equidistant": AxisType.EQUIDISTANT, "periodic": AxisType.EQUIDISTANT_PERIODIC,
        "chebyshev": AxisType.CHEBYSHEV, "log": AxisType.LOGARITHMIC}
GRIDS = [
    [("equidistant", 12, 0.0, 1.0)],
    [("chebyshev", 9, -1.0, 2.0)],
    [("equidistant", 10, 0.0, 2.0), ("chebyshev", 7, -1.0, 1.0)],
    [("log", 14, 0.1, 10.0), ("equidistant", 13, 0.0, 1.0)],
    [("chebyshev", 6, 0.0, 1.0), ("periodic", 8, 0.0, 2 * np.pi), ("equidistant", 11, 0.0, 1.0)],
]


def callable_value(face_coords):
    return sum((i + 1.5) * np.cos(c + 0.3 * i) for i, c in enumerate(face_coords)) + 0.25


def main():
    arrays, meta = {}, []
    rng = np.random.default_rng(5)
    for gi, specs in enumerate(GRIDS):
        axes = [create_axis(KIND[k], n, lo, hi) for k, n, lo, hi in specs]
        grid = Grid(*axes)
        coords = [a.coords for a in axes]
        u0 = rng.standard_normal(grid.shape)
        arrays[f"g{gi}_u"] = u0
        for axis, ax in enumerate(axes):
            if ax.periodic:
                continue
            for side in ("low", "high"):
                face = BoundaryFace(grid, axis, side)
                fshape = tuple(s for i, s in enumerate(grid.shape) if i != axis)
                garr = rng.standard_normal(fshape) if fshape else rng.standard_normal(1)
                for vname, value in (("scalar", 1.75), ("array", garr), ("callable", callable_value)):
                    key = f"g{gi}_a{axis}_{side}_{vname}"
                    g_or = bp.face_values(value, coords, axis, side)
                    d_ref = DirichletBC(face, value).apply(u0.copy())
                    n_ref = NeumannBC(face, value).apply(u0.copy())
                    assert np.array_equal(d_ref, bp.dirichlet_apply(u0, axis, side, g_or)), key
                    assert np.array_equal(n_ref, bp.neumann_apply(u0, axis, side, g_or, ax.coords)), key
                    arrays[key + "_g"] = g_or
                    arrays[key + "_dirichlet"] = d_ref
                    arrays[key + "_neumann"] = n_ref
                    meta.append({"grid": gi, "axis": axis, "side": side, "value": vname, "key": key})
                # system level (small grids only): identity / sign*D / a I + b sign D rows
                if grid.size <= 200:
                    L0 = sum(Diff(grid, 2, i).as_matrix() for i in range(grid.ndims))
                    rhs0 = rng.standard_normal(grid.size)
                    gv = bp.face_values(garr, coords, axis, side)
                    rows = bp.flat_face_indices(grid.shape, axis, side)
                    assert np.array_equal(rows, face.flat_indices)
                    D = Diff(grid, 1, axis).as_matrix().toarray()
                    sign = -1 if side == "low" else 1
                    eye = np.eye(grid.size)
                    for bname, bc, new_rows in (
                            ("dirichlet", DirichletBC(face, garr), eye[rows]),
                            ("neumann", NeumannBC(face, garr), sign * D[rows]),
                            ("robin", RobinBC(face, 2.0, 0.5, garr), 2.0 * eye[rows] + 0.5 * sign * D[rows])):
                        Lr, rr = bc.apply_to_system(L0, rhs0)
                        Lo, ro = bp.system_rows(L0.toarray(), rhs0, rows, new_rows, gv)
                        key = f"g{gi}_a{axis}_{side}_sys_{bname}"
                        assert np.array_equal(Lr.toarray(), Lo) and np.array_equal(rr, ro), key
                        arrays[key + "_L"] = Lr.toarray()
                        arrays[key + "_rhs"] = rr
                    arrays[f"g{gi}_a{axis}_{side}_sys_L0"] = L0.toarray()
                    arrays[f"g{gi}_a{axis}_{side}_sys_rhs0"] = rhs0
                    arrays[f"g{gi}_a{axis}_{side}_sys_g"] = garr
    arrays["__meta__"] = np.frombuffer(json.dumps({"grids": GRIDS, "cases": meta}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "bc_cases.npz"), **arrays)
    return len(meta)


if __name__ == "__main__":
    print("bc cases:", main())
    print(f"  bc_cases.npz: {os.path.getsize(os.path.join(HERE, 'bc_cases.npz')) / 1024:.0f} KiB")
