"""Boundary conditions: the reference's numgrids/boundary.py semantics.

CPU part: the NumPy restatement (oracle/boundary_port.py) against the golden vectors of the REAL
reference (tests/golden/bc_cases.npz, made by make_golden_bc.py), and the host-side
``apply_to_system`` assembly + the exceptions of the reference's tests/test_boundary.py.
GPU part: ``DirichletBC.apply`` / ``NeumannBC.apply`` through the C ABI (``ng_bc_apply``), bit for bit.
"""
import numpy as np
import pytest
import scipy.sparse

import numgrids_b200 as ng
from conftest import load_golden
from oracle import boundary_port as bp
from tests_support import build_axes

SIDES = ("low", "high")


def callable_value(face_coords):       # must match tests/golden/make_golden_bc.py
    return sum((i + 1.5) * np.cos(c + 0.3 * i) for i, c in enumerate(face_coords)) + 0.25


def _value(z, case, grid):
    if case["value"] == "scalar":
        return 1.75
    if case["value"] == "callable":
        return callable_value
    shape = tuple(s for i, s in enumerate(grid.shape) if i != case["axis"]) or (1,)
    return z[case["key"] + "_g"].reshape(shape)


# ------------------------------------------------------------------------------------------ CPU
def test_oracle_matches_reference_golden():
    z, meta = load_golden("bc_cases.npz")
    assert len(meta["cases"]) >= 40
    for case in meta["cases"]:
        specs = meta["grids"][case["grid"]]
        grid = ng.Grid(*build_axes(ng, specs))
        coords = [a.coords for a in grid.axes]
        u0 = z[f"g{case['grid']}_u"]
        g = bp.face_values(_value(z, case, grid), coords, case["axis"], case["side"])
        assert np.array_equal(g, z[case["key"] + "_g"]), case
        assert np.array_equal(bp.dirichlet_apply(u0, case["axis"], case["side"], g), z[case["key"] + "_dirichlet"]), case
        assert np.array_equal(bp.neumann_apply(u0, case["axis"], case["side"], g, coords[case["axis"]]),
                              z[case["key"] + "_neumann"]), case


def test_resolve_value_and_faces_match_reference():
    z, meta = load_golden("bc_cases.npz")
    for case in meta["cases"]:
        grid = ng.Grid(*build_axes(ng, meta["grids"][case["grid"]]))
        face = grid.faces[f"{case['axis']}_{case['side']}"]
        bc = ng.DirichletBC(face, _value(z, case, grid))
        assert np.array_equal(bc._resolve_value(), z[case["key"] + "_g"]), case
        assert np.array_equal(face.flat_indices, np.flatnonzero(face.mask))
        assert face.normal_sign == (-1 if case["side"] == "low" else 1)


def test_apply_to_system_matches_reference_golden():
    z, meta = load_golden("bc_cases.npz")
    n_checked = 0
    for gi, specs in enumerate(meta["grids"]):
        grid = ng.Grid(*build_axes(ng, specs))
        for axis in range(grid.ndims):
            for side in SIDES:
                key = f"g{gi}_a{axis}_{side}_sys"
                if key + "_L0" not in z.files:
                    continue
                face = ng.BoundaryFace(grid, axis, side)
                L0 = scipy.sparse.csc_matrix(z[key + "_L0"])
                rhs0, garr = z[key + "_rhs0"], z[key + "_g"]
                for name, bc in (("dirichlet", ng.DirichletBC(face, garr)),
                                 ("neumann", ng.NeumannBC(face, garr)),
                                 ("robin", ng.RobinBC(face, 2.0, 0.5, garr))):
                    L, rhs = bc.apply_to_system(L0, rhs0)
                    assert scipy.sparse.isspmatrix_csc(L) or L.format == "csc"
                    assert np.array_equal(L.toarray(), z[f"{key}_{name}_L"]), (key, name)
                    assert np.array_equal(rhs, z[f"{key}_{name}_rhs"]), (key, name)
                    assert np.array_equal(rhs0, z[key + "_rhs0"])          # inputs not mutated
                    n_checked += 1
    assert n_checked >= 12


def test_apply_bcs_last_wins_on_corners():
    ax = ng.create_axis(ng.AxisType.EQUIDISTANT, 8, 0.0, 1.0)
    grid = ng.Grid(ax, ax)
    L0 = sum(ng.Diff(grid, 2, i).as_matrix() for i in range(2))
    rhs0 = np.zeros(grid.size)
    bcs = [ng.DirichletBC(grid.faces["0_low"], 1.0), ng.DirichletBC(grid.faces["1_low"], 2.0)]
    L, rhs = ng.apply_bcs(bcs, L0, rhs0)
    assert rhs.reshape(grid.shape)[0, 0] == 2.0 and rhs.reshape(grid.shape)[0, 3] == 1.0
    Ld = L.toarray()
    for r in np.concatenate([f.flat_indices for f in (grid.faces["0_low"], grid.faces["1_low"])]):
        row = np.zeros(grid.size)
        row[r] = 1.0
        assert np.array_equal(Ld[r], row)


def test_face_and_bc_errors_like_reference():
    ax = ng.create_axis(ng.AxisType.EQUIDISTANT, 8, 0.0, 1.0)
    per = ng.create_axis(ng.AxisType.EQUIDISTANT_PERIODIC, 8, 0.0, 1.0)
    grid = ng.Grid(ax, per)
    with pytest.raises(ValueError, match="side must be"):
        ng.BoundaryFace(grid, 0, "left")
    with pytest.raises(ValueError, match="out of range"):
        ng.BoundaryFace(grid, 2, "low")
    with pytest.raises(ValueError, match="periodic"):
        ng.BoundaryFace(grid, 1, "low")
    assert sorted(grid.faces) == ["0_high", "0_low"] and grid.faces is grid.faces
    assert repr(grid.faces["0_low"]) == "BoundaryFace(axis=0, side='low')"
    with pytest.raises(NotImplementedError):
        ng.RobinBC(grid.faces["0_low"], 1.0, 1.0).apply(np.zeros(grid.shape))
    with pytest.raises(NotImplementedError):
        ng.boundary.BoundaryCondition(grid.faces["0_low"]).apply(np.zeros(grid.shape))


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_apply_golden_host_arrays(gpu_ready):
    z, meta = load_golden("bc_cases.npz")
    for case in meta["cases"]:
        grid = ng.Grid(*build_axes(ng, meta["grids"][case["grid"]]))
        face = ng.BoundaryFace(grid, case["axis"], case["side"])
        value = _value(z, case, grid)
        for cls, ref in ((ng.DirichletBC, "_dirichlet"), (ng.NeumannBC, "_neumann")):
            u = z[f"g{case['grid']}_u"].copy()
            out = cls(face, value).apply(u)
            assert out is u                                           # in place, like the reference
            assert np.array_equal(u, z[case["key"] + ref]), (case, cls.__name__)


@pytest.mark.gpu
def test_apply_golden_device_arrays(gpu_ready):
    import torch
    z, meta = load_golden("bc_cases.npz")
    for case in meta["cases"]:
        grid = ng.Grid(*build_axes(ng, meta["grids"][case["grid"]]))
        face = ng.BoundaryFace(grid, case["axis"], case["side"])
        value = _value(z, case, grid)
        if isinstance(value, np.ndarray) and case["axis"] % 2:
            value = torch.from_numpy(value).cuda()                    # device-resident face values
        for cls, ref in ((ng.DirichletBC, "_dirichlet"), (ng.NeumannBC, "_neumann")):
            u = torch.from_numpy(z[f"g{case['grid']}_u"]).cuda()
            ptr = u.data_ptr()
            out = cls(face, value).apply(u)
            assert out is u and u.data_ptr() == ptr
            assert np.array_equal(u.cpu().numpy(), z[case["key"] + ref]), (case, cls.__name__)


@pytest.mark.gpu
def test_heat_step_stays_on_device(gpu_ready):
    """A few explicit heat steps, laplacian -> update -> BCs, device-resident, against the oracle."""
    import torch
    from oracle import pencil
    ax = ng.create_axis(ng.AxisType.EQUIDISTANT, 48, 0.0, 1.0)
    ay = ng.create_axis(ng.AxisType.EQUIDISTANT, 40, 0.0, 2.0)
    grid = ng.Grid(ax, ay)
    X, Y = grid.meshed_coords
    u_ref = np.sin(np.pi * X) * np.cos(Y) + 0.1 * X
    u = torch.from_numpy(u_ref.copy()).cuda()
    bcs = [ng.DirichletBC(grid.faces["0_low"], 0.0), ng.DirichletBC(grid.faces["0_high"], lambda c: 0.1 + 0 * c[1]),
           ng.NeumannBC(grid.faces["1_low"], 0.0), ng.NeumannBC(grid.faces["1_high"], 0.5)]
    dxx, dyy = ng.Diff(grid, 2, 0), ng.Diff(grid, 2, 1)
    coords = [a.coords for a in grid.axes]
    dt = 1e-5
    for _ in range(5):
        u = u + dt * (dxx(u) + dyy(u))
        for bc in bcs:
            u = bc.apply(u)
        lap = pencil.fd_apply(u_ref, 0, coords[0][1] - coords[0][0], 2, 4) + \
            pencil.fd_apply(u_ref, 1, coords[1][1] - coords[1][0], 2, 4)
        u_ref = u_ref + dt * lap
        for bc in bcs:
            g = bp.face_values(bc._value_spec, coords, bc.face.axis_index, bc.face.side)
            fn = bp.dirichlet_apply if isinstance(bc, ng.DirichletBC) else None
            if fn is not None:
                u_ref = fn(u_ref, bc.face.axis_index, bc.face.side, g)
            else:
                u_ref = bp.neumann_apply(u_ref, bc.face.axis_index, bc.face.side, g, coords[bc.face.axis_index])
    assert np.array_equal(u.cpu().numpy(), u_ref)


@pytest.mark.gpu
def test_apply_accepts_what_the_reference_accepts(gpu_ready):
    """`u[mask] = g` in the reference takes any ndarray of the grid's shape (other float dtypes, non-contiguous views,
    complex) and broadcasts a one-element value (numgrids/boundary.py:157-166)."""
    import numgrids_b200 as ng
    from numgrids_b200.boundary import DirichletBC, NeumannBC
    ax = ng.create_axis(ng.AxisType.EQUIDISTANT, 9, 0.0, 1.0)
    ay = ng.create_axis(ng.AxisType.EQUIDISTANT, 7, 0.0, 2.0)
    g = ng.Grid(ax, ay)
    rng = np.random.default_rng(3)
    base = rng.standard_normal(g.shape)
    face = g.faces["0_low"]
    # float32
    u32 = base.astype(np.float32)
    DirichletBC(face, lambda c: 2.5).apply(u32)                  # callable returning ONE value: broadcast
    assert u32.dtype == np.float32 and np.all(u32[0] == np.float32(2.5)) and np.array_equal(u32[1:], base.astype(np.float32)[1:])
    # non-contiguous view of a larger array
    big = rng.standard_normal((9, 14))
    view = big[:, ::2]
    want = view.copy()
    h = ax.coords[1] - ax.coords[0]
    want[0] = want[1] + (h * -1.0) * 0.75
    NeumannBC(face, 0.75).apply(view)
    assert np.array_equal(view, want)
    # complex Dirichlet
    uc = base + 1j * rng.standard_normal(g.shape)
    keep = uc.copy()
    DirichletBC(g.faces["1_high"], 1.5).apply(uc)
    assert np.all(uc[:, -1] == 1.5 + 0j) and np.array_equal(uc[:, :-1], keep[:, :-1])
