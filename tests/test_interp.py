"""Linear interpolation / multigrid transfer: numgrids/interpol.py + MultiGrid semantics.

CPU part: the NumPy restatement of SciPy's linear arithmetic (oracle/interp_port.py) against the golden
vectors of the REAL reference (tests/golden/interp_cases.npz, made by make_golden_interp.py), plus the
host logic (levels, tables, exceptions).  GPU part: ``Interpolator(method="linear")`` and
``MultiGrid.transfer`` through the C ABI (``ng_interp_linear``) -- bit for bit.
"""
import numpy as np
import pytest

import numgrids_b200 as ng
from conftest import load_golden
from oracle import interp_port as ip
from tests_support import build_axes


def _grids(meta, pi):
    src, dst = meta["pairs"][pi]
    return ng.Grid(*build_axes(ng, src)), ng.Grid(*build_axes(ng, dst))


def test_oracle_matches_reference_golden():
    z, meta = load_golden("interp_cases.npz")
    for pi in range(len(meta["pairs"])):
        gs, gd = _grids(meta, pi)
        cs, cd = [a.coords for a in gs.axes], [a.coords for a in gd.axes]
        for name in ("smooth", "noise"):
            assert np.array_equal(ip.linear_on_grid(z[f"p{pi}_{name}_f"], cs, cd), z[f"p{pi}_{name}_out"]), (pi, name)
        assert np.array_equal(ip.linear_points(z[f"p{pi}_noise_f"], cs, z[f"p{pi}_pts"]), z[f"p{pi}_pts_out"]), pi


def test_multigrid_levels_and_errors():
    z, meta = load_golden("interp_cases.npz")
    for t in meta["transfers"]:
        mg = ng.MultiGrid(*build_axes(ng, meta["multigrids"][t["mg"]]), min_size=3)
        assert [list(g.shape) for g in mg.levels] == t["shapes"]
    ax = ng.create_axis(ng.AxisType.EQUIDISTANT, 16, 0.0, 1.0)
    per = ng.create_axis(ng.AxisType.EQUIDISTANT_PERIODIC, 16, 0.0, 1.0, name="phi")
    mg = ng.MultiGrid(ax, per)
    assert [g.shape for g in mg.levels] == [(16, 16), (8, 8), (4, 4), (2, 2)]
    assert all(g.get_axis(1).periodic and g.get_axis(1).name == "phi" for g in mg.levels)
    with pytest.raises(ValueError, match="does not match grid shape"):
        mg.transfer(np.zeros((3, 3)), 0, 1)
    with pytest.raises(ValueError, match="adjacent levels"):
        mg.transfer(np.zeros((16, 16)), 0, 2)
    g = ng.Grid(ax, ax)
    with pytest.raises(ValueError, match="not defined"):
        ng.Interpolator(g, np.zeros(g.shape), method="septic")
    with pytest.raises(ValueError, match="not defined"):
        ng.Interpolator(g, np.zeros(g.shape), method="quadratic")     # RegularGridInterpolator has no such method either
    tiny = ng.Grid(ng.create_axis(ng.AxisType.EQUIDISTANT, 3, 0.0, 1.0), ax)
    with pytest.raises(ValueError, match="requires at least"):
        ng.Interpolator(tiny, np.zeros(tiny.shape), method="cubic")
    with pytest.raises(ValueError, match="out of bounds in dimension 1"):
        ng.interpol.axis_table(np.linspace(0, 1, 5), np.array([0.5, 1.5]), 1)
    i, x0, x1 = ng.interpol.axis_table(np.linspace(0, 1, 5), np.array([0.0, 0.25, 0.3, 1.0]), 0)
    assert list(i) == [0, 1, 1, 3]


@pytest.mark.gpu
def test_interpolate_onto_grid_golden(gpu_ready):
    import torch
    z, meta = load_golden("interp_cases.npz")
    for pi in range(len(meta["pairs"])):
        gs, gd = _grids(meta, pi)
        for name in ("smooth", "noise"):
            f = z[f"p{pi}_{name}_f"]
            out = ng.Interpolator(gs, f, method="linear")(gd)
            assert isinstance(out, np.ndarray) and out.shape == gd.shape
            assert np.array_equal(out, z[f"p{pi}_{name}_out"]), (pi, name)
            out_d = ng.Interpolator(gs, torch.from_numpy(f).cuda(), method="linear")(gd)
            assert isinstance(out_d, torch.Tensor) and out_d.is_cuda
            assert np.array_equal(out_d.cpu().numpy(), z[f"p{pi}_{name}_out"]), (pi, name)


@pytest.mark.gpu
def test_interpolate_points_golden(gpu_ready):
    z, meta = load_golden("interp_cases.npz")
    for pi in meta["points"]:
        gs, _ = _grids(meta, pi)
        inter = ng.Interpolator(gs, z[f"p{pi}_noise_f"], method="linear")
        pts = z[f"p{pi}_pts"]
        out = inter([tuple(p) for p in pts])
        assert np.array_equal(out, z[f"p{pi}_pts_out"]), pi
        assert inter(tuple(pts[3])) == z[f"p{pi}_pts_out"][3]
        assert np.array_equal(inter(zip(*[pts[:, a] for a in range(pts.shape[1])])), z[f"p{pi}_pts_out"])
        with pytest.raises(ValueError, match="out of bounds"):
            inter(tuple(c[-1] + 1.0 for c in (a.coords for a in gs.axes)))


@pytest.mark.gpu
def test_multigrid_transfer_golden_device_resident(gpu_ready):
    import torch
    z, meta = load_golden("interp_cases.npz")
    for t in meta["transfers"]:
        mi = t["mg"]
        mg = ng.MultiGrid(*build_axes(ng, meta["multigrids"][mi]), min_size=3)
        f = torch.from_numpy(z[f"m{mi}_l0"]).cuda()
        for lv in range(len(mg.levels) - 1):
            f = mg.transfer(f, lv, lv + 1)
            assert f.is_cuda and np.array_equal(f.cpu().numpy(), z[f"m{mi}_down{lv + 1}"]), (mi, lv)
        for lv in range(len(mg.levels) - 1, 0, -1):
            f = mg.transfer(f, lv, lv - 1)
            assert f.is_cuda and np.array_equal(f.cpu().numpy(), z[f"m{mi}_up{lv - 1}"]), (mi, lv)


@pytest.mark.gpu
def test_transfer_large_against_oracle(gpu_ready):
    """256^3 -> 128^3 -> 256^3 on the device against the restatement (bit for bit)."""
    import torch
    ax = ng.create_axis(ng.AxisType.EQUIDISTANT, 256, 0.0, 1.0)
    ch = ng.create_axis(ng.AxisType.CHEBYSHEV, 256, 0.0, 1.0)
    mg = ng.MultiGrid(ax, ch, ax, min_size=64)
    X, Y, Z = mg.levels[0].meshed_coords
    f = np.sin(3 * X) * np.cos(2 * Y) * np.exp(-Z)
    c0, c1 = [a.coords for a in mg.levels[0].axes], [a.coords for a in mg.levels[1].axes]
    down = mg.transfer(torch.from_numpy(f).cuda(), 0, 1)
    ref_down = ip.linear_on_grid(f, c0, c1)
    assert np.array_equal(down.cpu().numpy(), ref_down)
    up = mg.transfer(down, 1, 0)
    assert np.array_equal(up.cpu().numpy(), ip.linear_on_grid(ref_down, c1, c0))


# ------------------------------------------------------------------------------------------ spline methods
def _spline_cases():
    z, meta = load_golden("spline_cases.npz")
    for ci, (method, src, dst) in enumerate(meta["cases"]):
        yield z, ci, method, ng.Grid(*build_axes(ng, src)), ng.Grid(*build_axes(ng, dst))


def test_spline_tables_reproduce_reference_golden():
    """Host side of the cubic / quadratic methods (SciPy's fit + basis tables) through the restated evaluation
    order against the reference's results.  The fit is iterative in N-D (gcrotmk): equal to rounding across hosts."""
    from numgrids_b200 import interpol
    for z, ci, method, gs, gd in _spline_cases():
        k = {"cubic": 3, "quadratic": 2}[method]
        cs, cd = [a.coords for a in gs.axes], [a.coords for a in gd.axes]
        for name in ("smooth", "noise"):
            f, ref = z[f"c{ci}_{name}_f"], z[f"c{ci}_{name}_out"]
            knots, coef = interpol.spline_fit(cs, f, k)
            fi, bv = interpol.spline_tables(knots, k, cs, cd)
            out = ip.bspline_from_tables(coef, fi, bv)
            assert np.max(np.abs(out - ref)) <= 1e-9 * np.max(np.abs(ref)), (ci, name)
    with pytest.raises(ValueError, match="out of bounds in dimension 0"):
        interpol.spline_tables(knots, k, cs, [np.array([c[-1] + 1.0]) for c in cs])


@pytest.mark.gpu
def test_spline_interpolator_golden_and_scipy_order(gpu_ready):
    """Device evaluation: equal to the golden vectors of the real reference (to the fit's portability) and
    bit-identical to the restatement of SciPy's evaluation order fed with the same host tables."""
    import torch
    from numgrids_b200 import interpol
    for z, ci, method, gs, gd in _spline_cases():
        k = {"cubic": 3, "quadratic": 2}[method]
        cs, cd = [a.coords for a in gs.axes], [a.coords for a in gd.axes]
        for name in ("smooth", "noise"):
            f, ref = z[f"c{ci}_{name}_f"], z[f"c{ci}_{name}_out"]
            out = ng.Interpolator(gs, f, method=method)(gd)
            assert isinstance(out, np.ndarray) and out.shape == gd.shape
            assert np.max(np.abs(out - ref)) <= 1e-9 * np.max(np.abs(ref)), (ci, name)
            knots, coef = interpol.spline_fit(cs, f, k)
            assert np.array_equal(out, ip.bspline_from_tables(coef, *interpol.spline_tables(knots, k, cs, cd))), (ci, name)
            out_d = ng.Interpolator(gs, torch.from_numpy(f).cuda(), method=method)(gd)
            assert out_d.is_cuda and np.array_equal(out_d.cpu().numpy(), out)
        f, pts, refp = z[f"c{ci}_noise_f"], z[f"c{ci}_pts"], z[f"c{ci}_pts_out"]
        inter = ng.Interpolator(gs, f, method=method)
        outp = inter([tuple(p) for p in pts])
        assert np.max(np.abs(outp - refp)) <= 1e-9 * np.max(np.abs(refp)), ci
        assert inter(tuple(pts[2])) == outp[2]
        if method == "cubic":
            assert ng.interpolate(gs, f, tuple(pts[3])) == outp[3]          # the reference's helper: always cubic
        with pytest.raises(ValueError, match="out of bounds"):
            inter(tuple(c[-1] + 1.0 for c in cs))
