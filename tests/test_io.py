"""save_grid / load_grid: the reference's .npz layout (numgrids/io.py), both directions."""
import os

import numpy as np
import pytest

import numgrids_b200 as ng
from conftest import GOLDEN
from oracle import reference_loader

A = ng.AxisType


def _grids():
    yield ng.Grid(ng.create_axis(A.EQUIDISTANT, 7, -1.0, 2.0, name="x"), ng.create_axis(A.LOGARITHMIC, 6, 0.1, 10.0))
    yield ng.Grid(ng.create_axis(A.EQUIDISTANT_PERIODIC, 8, 0.0, 1.0))
    yield ng.SphericalGrid(ng.create_axis(A.CHEBYSHEV, 5, 0.5, 2.0), ng.create_axis(A.CHEBYSHEV, 4, 0.3, 2.8),
                           ng.create_axis(A.EQUIDISTANT_PERIODIC, 6, 0.0, 2 * np.pi, name="phi"))
    yield ng.CylindricalGrid(ng.create_axis(A.CHEBYSHEV, 5, 0.5, 2.0),
                             ng.create_axis(A.EQUIDISTANT_PERIODIC, 6, 0.0, 2 * np.pi), ng.create_axis(A.EQUIDISTANT, 4, 0.0, 1.0))
    yield ng.PolarGrid(ng.create_axis(A.CHEBYSHEV, 5, 0.5, 2.0), ng.create_axis(A.EQUIDISTANT_PERIODIC, 6, 0.0, 2 * np.pi))


def _same_grid(a, b):
    assert type(a).__name__ == type(b).__name__ and a.shape == b.shape
    for x, y in zip(a.axes, b.axes):
        assert type(x).__name__ == type(y).__name__ and x.periodic == y.periodic and x.name == y.name
        assert np.allclose(x.coords, y.coords, rtol=1e-14, atol=0)


def test_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    for k, g in enumerate(_grids()):
        f, h = rng.standard_normal(g.shape), rng.standard_normal(g.shape)
        path = tmp_path / f"g{k}"                       # NumPy appends .npz; load_grid accepts the bare name
        ng.save_grid(path, g, f=f, other=h)
        g2, data = ng.load_grid(path)
        _same_grid(g, g2)
        assert sorted(data) == ["f", "other"] and np.array_equal(data["f"], f) and np.array_equal(data["other"], h)
    with pytest.raises(ValueError, match="expected grid shape"):
        ng.save_grid(tmp_path / "bad", g, f=np.zeros((2, 2)))


def test_loads_file_written_by_the_reference():
    g, data = ng.load_grid(os.path.join(GOLDEN, "io_ref.npz"))
    assert type(g).__name__ == "SphericalGrid" and g.shape == (6, 5, 8)
    assert g.get_axis(0).name == "r" and g.get_axis(2).periodic and g.get_axis(2).name == "phi"
    R, T, P = g.meshed_coords
    assert np.allclose(data["density"], R * np.sin(T), rtol=1e-14) and np.allclose(data["phase"], np.cos(P), rtol=1e-14)


@pytest.mark.skipif(not reference_loader.available(), reason="reference sources only exist in the build container")
def test_reference_loads_our_files(tmp_path):
    ref = reference_loader.load()
    for k, g in enumerate(_grids()):
        f = np.arange(g.size, dtype=np.float64).reshape(g.shape)
        ng.save_grid(tmp_path / f"h{k}.npz", g, field=f)
        gr, data = ref.load_grid(tmp_path / f"h{k}.npz")
        assert type(gr).__name__ == type(g).__name__ and gr.shape == g.shape
        for x, y in zip(g.axes, gr.axes):
            # (axes are rebuilt from coords[0] / coords[-1], so e.g. a log axis comes back to rounding, in either package)
            assert np.allclose(x.coords, y.coords, rtol=1e-14, atol=0) and x.periodic == y.periodic and x.name == y.name
        assert np.array_equal(data["field"], f)
