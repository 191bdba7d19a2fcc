"""CPU: the pencil oracle (oracle/pencil.py) reproduces every golden vector bit for bit.

The golden vectors were produced by the REAL reference in the build container
(tests/golden/make_golden.py); this test pins the oracle wherever the suite runs.
"""
import numpy as np
import pytest

from conftest import load_golden
from oracle import pencil


def _axes(specs):
    return [pencil.make_axis(*s) for s in specs]


def test_diff_cases_bit_exact():
    z, meta = load_golden("diff_cases.npz")
    assert len(meta) > 100
    for m in meta:
        paxes = _axes(m["axes"])
        out = pencil.make_diff(paxes, m["order"], m["axis"], m["acc"])(z[m["f"]])
        assert np.array_equal(out, z[m["id"]]), m


def test_curvilinear_cases_bit_exact():
    z, meta = load_golden("curv_cases.npz")
    from tests_support import scale_factor_arrays
    for m in meta:
        paxes = _axes(m["axes"])
        nd = m["ndim"]
        D = [pencil.make_diff(paxes, 1, i, 4) for i in range(nd)]
        h = scale_factor_arrays(m["grid"], [a["coords"] for a in paxes])
        cid = m["id"]
        f = z[f"{cid}_f"]
        v = [z[f"{cid}_v{i}"] for i in range(nd)]
        assert np.array_equal(pencil.laplacian(f, h, D), z[f"{cid}_lap"]), (cid, "lap")
        assert np.array_equal(pencil.divergence(v, h, D), z[f"{cid}_div"]), (cid, "div")
        g = pencil.gradient(f, h, D)
        for i in range(nd):
            assert np.array_equal(g[i], z[f"{cid}_grad{i}"]), (cid, "grad", i)
        c = pencil.curl(v, h, D)
        c = (c,) if nd == 2 else c
        for i in range(len(c)):
            assert np.array_equal(c[i], z[f"{cid}_curl{i}"]), (cid, "curl", i)


def test_cartesian_laplacian_bit_exact():
    z, meta = load_golden("cartlap_cases.npz")
    for m in meta:
        paxes = _axes(m["axes"])
        hs = [a["coords"][1] - a["coords"][0] for a in paxes]
        for acc in (2, 4, 6):
            out = pencil.cartesian_laplacian(z[f"{m['id']}_f"], hs, acc)
            assert np.array_equal(out, z[f"{m['id']}_acc{acc}"])


def test_table_kats():
    z, _ = load_golden("table_kats.npz")
    # reference tests/test_diff.py:247-249: exact 3-point Chebyshev matrix
    expect3 = -np.array([[1.5, -2, .5], [.5, 0, -.5], [-.5, 2, -1.5]])
    np.testing.assert_array_almost_equal(z["cheb_n3_o1"], expect3, decimal=14)
    for key in z.files:
        if key.startswith("cheb_n") and "_o" in key:
            n = int(key.split("_")[1][1:])
            order = int(key.split("_o")[1])
            coords = z[f"cheb_n{n}_coords"]
            M = pencil.chebyshev_matrix(n, float(coords[0]), float(coords[-1]), order)
            assert np.array_equal(M, z[key]), key
    # SURVEY A.4 KAT: n=8 on [0, 2pi): W = i*[0,1,2,3,0,-3,-2,-1]
    np.testing.assert_allclose(z["fftW_n8_o1"].imag, [0, 1, 2, 3, 0, -3, -2, -1], atol=1e-14)
    np.testing.assert_allclose(z["fftW_n8_o2"].real, -np.array([0, 1, 4, 9, 16, 9, 4, 1.0]), atol=1e-13)
    np.testing.assert_allclose(z["fftW_n9_o1"].imag, [0, 1, 2, 3, 4, -4, -3, -2, -1], atol=1e-14)


def test_fd_weights_rational():
    # SURVEY A.2 table (sanity only; the bits come from numpy.linalg.solve)
    c = pencil.coefficients(1, 4)
    np.testing.assert_allclose(c["center"]["coefficients"], [1 / 12, -2 / 3, 0, 2 / 3, -1 / 12], atol=1e-14)
    np.testing.assert_allclose(c["forward"]["coefficients"], [-25 / 12, 4, -3, 4 / 3, -1 / 4], atol=1e-13)
    c = pencil.coefficients(2, 4)
    np.testing.assert_allclose(c["center"]["coefficients"], [-1 / 12, 4 / 3, -5 / 2, 4 / 3, -1 / 12], atol=1e-13)
    np.testing.assert_allclose(c["forward"]["coefficients"],
                               [15 / 4, -77 / 6, 107 / 6, -13, 61 / 12, -5 / 6], atol=1e-12)
    assert c["backward"]["offsets"] == [0, -1, -2, -3, -4, -5]
    with pytest.raises(ValueError):
        pencil.coefficients(1, 3)


def test_reference_analytic_checks():
    """Mirrors reference tests/test_diff.py:16-39 on the oracle (x^4 derivatives, 2-D Laplacian)."""
    ax = pencil.make_axis("equidistant", 100, 0, 1)
    x = ax["coords"]
    f = x ** 4
    np.testing.assert_array_almost_equal(pencil.make_diff([ax], 1, 0)(f), 4 * x ** 3, decimal=7)
    np.testing.assert_array_almost_equal(pencil.make_diff([ax], 2, 0)(f), 12 * x ** 2, decimal=7)
    X, Y = np.meshgrid(x, x, indexing="ij")
    lap = pencil.cartesian_laplacian(X ** 3 + Y ** 3, [x[1] - x[0]] * 2)
    np.testing.assert_array_almost_equal(6 * (X + Y), lap)
