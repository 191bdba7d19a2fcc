"""CPU: host logic of the drop-in package -- tables, dispatch, validation, lazy caches, as_matrix.

Mirrors the reference's tests/test_api.py:18-102, :1026-1090, tests/test_diff.py:242-283, :327-366
and tests/test_curvilinear.py:17-66 (construction/validation parts; the numerical parts run on the
GPU, see test_gpu_*.py).  No compute call is made here.
"""
import numpy as np
import pytest

import numgrids_b200 as ng
from numgrids_b200 import _tables
from numgrids_b200.axes import ChebyshevAxis, EquidistantAxis, LogAxis
from numgrids_b200.diff import ChebyshevDiff, FFTDiff, FiniteDifferenceDiff, LogDiff
from conftest import load_golden
from oracle import pencil


# ------------------------------------------------------------------ tables == oracle == reference
def test_axis_coordinates_match_oracle():
    for n, lo, hi in ((4, 0.0, 1.0), (33, -1.0, 2.5), (128, 0.3, np.pi - 0.3)):
        a = ng.create_axis(ng.AxisType.CHEBYSHEV, n, lo, hi)
        ci, c = pencil.chebyshev_coords(n, lo, hi)
        assert np.array_equal(a.coords, c) and np.array_equal(a.coords_internal, ci)
        a = ng.create_axis(ng.AxisType.EQUIDISTANT, n, lo, hi)
        assert np.array_equal(a.coords, pencil.equidistant_coords(n, lo, hi)[1])
        a = ng.create_axis(ng.AxisType.EQUIDISTANT_PERIODIC, n, lo, hi)
        assert np.array_equal(a.coords, pencil.equidistant_coords(n, lo, hi, True)[1])
        a = ng.create_axis(ng.AxisType.LOGARITHMIC, n, 0.1 + lo * 0 + 0.1, 5.0)
        assert np.array_equal(a.coords, pencil.log_coords(n, 0.2, 5.0)[1])
    # SURVEY A.1 KAT
    assert list(ChebyshevAxis(4, 0, 1).coords) == [0.0, 0.2500000000000001, 0.75, 1.0]


def test_fd_schemes_match_oracle_bits():
    for order in (1, 2, 3, 4):
        for acc in (2, 4, 6):
            s = _tables.fd_schemes(order, acc)
            c = pencil.coefficients(order, acc)
            for key in ("center", "forward", "backward"):
                assert list(s[key][0]) == list(c[key]["offsets"])
                assert np.array_equal(s[key][1], c[key]["coefficients"])
    with pytest.raises(ValueError):
        _tables.fd_schemes(1, 3)
    w = _tables.snap_unit_weights(np.array([1.0000000000000002, 0.5, 1 - 2e-14]))
    assert w[0] == 1.0 and w[1] == 0.5 and w[2] != 1.0
    assert _tables.fd_min_points(2, 4) == 7 and _tables.fd_min_points(1, 4) == 6


def test_chebyshev_tables_match_reference_kats():
    z, _ = load_golden("table_kats.npz")
    for key in z.files:
        if key.startswith("cheb_n") and "_o" in key:
            n = int(key.split("_")[1][1:])
            order = int(key.split("_o")[1])
            coords = z[f"cheb_n{n}_coords"]
            ax = ChebyshevAxis(n, float(coords[0]), float(coords[-1]))
            M = _tables.chebyshev_matrix(ax.coords_internal, ax.coords, order, False)
            assert np.array_equal(M, z[key]), key
    assert _tables.chebyshev_descending(2, 0, 3) and not _tables.chebyshev_descending(2, 2, 3)
    assert not _tables.chebyshev_descending(1, 0, 3) and not _tables.chebyshev_descending(2, 0, 1)
    assert not _tables.chebyshev_descending(3, 0, 2) and _tables.chebyshev_descending(4, 1, 3)


def test_fft_tables_match_reference_kats():
    z, _ = load_golden("table_kats.npz")
    for n, lo, hi in ((8, 0.0, 2 * np.pi), (9, 0.0, 2 * np.pi), (16, 0.0, 1.0)):
        ax = EquidistantAxis(n, lo, hi, periodic=True)
        for order in (1, 2, 3):
            assert np.array_equal(_tables.fft_multiplier(n, ax.spacing, order), z[f"fftW_n{n}_o{order}"])
    W = _tables.fft_multiplier(12, 2 * np.pi / 12, 1)
    D = _tables.fft_dense_matrix(W)
    f = np.exp(np.sin(np.arange(12) * 2 * np.pi / 12))
    np.testing.assert_allclose(D @ f, pencil.fft_apply(f, 0, W), rtol=0, atol=1e-13)


def test_compact_coefficient_roundtrip():
    r = np.linspace(0.5, 2, 5)
    th = np.linspace(0.3, 2.8, 4)
    R, T, P = np.meshgrid(r, th, np.arange(3.0), indexing="ij")
    for full, nvals in ((R * np.sin(T), 20), (np.ones_like(R), 1), (R, 5), (R + P, 15), (R * T * (P + 1), 60)):
        vals, strides = _tables.compact_coefficient(full, full.shape)
        assert len(vals) == nvals
        idx = sum(np.indices(full.shape)[a] * strides[a] for a in range(3))
        assert np.array_equal(vals[idx], full)
    nanfull = np.full((3, 4), np.nan)
    vals, strides = _tables.compact_coefficient(nanfull, (3, 4))
    assert len(vals) == 1 and strides == [0, 0]


# ------------------------------------------------------------------ dispatch / validation
def test_dispatch_types_and_acc():
    eq = ng.create_axis(ng.AxisType.EQUIDISTANT, 20, 0, 1)
    per = ng.create_axis(ng.AxisType.EQUIDISTANT_PERIODIC, 20, 0, 1)
    ch = ng.create_axis(ng.AxisType.CHEBYSHEV, 20, 0, 1)
    lg = ng.create_axis(ng.AxisType.LOGARITHMIC, 20, 0.1, 1)
    g = ng.Grid(eq, per, ch, lg)
    assert type(ng.Diff(g, 1, 0).operator) == FiniteDifferenceDiff
    assert type(ng.Diff(g, 1, 1).operator) == FFTDiff
    assert type(ng.Diff(g, 1, 2).operator) == ChebyshevDiff
    assert type(ng.Diff(g, 1, 3).operator) == LogDiff
    assert ng.Diff(g, 1, 0).operator.acc == 4
    assert ng.Diff(g, 1, 0, acc=6).operator.acc == 6
    assert ng.Diff(g, 1, 1, acc=2).operator.acc == 2
    assert ng.Diff(g, 2, 3, acc=2).operator.acc == 2
    assert FFTDiff(ng.Grid(per), 1, 0).acc == 6 and LogDiff(ng.Grid(lg), 1, 0).acc == 6
    d = ng.Diff(g, 2, 2).operator
    assert (d.order, d.axis_index, d.grid, d.axis) == (2, 2, g, ch)


def test_validation_errors():
    g = ng.Grid(EquidistantAxis(20, 0, 1))
    with pytest.raises(ValueError):
        ng.Diff(g, 0, 0)
    with pytest.raises(ValueError):
        ng.Diff(g, -1, 0)
    with pytest.raises(TypeError):
        ng.Diff("grid", 1, 0)
    with pytest.raises(ValueError):
        ng.Diff(g, 1, -1)
    with pytest.raises(ValueError):
        ng.Diff(g, 1, 1)
    with pytest.raises(TypeError):
        FiniteDifferenceDiff(ng.Grid(ChebyshevAxis(10, 0, 1)), 1, 0)
    with pytest.raises(TypeError):
        FFTDiff(ng.Grid(ChebyshevAxis(10, 0, 1)), 1, 0)
    with pytest.raises(TypeError):
        FFTDiff(g, 1, 0)                       # equidistant but not periodic
    with pytest.raises(TypeError):
        LogDiff(g, 1, 0)
    with pytest.raises(NotImplementedError):
        ng.create_axis("nonsense", 10, 0, 1)
    with pytest.raises(ValueError):
        EquidistantAxis(0, 0, 1)
    with pytest.raises(ValueError):
        EquidistantAxis(10, 1, 1)
    with pytest.raises(ValueError):
        LogAxis(10, 0.0, 1.0)
    with pytest.raises(ValueError):
        ng.Diff(ng.Grid(EquidistantAxis(5, 0, 1)), 2, 0)      # too short for the 6-tap one-sided scheme
    assert ng.Axis is ng.create_axis


def test_grid_surface_is_lazy_and_compatible():
    ax = EquidistantAxis(6, 0, 1)
    per = EquidistantAxis(4, 0, 1, periodic=True)
    g = ng.Grid(ax, per)
    assert g._meshed_coords is None and g.shape == (6, 4) and g.size == 24 and g.ndims == 2
    X, Y = g.meshed_coords
    assert X.shape == (6, 4) and g.meshed_coords[0] is X
    b = g.boundary
    assert b[0].all() and b[-1].all() and not b[1:-1].any()
    assert set(g.cache) == {"diffs", "integrals", "interpolators", "faces"}
    assert np.array_equal(g[2, 5], [ax[2], per[1]])
    assert g.refine().shape == (12, 8) and g.coarsen().shape == (3, 2)
    assert g.refine_axis(0, 1.5).shape == (9, 4)
    with pytest.raises(ValueError):
        g.refine_axis(2)
    assert ng.Grid(ax).coords is ax.coords and repr(g) == "Grid(shape=(6, 4), ndims=2)"
    big = ng.Grid(EquidistantAxis(1024), EquidistantAxis(1024), EquidistantAxis(1024))   # no 26 GB meshgrid
    assert big.size == 1024 ** 3


def test_diff_helper_cache_keys():
    g = ng.Grid(EquidistantAxis(20, 0, 1))
    assert g.cache["diffs"] == {}
    for key in ((1, 0, 4), (2, 0, 4), (1, 0, 6)):
        try:
            ng.diff(g, np.zeros(20), *key)
        except ng.NumgridsB200Error:
            pass                                 # no GPU here: the operator is cached before the apply
        assert key in g.cache["diffs"]
    assert len(g.cache["diffs"]) == 3


def test_curvilinear_construction_and_lazy_caches():
    r = ng.create_axis(ng.AxisType.CHEBYSHEV, 8, 0.5, 2)
    th = ng.create_axis(ng.AxisType.CHEBYSHEV, 7, 0.3, 2.8)
    ph = ng.create_axis(ng.AxisType.EQUIDISTANT_PERIODIC, 8, 0, 2 * np.pi)
    with pytest.raises(ValueError):
        ng.CurvilinearGrid(r, th, scale_factors=(lambda c: c[0],))
    g = ng.SphericalGrid(r, th, ph)
    assert isinstance(g, ng.Grid) and g._diff_ops is None and g._h_arrays is None and g._jacobian is None
    h = g._evaluate_scale_factors()
    assert g._evaluate_scale_factors() is h and len(h) == 3
    R, T, _ = g.meshed_coords
    assert np.array_equal(h[2], R * np.sin(T))
    assert np.array_equal(g._evaluate_jacobian(), np.ones(g.shape) * h[0] * h[1] * h[2])
    g._ensure_diff_ops()
    ops = g._diff_ops
    g._ensure_diff_ops()
    assert g._diff_ops is ops and [type(ops[i].operator) for i in range(3)] == [ChebyshevDiff, ChebyshevDiff, FFTDiff]
    with pytest.raises(ValueError):
        g.divergence(np.zeros(g.shape))
    g4 = ng.CurvilinearGrid(r, r, r, r, scale_factors=(lambda c: c[0],) * 4)
    with pytest.raises(ValueError):
        g4.curl(*[np.zeros(g4.shape)] * 4)


# ------------------------------------------------------------------ as_matrix (host, lazy)
def test_chebyshev_matrix_golden_and_kron_layout():
    # reference tests/test_diff.py:242-283
    D = ChebyshevDiff(ng.Grid(ChebyshevAxis(3, -1, 1)), 1, 0).as_matrix().toarray()
    np.testing.assert_array_almost_equal(D, -np.array([[1.5, -2, .5], [.5, 0, -.5], [-.5, 2, -1.5]]))
    ax = ChebyshevAxis(3, -1, 1)
    g = ng.Grid(ax, ax)
    D1 = ChebyshevDiff(ng.Grid(ax), 1, 0).as_matrix().toarray()
    np.testing.assert_array_almost_equal(ChebyshevDiff(g, 1, 0).as_matrix().toarray(), np.kron(D1, np.eye(3)))
    np.testing.assert_array_almost_equal(ChebyshevDiff(g, 1, 1).as_matrix().toarray(), np.kron(np.eye(3), D1))


@pytest.mark.parametrize("specs,axis,order", [
    ([("equidistant", 14, 0.0, 1.0), ("periodic", 9, 0.0, 3.0)], 0, 1),
    ([("equidistant", 14, 0.0, 1.0), ("periodic", 9, 0.0, 3.0)], 1, 2),
    ([("log", 15, 0.1, 10.0), ("chebyshev", 8, 0.0, 1.0)], 0, 2),
    ([("log", 15, 0.1, 10.0), ("chebyshev", 8, 0.0, 1.0)], 1, 2),
    ([("periodic", 16, 0.0, 2 * np.pi)], 0, 1),
])
def test_as_matrix_equals_oracle_operator(specs, axis, order):
    """matrix == operator (reference tests/test_diff.py:391-420, :544-582), operator = the oracle."""
    from tests_support import build_axes, smooth_field
    axes = build_axes(ng, specs)
    g = ng.Grid(*axes)
    f = smooth_field([a.coords for a in axes])
    M = ng.Diff(g, order, axis).as_matrix()
    assert M.shape == (g.size, g.size)
    ref = pencil.make_diff([pencil.make_axis(*s) for s in specs], order, axis, 4)(f)
    np.testing.assert_allclose((M @ f.reshape(-1)).reshape(g.shape), ref, rtol=0,
                               atol=1e-10 * np.max(np.abs(ref)))


def test_mixed_radix_lengths():
    """Which periodic-axis lengths the library's native mixed-radix transform serves (mirrors ng_fft_mixed_supported,
    csrc/fft_mixed.cu): not a power of two, prime factors up to 13, at most 8192 points."""
    from numgrids_b200 import _tables
    yes = [3, 6, 9, 12, 45, 96, 100, 360, 1000, 1536, 2310, 3000, 6000, 2 * 3 * 5 * 7 * 11 * 3, 13 ** 3]
    no = [1, 2, 4, 64, 1024, 8192, 17, 34, 118, 2 * 59, 19 * 5, 8193, 9000, 10 ** 4]
    assert all(_tables.fft_mixed_supported(n) for n in yes)
    assert not any(_tables.fft_mixed_supported(n) for n in no)


def test_periodic_axis_routing():
    """FFTDiff.route: which transform serves a periodic axis (numgrids/diff.py:133-143 is one numpy.fft call for all of them)."""
    from numgrids_b200.diff import FFTDiff
    r = FFTDiff.route
    assert FFTDiff.NONPOW2 == "auto"
    assert [r(n, 1) for n in (2, 64, 1024, 8192)] == ["pow2"] * 4
    # plain first / second derivatives: zero-padded where measured faster (even n, 65..512, not just above 128)
    assert [r(n, 1) for n in (96, 100, 200, 360, 500)] == ["padded"] * 5
    assert [r(n, 2) for n in (96, 500)] == ["padded"] * 2
    assert [r(n, 1) for n in (6, 45, 60, 65, 81, 150, 180, 1000, 1536, 3000, 6000)] == ["mixed"] * 11
    # higher orders and the composites: native wherever the length factors into 2 3 5 7 11 13
    assert [r(n, 3) for n in (12, 96, 360, 1000)] == ["mixed"] * 4
    assert [r(n, 1, prefer_native=True) for n in (24, 100, 360, 500)] == ["mixed"] * 4
    # a prime factor above 13: padded for orders 1-2 from 65 points on, else the dense contraction
    assert [r(n, 1) for n in (118, 177, 1018)] == ["padded"] * 3
    assert [r(n, 1, prefer_native=True) for n in (118,)] == ["padded"]
    assert [r(n, 3) for n in (118, 34)] == ["dense"] * 2
    assert r(34, 1) == "dense"


def test_mixed_radix_factorisation():
    """ng_fft_mixed_radices (host only): the passes of the native transform -- product n, allowed radices, powers of two
    first, as few passes as any factorisation into radices <= 16 allows -- and the lengths it refuses."""
    import ctypes as C
    from numgrids_b200 import _lib, _tables
    lib = _lib.load()
    allowed = set(range(2, 17))

    def radices(n):
        buf, cnt = (C.c_int32 * 12)(), C.c_int32(0)
        rc = lib.ng_fft_mixed_radices(n, buf, C.byref(cnt))
        return rc, list(buf[:cnt.value])

    def fewest(n, memo={1: 0}):
        if n not in memo:
            memo[n] = 1 + min(fewest(n // r) for r in range(2, 17) if n % r == 0)
        return memo[n]

    for n in (6, 12, 24, 45, 48, 60, 96, 100, 182, 200, 360, 500, 1000, 1536, 2187, 2310, 3000, 6000, 8190):
        rc, rs = radices(n)
        assert rc == 0 and int(np.prod(rs)) == n and set(rs) <= allowed, (n, rs)
        assert len(rs) == fewest(n), (n, rs)
        pow2 = [r for r in rs if r & (r - 1) == 0]
        assert rs[:len(pow2)] == pow2, (n, rs)
    assert radices(1000)[1] == [10, 10, 10] and radices(1536)[1] == [16, 8, 12] and radices(360)[1] == [8, 5, 9]
    for n in (1, 2, 64, 1024, 17, 34, 118, 8193, 9000):
        assert not _tables.fft_mixed_supported(n)
        assert radices(n)[0] != 0
    # the host layer's rule and the library's agree on every length
    for n in range(1, 8300):
        rc, rs = radices(n)
        assert (rc == 0) == _tables.fft_mixed_supported(n), n
        if rc == 0:
            assert int(np.prod(rs)) == n and len(rs) == fewest(n), (n, rs)
