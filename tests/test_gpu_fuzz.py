"""GPU parity fuzz: random grids (axis kinds, sizes, orders, accuracies, 1-4 dimensions) against
the pencil oracle.  FD / Log / Chebyshev must be bit-exact, periodic axes within 1e-11.
Seeds are fixed, so failures reproduce."""
import numpy as np
import pytest

import numgrids_b200 as ng
from conftest import TOL_FFT, rel_linf
from oracle import pencil
from tests_support import build_axes

pytestmark = pytest.mark.gpu


def _random_spec(rng, kind, big):
    if kind == "equidistant":
        return (kind, int(rng.integers(12, 70 if big else 30)), 0.0, float(rng.uniform(0.5, 3.0)))
    if kind == "periodic":
        n = int(rng.choice([8, 16, 32, 64, 128, 256, 12, 21, 30, 50]) if big else rng.choice([8, 9, 16, 20, 64]))
        return (kind, n, 0.0, float(rng.uniform(1.0, 7.0)))
    if kind == "chebyshev":
        return (kind, int(rng.integers(4, 140 if big else 24)), float(rng.uniform(-1, 0.5)), float(rng.uniform(1.0, 3.0)))
    return ("log", int(rng.integers(14, 40)), float(rng.uniform(0.05, 0.5)), float(rng.uniform(2.0, 20.0)))


@pytest.mark.parametrize("seed", range(12))
def test_random_grids_match_oracle(gpu_ready, seed):
    rng = np.random.default_rng(1000 + seed)
    nd = int(rng.integers(1, 5))
    kinds = rng.choice(["equidistant", "periodic", "chebyshev", "log"], size=nd)
    specs = [_random_spec(rng, k, big=(nd <= 2)) for k in kinds]
    axes = build_axes(ng, specs)
    g = ng.Grid(*axes)
    paxes = [pencil.make_axis(*s) for s in specs]
    f = rng.standard_normal(g.shape)
    for axis in range(nd):
        kind = specs[axis][0]
        for order in (1, 2):
            acc = int(rng.choice([2, 4, 6])) if kind in ("equidistant", "log") else 4
            if kind in ("equidistant", "log") and specs[axis][1] < 2 * (order + acc):
                acc = 2
            out = ng.Diff(g, order, axis, acc=acc)(f)
            ref = pencil.make_diff(paxes, order, axis, acc)(f)
            if kind == "periodic":
                assert rel_linf(out, ref) <= TOL_FFT, (specs, axis, order, rel_linf(out, ref))
            else:
                assert np.array_equal(out, ref), (specs, axis, order, acc, rel_linf(out, ref))


@pytest.mark.parametrize("seed", range(4))
def test_random_curvilinear_match_oracle(gpu_ready, seed):
    """Custom metric on a random 2-D / 3-D grid: all four operators against the oracle transcription."""
    rng = np.random.default_rng(2000 + seed)
    nd = 2 + seed % 2
    kinds = rng.choice(["equidistant", "periodic", "chebyshev"], size=nd)
    specs = [_random_spec(rng, k, big=False) for k in kinds]
    specs = [(k, max(n, 14) if k == "equidistant" else n, lo, hi) for k, n, lo, hi in specs]
    axes = build_axes(ng, specs)
    fns = [lambda c, i=i: 1.0 + 0.3 * i + 0.5 * np.cos(c[i % len(c)]) ** 2 + 0.1 * c[0] for i in range(nd)]
    g = ng.CurvilinearGrid(*axes, scale_factors=fns)
    paxes = [pencil.make_axis(*s) for s in specs]
    D = [pencil.make_diff(paxes, 1, i, 4) for i in range(nd)]
    mesh = np.meshgrid(*[a.coords for a in axes], indexing="ij")
    h = tuple(fn(mesh) for fn in fns)
    f = rng.standard_normal(g.shape)
    v = [rng.standard_normal(g.shape) for _ in range(nd)]
    has_fft = any(k == "periodic" for k in kinds)
    tol = 1e-10 if has_fft else 0.0          # white-noise fields: spectral noise is relative to max|ref|

    def check(a, b, what):
        if tol == 0.0:
            assert np.array_equal(a, b), (specs, what, rel_linf(a, b))
        else:
            assert rel_linf(a, b) <= tol, (specs, what, rel_linf(a, b))

    check(g.laplacian(f), pencil.laplacian(f, h, D), "lap")
    check(g.divergence(*v), pencil.divergence(v, h, D), "div")
    for i, (a, b) in enumerate(zip(g.gradient(f), pencil.gradient(f, h, D))):
        check(a, b, ("grad", i))
    c, cr = g.curl(*v), pencil.curl(v, h, D)
    c, cr = ((c,), (cr,)) if nd == 2 else (c, cr)
    for i, (a, b) in enumerate(zip(c, cr)):
        check(a, b, ("curl", i))


@pytest.mark.parametrize("nphi,nz", [(16, 11), (64, 7), (256, 5), (512, 3), (12, 9)])
def test_singular_rows_do_not_leak_through_paired_fft(gpu_ready, nphi, nz):
    """Cylindrical grid including r = 0 with an ODD z extent: the FFT kernels transform two real
    pencils as one complex one, and the rows at r = 0 are inf/nan before the final isfinite select
    (reference numgrids/curvilinear.py:236-244).  A singular pencil must not poison its partner --
    every kernel family (radix-2, two-pass, ring, dense fallback), phi as a middle and as the last axis."""
    for order_axes in ("r_phi_z", "z_r_phi"):
        r = ("chebyshev", 9, 0.0, 1.5)
        phi = ("periodic", nphi, 0.0, 2 * np.pi)
        z = ("equidistant", nz + 8, -1.0, 1.0)
        specs = [r, phi, z] if order_axes == "r_phi_z" else [z, r, phi]
        ir, iphi = (0, 1) if order_axes == "r_phi_z" else (1, 2)
        ones = lambda c: np.ones_like(c[0])
        fns = [ones, ones, ones]
        fns[iphi] = lambda c: c[ir]                       # h_phi = r, zero on the axis
        axes = build_axes(ng, specs)
        g = ng.CurvilinearGrid(*axes, scale_factors=fns)
        paxes = [pencil.make_axis(*s) for s in specs]
        D = [pencil.make_diff(paxes, 1, i, 4) for i in range(3)]
        mesh = np.meshgrid(*[a.coords for a in axes], indexing="ij")
        h = tuple(fn(mesh) for fn in fns)
        rng = np.random.default_rng(nphi + nz)
        f = rng.standard_normal(g.shape)
        v = [rng.standard_normal(g.shape) for _ in range(3)]
        with np.errstate(all="ignore"):
            refs = {"lap": pencil.laplacian(f, h, D), "div": pencil.divergence(v, h, D)}
            for i, x in enumerate(pencil.gradient(f, h, D)):
                refs["grad%d" % i] = x
            for i, x in enumerate(pencil.curl(v, h, D)):
                refs["curl%d" % i] = x
        got = {"lap": g.laplacian(f), "div": g.divergence(*v)}
        for i, x in enumerate(g.gradient(f)):
            got["grad%d" % i] = x
        for i, x in enumerate(g.curl(*v)):
            got["curl%d" % i] = x
        for k in refs:
            assert np.all(np.isfinite(got[k])), (order_axes, k)
            assert np.array_equal(got[k] == 0.0, refs[k] == 0.0) or rel_linf(got[k], refs[k]) <= 1e-10, (order_axes, k)
            assert rel_linf(got[k], refs[k]) <= 1e-10, (order_axes, k, rel_linf(got[k], refs[k]))
