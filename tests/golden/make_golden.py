"""Generate tests/golden/*.npz by running the REAL reference (build container only).

Run:  python tests/golden/make_golden.py      (needs /root/reference; findiff is
supplied by oracle/findiff_port.py because the real package is not installable
offline -- see oracle/__init__.py).

For every case the script
  1. builds the grid with the reference's own classes (numgrids.api / .diff /
     .curvilinear imported from /root/reference),
  2. evaluates the reference operator on a deterministic smooth field,
  3. asserts the pencil oracle (oracle/pencil.py) reproduces it with
     ``np.array_equal`` (bit-exact; FFT included since both call numpy.fft),
  4. stores inputs + outputs in an .npz next to this file.

The fixtures are what the GPU parity tests and the CPU oracle tests compare to;
/root/reference itself never travels to the GPU box.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import pencil, reference_loader  # noqa: E402

ref = reference_loader.load()
from numgrids import (AxisType, CurvilinearGrid, CylindricalGrid, Diff, Grid,  # noqa: E402
                      PolarGrid, SphericalGrid, create_axis)

KIND = {"equidistant": AxisType.EQUIDISTANT, "periodic": AxisType.EQUIDISTANT_PERIODIC,
        "chebyshev": AxisType.CHEBYSHEV, "log": AxisType.LOGARITHMIC}


def smooth_field(coords, shift=0.0):
    """Deterministic smooth field from 1-D coordinate vectors (no meshgrid needed)."""
    nd = len(coords)
    f = np.ones([len(c) for c in coords])
    for i, c in enumerate(coords):
        shp = [1] * nd
        shp[i] = len(c)
        c = np.asarray(c).reshape(shp)
        f = f * np.sin((i + 1.3) * c + 0.2 * i + shift) + 0.1 * np.cos(c + shift)
    return f


def ref_axes(specs):
    return [create_axis(KIND[k], n, lo, hi) for k, n, lo, hi in specs]


# ------------------------------------------------------------------ Diff cases
DIFF_GRIDS = [
    [("equidistant", 24, 0.0, 1.0)],
    [("periodic", 16, 0.0, 2 * np.pi)],
    [("periodic", 21, 0.0, 3.0)],
    [("chebyshev", 17, -1.0, 2.0)],
    [("log", 25, 0.1, 10.0)],
    [("equidistant", 20, 0.0, 2.0), ("equidistant", 33, -1.0, 1.0)],
    [("chebyshev", 12, 0.0, 1.0), ("chebyshev", 9, -1.0, 2.0)],
    [("periodic", 32, 0.0, 2 * np.pi), ("periodic", 15, 0.0, 1.0)],
    [("log", 22, 0.5, 4.0), ("equidistant", 21, 0.0, 1.0)],
    [("chebyshev", 9, 0.1, 2.0), ("periodic", 10, 0.0, 2 * np.pi), ("equidistant", 16, 0.0, 1.0)],
    [("equidistant", 14, 0.0, 1.0), ("log", 15, 0.1, 10.0), ("periodic", 32, 0.0, 3.0)],
    [("chebyshev", 8, 0.0, 1.0), ("chebyshev", 11, 0.5, 1.0), ("chebyshev", 10, -1.0, 1.0)],
    [("periodic", 6, 0.0, 1.0), ("chebyshev", 17, 0.0, 1.0), ("periodic", 64, 0.0, 2 * np.pi)],
    [("chebyshev", 5, 0.0, 1.0), ("equidistant", 16, 0.0, 1.0), ("chebyshev", 6, 0.0, 1.0),
     ("periodic", 8, 0.0, 1.0)],
]


def gen_diff():
    arrays, meta = {}, []
    cid = 0
    for specs in DIFF_GRIDS:
        axes = ref_axes(specs)
        g = Grid(*axes)
        paxes = [pencil.make_axis(*s) for s in specs]
        for pa, ra in zip(paxes, axes):
            assert np.array_equal(pa["coords"], ra.coords)
        f = smooth_field([a.coords for a in axes])
        fkey = f"f{cid}"
        arrays[fkey] = f
        for axis in range(g.ndims):
            kind = specs[axis][0]
            accs = (2, 4, 6) if kind in ("equidistant", "log") and g.ndims <= 2 else (4,)
            orders = (1, 2, 3) if kind != "chebyshev" else (1, 2, 3, 4)
            for order in orders:
                for acc in accs:
                    if kind in ("equidistant", "log") and order == 3 and acc == 6:
                        continue
                    if kind in ("equidistant", "log") and len(axes[axis]) < 2 * (order + acc):
                        continue
                    r = Diff(g, order, axis, acc=acc)(f)
                    o = pencil.make_diff(paxes, order, axis, acc)(f)
                    assert np.array_equal(r, o), (specs, axis, order, acc)
                    key = f"d{cid}"
                    arrays[key] = r
                    meta.append({"id": key, "f": fkey, "axes": specs, "axis": axis,
                                 "order": order, "acc": acc, "kind": kind})
                    cid += 1
        cid += 1
    np.savez_compressed(os.path.join(HERE, "diff_cases.npz"),
                        __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    return len(meta)


# ------------------------------------------------------------ curvilinear cases
def _ones(c):
    return np.ones_like(c[0])


CURV = [
    ("spherical", [("chebyshev", 12, 0.5, 2.0), ("chebyshev", 10, 0.3, np.pi - 0.3),
                   ("periodic", 16, 0.0, 2 * np.pi)]),
    ("spherical", [("chebyshev", 9, 0.0, 1.0), ("chebyshev", 8, 0.0, np.pi),       # singular r=0, theta=0,pi
                   ("periodic", 9, 0.0, 2 * np.pi)]),
    ("spherical", [("equidistant", 14, 0.5, 2.0), ("equidistant", 13, 0.3, np.pi - 0.3),
                   ("periodic", 12, 0.0, 2 * np.pi)]),
    ("cylindrical", [("chebyshev", 11, 0.1, 2.0), ("periodic", 12, 0.0, 2 * np.pi),
                     ("chebyshev", 9, -1.0, 1.0)]),
    ("cylindrical", [("log", 15, 0.1, 2.0), ("periodic", 8, 0.0, 2 * np.pi),
                     ("equidistant", 14, -1.0, 1.0)]),
    ("polar", [("chebyshev", 13, 0.1, 1.0), ("periodic", 20, 0.0, 2 * np.pi)]),
    ("polar", [("chebyshev", 10, 0.0, 1.0), ("periodic", 14, 0.0, 2 * np.pi)]),          # singular r=0
    ("unit3", [("equidistant", 14, 0.0, 1.0), ("equidistant", 13, 0.0, 2.0),
               ("equidistant", 16, -1.0, 1.0)]),
    ("unit2", [("chebyshev", 11, 0.0, 1.0), ("chebyshev", 12, 0.0, 2.0)]),
    ("custom3", [("chebyshev", 9, 0.5, 1.5), ("equidistant", 15, 0.2, 1.2), ("periodic", 10, 0.0, 1.0)]),
]

CUSTOM3 = (lambda c: 1.0 + c[0] ** 2, lambda c: c[0] * np.exp(c[1]), lambda c: 2.0 + np.sin(2 * np.pi * c[2]) * c[1])


def build_curv(name, axes):
    if name == "spherical":
        return SphericalGrid(*axes)
    if name == "cylindrical":
        return CylindricalGrid(*axes)
    if name == "polar":
        return PolarGrid(*axes)
    if name == "unit3":
        return CurvilinearGrid(*axes, scale_factors=(_ones, _ones, _ones))
    if name == "unit2":
        return CurvilinearGrid(*axes, scale_factors=(_ones, _ones))
    if name == "custom3":
        return CurvilinearGrid(*axes, scale_factors=CUSTOM3)
    raise KeyError(name)


def gen_curv():
    arrays, meta = {}, []
    for cid, (name, specs) in enumerate(CURV):
        axes = ref_axes(specs)
        g = build_curv(name, axes)
        nd = g.ndims
        paxes = [pencil.make_axis(*s) for s in specs]
        D = [pencil.make_diff(paxes, 1, i, 4) for i in range(nd)]
        h = g._evaluate_scale_factors()
        f = smooth_field([a.coords for a in axes])
        v = [smooth_field([a.coords for a in axes], shift=0.37 * (i + 1)) for i in range(nd)]
        arrays[f"c{cid}_f"] = f
        for i in range(nd):
            arrays[f"c{cid}_v{i}"] = v[i]
        lap = g.laplacian(f)
        assert np.array_equal(lap, pencil.laplacian(f, h, D)), (name, "lap")
        arrays[f"c{cid}_lap"] = lap
        grad = g.gradient(f)
        og = pencil.gradient(f, h, D)
        div = g.divergence(*v)
        assert np.array_equal(div, pencil.divergence(v, h, D)), (name, "div")
        arrays[f"c{cid}_div"] = div
        for i in range(nd):
            assert np.array_equal(grad[i], og[i]), (name, "grad", i)
            arrays[f"c{cid}_grad{i}"] = grad[i]
        cu = g.curl(*v)
        oc = pencil.curl(v, h, D)
        if nd == 2:
            assert np.array_equal(cu, oc), (name, "curl")
            arrays[f"c{cid}_curl0"] = cu
        else:
            for i in range(3):
                assert np.array_equal(cu[i], oc[i]), (name, "curl", i)
                arrays[f"c{cid}_curl{i}"] = cu[i]
        meta.append({"id": f"c{cid}", "grid": name, "axes": specs, "ndim": nd,
                     "all_finite_in": bool(np.all(np.isfinite(lap)))})
    np.savez_compressed(os.path.join(HERE, "curv_cases.npz"),
                        __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    return len(meta)


# -------------------------------------------------------------- Cartesian Laplacian
def gen_cartlap():
    """Reference idiom d2_dx2(f)+d2_dy2(f)(+d2_dz2(f)) (reference tests/test_diff.py:35-37)."""
    arrays, meta = {}, []
    cases = [
        [("equidistant", 20, 0.0, 1.0), ("equidistant", 33, -1.0, 1.0)],
        [("equidistant", 18, 0.0, 1.0), ("equidistant", 17, 0.0, 2.0), ("equidistant", 40, 0.0, 1.0)],
        [("equidistant", 12, 0.0, 1.0), ("equidistant", 70, 0.0, 2.0), ("equidistant", 13, 0.0, 3.0)],
    ]
    for cid, specs in enumerate(cases):
        axes = ref_axes(specs)
        g = Grid(*axes)
        f = smooth_field([a.coords for a in axes])
        for acc in (2, 4, 6):
            r = None
            for a in range(g.ndims):
                d = Diff(g, 2, a, acc=acc)(f)
                r = d if r is None else r + d
            o = pencil.cartesian_laplacian(f, [a.spacing for a in axes], acc)
            assert np.array_equal(r, o)
            arrays[f"l{cid}_acc{acc}"] = r
        arrays[f"l{cid}_f"] = f
        meta.append({"id": f"l{cid}", "axes": specs})
    np.savez_compressed(os.path.join(HERE, "cartlap_cases.npz"),
                        __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    return len(meta)


# -------------------------------------------------------------- table KATs
def gen_tables():
    """1-D matrices / multipliers straight from reference objects (host-table KATs)."""
    arrays = {}
    for n, lo, hi in ((3, -1.0, 1.0), (4, 0.0, 1.0), (17, -1.0, 2.0), (32, 0.5, 2.0)):
        ax = create_axis(AxisType.CHEBYSHEV, n, lo, hi)
        for order in (1, 2, 3, 4):
            M = Diff(Grid(ax), order, 0).as_matrix().toarray()
            assert np.array_equal(M, pencil.chebyshev_matrix(n, lo, hi, order))
            arrays[f"cheb_n{n}_o{order}"] = M
        arrays[f"cheb_n{n}_coords"] = ax.coords
    for n, lo, hi in ((8, 0.0, 2 * np.pi), (9, 0.0, 2 * np.pi), (16, 0.0, 1.0)):
        ax = create_axis(AxisType.EQUIDISTANT_PERIODIC, n, lo, hi)
        for order in (1, 2, 3):
            W = Diff(Grid(ax), order, 0).operator._W_1d
            assert np.array_equal(W, pencil.fft_multiplier(n, ax.spacing, order))
            arrays[f"fftW_n{n}_o{order}"] = W
    np.savez_compressed(os.path.join(HERE, "table_kats.npz"), **arrays)
    return len(arrays)


if __name__ == "__main__":
    print("diff cases    :", gen_diff())
    print("curv cases    :", gen_curv())
    print("cartlap cases :", gen_cartlap())
    print("table KATs    :", gen_tables())
    print("findiff real? :", reference_loader.real_findiff_available())
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(f"  {fn}: {os.path.getsize(os.path.join(HERE, fn)) / 1024:.0f} KiB")
