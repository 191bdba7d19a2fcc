"""Single-pass fused curvilinear Laplacian (csrc/curvlap3d.cu) on all-finite-difference 3-D grids: bit for bit
against the oracle's literal transcription of numgrids/curvilinear.py:235-244 (oracle/pencil.py::laplacian) and
against the six-pass path it replaces, on unit-scale grids (the second reference idiom of the Cartesian
Laplacian), general metrics, logarithmic axes and singular metrics; shapes cover several tiles and x chunks,
ragged y tiles and the single-z-tile case."""
import zlib

import numpy as np
import pytest

import numgrids_b200 as ng
from conftest import rel_linf
from numgrids_b200 import _lib
from oracle import pencil

pytestmark = pytest.mark.gpu
A = ng.AxisType
ONE = lambda c: np.ones_like(c[0])

METRICS = {
    "unit": (ONE, ONE, ONE),
    "cylindrical_like": (ONE, lambda c: c[0], ONE),                       # J = r: varies along axis 0 only
    "custom": (lambda c: 1.0 + c[0] ** 2, lambda c: c[0] * np.exp(c[1]), lambda c: 2.0 + np.sin(2 * np.pi * c[2]) * c[1]),
    "singular": (ONE, lambda c: c[0], lambda c: c[0] * np.sin(c[1])),     # r = 0 and sin(theta) = 0 on faces
}


def _axes(kinds, shape, lows=(0.0, 0.0, 0.0)):
    out, spec = [], []
    for k, n, lo in zip(kinds, shape, lows):
        if k == "log":
            out.append(ng.create_axis(A.LOGARITHMIC, n, 0.1, 10.0))
            spec.append(pencil.make_axis("log", n, 0.1, 10.0))
        else:
            hi = lo + 1.0 + 0.25 * len(out)
            out.append(ng.create_axis(A.EQUIDISTANT, n, lo, hi))
            spec.append(pencil.make_axis("equidistant", n, lo, hi))
    return out, spec


def _oracle(spec, fns, f):
    mesh = np.meshgrid(*[s["coords"] for s in spec], indexing="ij")
    h = tuple(fn(mesh) for fn in fns)
    D = [pencil.make_diff(spec, 1, i) for i in range(3)]
    return pencil.laplacian(f, h, D)


def _six_pass(g, f, monkeypatch):
    monkeypatch.setenv("NG_CURV_FUSED", "0")
    _lib.tuning_reload()
    try:
        out = g.laplacian(f)
        assert not _lib.last_kernel().startswith("curvlap3d")
    finally:
        monkeypatch.delenv("NG_CURV_FUSED")
        _lib.tuning_reload()
    return out


CASES = [
    ("unit", ("eq", "eq", "eq"), (32, 32, 64)),
    ("unit", ("eq", "eq", "eq"), (130, 72, 192)),          # 2 x chunks, ragged y tile (72 = 4.5 tiles of 16), 3 z tiles
    ("unit", ("eq", "eq", "eq"), (16, 16, 64)),            # smallest supported
    ("cylindrical_like", ("eq", "eq", "eq"), (40, 36, 128)),
    ("custom", ("eq", "eq", "eq"), (48, 40, 128)),
    ("custom", ("log", "eq", "eq"), (36, 32, 64)),
    ("custom", ("eq", "log", "log"), (24, 44, 128)),
    ("unit", ("log", "log", "log"), (20, 24, 64)),
    ("singular", ("eq", "eq", "eq"), (33, 28, 64)),
]


@pytest.mark.parametrize("metric,kinds,shape", CASES)
def test_fused_laplacian_equals_oracle_and_six_pass(gpu_ready, monkeypatch, metric, kinds, shape):
    lows = (0.0, 0.0, 0.0) if metric == "singular" else (0.5, 0.25, 0.0)
    axes, spec = _axes(kinds, shape, lows)
    fns = METRICS[metric]
    g = ng.CurvilinearGrid(*axes, scale_factors=fns)
    rng = np.random.default_rng(zlib.crc32(repr((metric, kinds, shape)).encode()))
    X, Y, Z = g.meshed_coords
    f = np.sin(3 * X + 0.2) * np.cos(2 * Y) * np.exp(0.3 * Z) + 0.05 * rng.standard_normal(shape)
    out = g.laplacian(f)
    kernel = _lib.last_kernel()
    gen = metric != "unit" or "log" in kinds
    assert kernel == ("curvlap3d<gen>" if gen else "curvlap3d<unit>"), kernel
    ref = _oracle(spec, fns, f)
    assert np.array_equal(out, ref), (metric, kinds, shape, rel_linf(out, ref),
                                      np.argwhere(out != ref)[:5].tolist(), int((out != ref).sum()))
    assert np.array_equal(out, _six_pass(g, f, monkeypatch))
    assert np.all(np.isfinite(out))


def test_fused_laplacian_non_finite_input_rows(gpu_ready, monkeypatch):
    """inf / nan in the field spread exactly as far as in the reference (two chained radius-2 stencils) and are
    then zeroed by the final select."""
    axes, spec = _axes(("eq", "eq", "eq"), (40, 32, 128), (0.5, 0.25, 0.0))
    fns = METRICS["custom"]
    g = ng.CurvilinearGrid(*axes, scale_factors=fns)
    rng = np.random.default_rng(9)
    f = rng.standard_normal(g.shape)
    f[7, 9, 70] = np.inf
    f[20, 31, 0] = np.nan
    f[39, 0, 127] = -np.inf
    out = g.laplacian(f)
    assert _lib.last_kernel() == "curvlap3d<gen>"
    ref = _oracle(spec, fns, f)
    assert np.array_equal(out, ref), int((out != ref).sum())
    assert np.array_equal(out, _six_pass(g, f, monkeypatch))


def test_fused_laplacian_512_cube_vs_c_oracle_blocks(gpu_ready):
    """cfg5 in the reference's second idiom at 512^3: unit-scale CurvilinearGrid.laplacian = sum_i D_i(D_i f) with
    first-derivative stencils.  x blocks at both faces and in the middle against the C oracle's chained first
    derivatives (oracle/pencil_c.c::po_fd_apply, pinned bit for bit to the NumPy restatement)."""
    import torch
    from oracle import pencil_c
    n = 512
    ax = ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0)
    g = ng.CurvilinearGrid(ax, ax, ax, scale_factors=(ONE, ONE, ONE))
    x = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device="cuda")
    f = (torch.sin(2 * np.pi * x)[:, None, None] * torch.cos(2 * np.pi * x)[None, :, None] * torch.exp(x)[None, None, :]).contiguous()
    out = g.laplacian(f)
    assert _lib.last_kernel() == "curvlap3d<unit>"
    h = ax.spacing
    pencil_c.use_all_cores()
    for a, b in ((0, 16), (248, 264), (496, 512)):
        lo, hi = max(a - 12, 0), min(b + 12, n)                    # two chained stencils: an artificial cut must be > 8 away
        blk = f[lo:hi].cpu().numpy()
        ref = np.zeros_like(blk)
        for axis in range(3):
            ref = ref + pencil_c.fd_apply(1.0 * pencil_c.fd_apply(blk, axis, h, 1, 4), axis, h, 1, 4)
        ref = ref / 1.0
        ref = np.where(np.isfinite(ref), ref, 0.0)[a - lo:b - lo]
        got = out[a:b].cpu().numpy()
        assert np.array_equal(got, ref), ((a, b), rel_linf(got, ref), int((got != ref).sum()))
