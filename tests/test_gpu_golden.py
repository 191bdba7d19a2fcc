"""GPU parity against the golden vectors produced by the REAL reference (tests/golden/*.npz).

Every call goes through the public drop-in API -> ctypes -> C ABI -> CUDA kernels.
Bars (BASELINE.json north_star): stencil and Chebyshev paths bit-exact (asserted with
np.array_equal, which implies the 1e-12 bound); FFT path rel. L-inf <= 1e-11.
"""
import numpy as np
import pytest

import numgrids_b200 as ng
from conftest import TOL_FFT, load_golden, rel_linf
from tests_support import build_axes, build_curv_grid

pytestmark = pytest.mark.gpu


def _check(out, ref, exact, what):
    assert out.shape == ref.shape and out.dtype == np.float64, what
    if exact:
        assert np.array_equal(out, ref), (what, rel_linf(out, ref))
    else:
        assert rel_linf(out, ref) <= TOL_FFT, (what, rel_linf(out, ref))


def test_diff_golden(gpu_ready):
    z, meta = load_golden("diff_cases.npz")
    grids = {}
    worst_fft = 0.0
    for m in meta:
        key = m["f"]
        if key not in grids:
            grids[key] = ng.Grid(*build_axes(ng, m["axes"]))
        g = grids[key]
        f = z[m["f"]]
        out = ng.Diff(g, m["order"], m["axis"], acc=m["acc"])(f)
        exact = m["kind"] != "periodic"
        if not exact:
            worst_fft = max(worst_fft, rel_linf(out, z[m["id"]]))
        _check(out, z[m["id"]], exact, m)
    print(f"worst FFT rel Linf = {worst_fft:.3e}")


def test_diff_golden_device_tensors(gpu_ready):
    """Same through CUDA torch tensors (device-resident in/out) and __cuda_array_interface__."""
    import torch
    z, meta = load_golden("diff_cases.npz")
    for m in meta[::7]:
        g = ng.Grid(*build_axes(ng, m["axes"]))
        f = torch.from_numpy(z[m["f"]]).cuda()
        op = ng.Diff(g, m["order"], m["axis"], acc=m["acc"])
        out = op(f)
        assert isinstance(out, torch.Tensor) and out.is_cuda and out.data_ptr() != f.data_ptr()
        _check(out.cpu().numpy(), z[m["id"]], m["kind"] != "periodic", m)
        assert torch.equal(f.cpu(), torch.from_numpy(z[m["f"]]))      # input not mutated

        class CAI:                                                     # bare __cuda_array_interface__
            def __init__(self, t):
                self.__cuda_array_interface__ = t.__cuda_array_interface__
        out2 = op(CAI(f))
        assert torch.equal(out2, out)


def test_curvilinear_golden(gpu_ready):
    z, meta = load_golden("curv_cases.npz")
    for m in meta:
        axes = build_axes(ng, m["axes"])
        g = build_curv_grid(ng, m["grid"], axes)
        nd = m["ndim"]
        cid = m["id"]
        has_fft = any(s[0] == "periodic" for s in m["axes"])
        f = z[f"{cid}_f"]
        v = [z[f"{cid}_v{i}"] for i in range(nd)]
        lap = g.laplacian(f)
        _check(lap, z[f"{cid}_lap"], not has_fft, (cid, m["grid"], "lap"))
        assert np.array_equal(g.laplacian(f), lap)                    # deterministic (reference tests/test_api.py:261-266)
        _check(g.divergence(*v), z[f"{cid}_div"], not has_fft, (cid, m["grid"], "div"))
        grad = g.gradient(f)
        assert isinstance(grad, tuple) and len(grad) == nd
        for i in range(nd):
            _check(grad[i], z[f"{cid}_grad{i}"], m["axes"][i][0] != "periodic", (cid, m["grid"], "grad", i))
        c = g.curl(*v)
        c = (c,) if nd == 2 else c
        for i in range(len(c)):
            _check(c[i], z[f"{cid}_curl{i}"], not has_fft, (cid, m["grid"], "curl", i))
        for arr in (lap, *grad, *c):
            assert np.all(np.isfinite(arr))


def test_curvilinear_no_warnings_at_singularities(gpu_ready):
    """reference tests/test_api.py:130-176: no RuntimeWarning escapes; no NaN/Inf in the result."""
    import warnings
    r = ng.create_axis(ng.AxisType.CHEBYSHEV, 12, 0.0, 1.0)
    th = ng.create_axis(ng.AxisType.CHEBYSHEV, 10, 0.0, np.pi)
    ph = ng.create_axis(ng.AxisType.EQUIDISTANT_PERIODIC, 8, 0, 2 * np.pi)
    g = ng.SphericalGrid(r, th, ph)
    R, T, P = g.meshed_coords
    f = R ** 2 * np.sin(T) ** 2
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        outs = [g.laplacian(f), *g.gradient(f), g.divergence(f, f, f), *g.curl(f, f, f)]
    for o in outs:
        assert np.all(np.isfinite(o))


def test_cartesian_laplacian_golden(gpu_ready):
    z, meta = load_golden("cartlap_cases.npz")
    for m in meta:
        g = ng.Grid(*build_axes(ng, m["axes"]))
        f = z[f"{m['id']}_f"]
        for acc in (2, 4, 6):
            out = ng.cartesian_laplacian(g, f, acc=acc)
            assert np.array_equal(out, z[f"{m['id']}_acc{acc}"]), (m, acc)
            # and the three-call reference idiom on the same device path
            idiom = None
            for a in range(g.ndims):
                d = ng.Diff(g, 2, a, acc=acc)(f)
                idiom = d if idiom is None else idiom + d
            assert np.array_equal(idiom, out)


def test_host_apply_staging_can_be_released(gpu_ready):
    """ng_apply_host keeps its two device staging arrays with the plan; ng_plan_release_staging hands them back and the
    next host apply allocates them again (same result)."""
    from numgrids_b200 import _lib
    ax = ng.create_axis(ng.AxisType.EQUIDISTANT, 40, 0.0, 1.0)
    g = ng.Grid(ax, ax)
    f = np.sin(3 * g.meshed_coords[0]) * g.meshed_coords[1]
    d = ng.Diff(g, 1, 0)
    first = d(f)
    op = d.operator.operator
    lib = _lib.load()
    assert lib.ng_plan_release_staging(op.plan.handle) == 0
    assert lib.ng_plan_release_staging(op.plan.handle) == 0        # idempotent
    assert np.array_equal(d(f), first)
