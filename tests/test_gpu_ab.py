"""A/B parity of the fast kernels against the simple kernels they replace (same library, other code path):
streaming finite-difference kernels vs the gather kernel (bit for bit), the general Chebyshev kernel vs the
fallback (bit for bit), the strided-pencil FFT kernel vs the radix-2 shared-memory kernel (<= 5e-13).
Each tool draws random shapes / orders / fused operands; the simple kernels are pinned to the oracle and the
golden vectors by the other GPU tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tool,cases,seed", [("check_fd_stream.py", 120, 21), ("check_dense_general.py", 80, 22),
                                             ("check_fft_strided.py", 50, 23)])
def test_fast_kernels_equal_simple_kernels(gpu_ready, tool, cases, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(cases), str(seed)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout


def test_fused_laplacian_kernels_equal_the_simple_kernel(gpu_ready):
    """Plane-ring kernel (every ring depth, several x chunks), register-march fallback, slab halos and the interior /
    boundary split, all against the thread-per-point kernel: tools/check_fdlap.py."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_fdlap.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bad 0" in r.stdout


@pytest.mark.parametrize("kind", ["spherical", "spherical_singular", "cylindrical", "polar"])
def test_laplacian_double_apply_equals_two_launches(gpu_ready, kind):
    """Periodic axes of the curvilinear Laplacian run D, the (J/h^2) multiply and D again in ONE launch (FuseDev::mid,
    csrc/fft2.cu); NG_FFT_TWICE=0 restores the two launches through the workspace array.  Same arithmetic in the same
    order, so the results must agree bit for bit.  On a grid whose coefficient table is non-finite on an axis (theta = 0:
    the reference's inf / nan rows before its isfinite select, numgrids/curvilinear.py:236-244) the partner pencil of a
    poisoned pair is redone by the warp-cooperative radix-2 routine -- twice in one launch, once with two launches --
    so there the agreement is to rounding (<= 1e-13 of the result's L-inf norm; the path's tolerance is 1e-11)."""
    import numpy as np
    import torch
    import numgrids_b200 as ng
    from numgrids_b200 import _lib
    A = ng.AxisType
    if kind.startswith("spherical"):
        th0 = 0.0 if kind.endswith("singular") else 0.3
        g = ng.SphericalGrid(ng.create_axis(A.CHEBYSHEV, 20, 0.5, 2.0), ng.create_axis(A.CHEBYSHEV, 18, th0, np.pi - th0),
                             ng.create_axis(A.EQUIDISTANT_PERIODIC, 256, 0, 2 * np.pi))
    elif kind == "cylindrical":           # phi in the middle: strided pencils
        g = ng.CylindricalGrid(ng.create_axis(A.CHEBYSHEV, 16, 0.2, 2.0), ng.create_axis(A.EQUIDISTANT_PERIODIC, 128, 0, 2 * np.pi),
                               ng.create_axis(A.EQUIDISTANT, 22, 0.0, 1.0))
    else:
        g = ng.PolarGrid(ng.create_axis(A.CHEBYSHEV, 40, 0.0, 2.0), ng.create_axis(A.EQUIDISTANT_PERIODIC, 64, 0, 2 * np.pi))
    M = g.meshed_coords
    f = np.sin(M[0]) * np.cos(2 * M[1]) + (M[-1] if len(M) == 3 else M[0]) ** 2
    f[(3,) * f.ndim] = np.inf
    fd = torch.from_numpy(f).cuda()
    lib = _lib.load()
    outs = {}
    try:
        os.environ["NG_CURV_SQUARE"] = "0"          # (the default evaluates c * (D D f) as ONE transform, see below)
        for mode in ("1", "0"):
            os.environ["NG_FFT_TWICE"] = mode
            lib.ng_tuning_reload()
            n0 = _lib.launch_count()
            outs[mode] = g.laplacian(fd).cpu().numpy()
            outs[mode + "n"] = _lib.launch_count() - n0
        os.environ.pop("NG_CURV_SQUARE")
        lib.ng_tuning_reload()
        n0 = _lib.launch_count()
        outs["sq"] = g.laplacian(fd).cpu().numpy()
        outs["sqn"] = _lib.launch_count() - n0
    finally:
        os.environ.pop("NG_FFT_TWICE", None)
        os.environ.pop("NG_CURV_SQUARE", None)
        lib.ng_tuning_reload()
    # default path: the built-in metrics' J/h_phi**2 does not vary along phi, so D(c D f) = c (D D f) is ONE transform
    # with the squared multiplier (ng_curv_set_axis_square): same operator, same launch count as the double apply,
    # agreement to rounding
    assert outs["sqn"] == outs["1n"], (outs["sqn"], outs["1n"])
    assert np.all(np.isfinite(outs["sq"]))
    assert float(np.max(np.abs(outs["sq"] - outs["1"]))) <= 1e-12 * float(np.max(np.abs(outs["1"])))
    assert outs["1n"] == outs["0n"] - 1, (outs["1n"], outs["0n"])
    assert np.all(np.isfinite(outs["1"])) and np.all(np.isfinite(outs["0"]))
    if kind.endswith("singular"):
        assert float(np.max(np.abs(outs["1"] - outs["0"]))) <= 1e-13 * float(np.max(np.abs(outs["0"])))
    else:
        assert np.array_equal(outs["1"], outs["0"])
