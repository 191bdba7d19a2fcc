"""CPU: the C-ABI library loads, exports every symbol include/numgrids_b200.h declares, validates
arguments, and fails loudly (never falls back) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import numgrids_b200._lib as L

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "numgrids_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ng_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = header_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature"
    assert set(L.SIGNATURES) == set(names)
    assert lib.ng_abi_version() == 1


def test_fuse_struct_layout_matches_header():
    assert C.sizeof(L.NgCoef) == 8 + 8 * L.NG_MAX_DIM
    assert C.sizeof(L.NgFuse) == 3 * C.sizeof(L.NgCoef) + 8 + 8 + 8 + 8   # ptrs + padded int32s


def _fd_create(shape, axis, wc, oc, wf, of, wb, ob):
    lib = L.load()
    shp, shp_p = L.as_i64(shape)
    a = [L.as_f64(x) for x in (wc, wf, wb)]
    o = [L.as_i32(x) for x in (oc, of, ob)]
    out = C.c_void_p()
    rc = lib.ng_fd_plan_create(len(shape), shp_p, axis, len(wc), a[0][1], o[0][1], len(wf), a[1][1],
                               o[1][1], a[2][1], o[2][1], 1.0, None, C.byref(out))
    return rc, out


def test_fd_plan_validation_without_gpu():
    lib = L.load()
    wc, oc = [0.5, 0.0, -0.5], [-1, 0, 1]
    wf, of = [-1.5, 2.0, -0.5], [0, 1, 2]
    wb, ob = [1.5, -2.0, 0.5], [0, -1, -2]
    rc, plan = _fd_create([8, 8], 1, wc, oc, wf, of, wb, ob)
    assert rc == 0 and plan.value
    assert lib.ng_plan_halo_width(plan) == 1
    assert lib.ng_plan_device_bytes(plan) == 0
    assert lib.ng_plan_destroy(plan) == 0
    rc, _ = _fd_create([8, 2], 1, wc, oc, wf, of, wb, ob)          # axis too short for the stencil
    assert rc == 1 and b"too short" in lib.ng_last_error()
    rc, _ = _fd_create([8, 8], 2, wc, oc, wf, of, wb, ob)          # axis out of range
    assert rc == 1
    rc, _ = _fd_create([8, 8, 8, 8, 8], 0, wc, oc, wf, of, wb, ob)  # ndim > NG_MAX_DIM
    assert rc == 1


@pytest.mark.skipif(L.load().ng_device_count() > 0, reason="this check is for GPU-less hosts")
def test_compute_fails_loudly_without_gpu():
    import numgrids_b200 as ng
    g = ng.Grid(ng.create_axis(ng.AxisType.EQUIDISTANT, 16, 0, 1))
    with pytest.raises(ng.NumgridsB200Error):
        ng.Diff(g, 1, 0)(np.zeros(16))
    gc = ng.Grid(ng.create_axis(ng.AxisType.CHEBYSHEV, 16, 0, 1))
    with pytest.raises(ng.NumgridsB200Error):
        ng.Diff(gc, 1, 0)(np.zeros(16))
    with pytest.raises(ng.NumgridsB200Error):
        ng.cartesian_laplacian(ng.Grid(*[ng.create_axis(ng.AxisType.EQUIDISTANT, 16, 0, 1)] * 2), np.zeros((16, 16)))
    rc, plan = _fd_create([8, 8], 1, [0.5, 0.0, -0.5], [-1, 0, 1], [-1.5, 2.0, -0.5], [0, 1, 2],
                          [1.5, -2.0, 0.5], [0, -1, -2])
    a = np.zeros((8, 8))
    b = np.zeros((8, 8))
    assert L.load().ng_apply_host(plan, a.ctypes.data, b.ctypes.data) == 2      # NG_ECUDA
    assert b"no CPU fallback" in L.load().ng_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "numgrids_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn
