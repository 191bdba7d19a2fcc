"""Adaptive refinement: the reference's numgrids/amr.py loop on device arrays.

The golden histories (tests/golden/amr_cases.npz, make_golden_amr.py) come from the REAL reference.
Interpolation is bit-identical and the max norm is exact, so norm="max" histories must agree exactly;
the l2 / mean norms sum in a different order than NumPy's pairwise sum: relative 1e-12.
"""
import json

import numpy as np
import pytest

import numgrids_b200 as ng
from conftest import load_golden
from tests_support import amr_field, build_axes


def _close(a, b, norm):
    return a == b if norm == "max" else abs(a - b) <= 1e-12 * max(abs(b), 1e-300)


def test_estimator_argument_checks():
    ax = ng.create_axis(ng.AxisType.EQUIDISTANT, 8, 0.0, 1.0)
    with pytest.raises(ValueError, match="Unknown norm"):
        ng.ErrorEstimator(ng.Grid(ax), lambda g: g.meshed_coords[0], norm="l7")
    r = ng.AdaptationResult(grid=None, f=None, errors={}, global_error=0.0, iterations=0, converged=True)
    assert r.history == []
    z, meta = load_golden("amr_cases.npz")
    assert len(meta) >= 5 and all("history" in c for c in meta)


@pytest.mark.gpu
@pytest.mark.parametrize("device_func", [False, True])
def test_adapt_matches_reference_history(gpu_ready, device_func):
    import torch
    z, meta = load_golden("amr_cases.npz")
    for c in meta:
        case = c["case"]
        grid = ng.Grid(*build_axes(ng, case["axes"]))

        def func(g, name=case["field"]):
            f = amr_field(name, g.meshed_coords)
            return torch.from_numpy(f).cuda() if device_func else f

        res = ng.adapt(grid, func, tol=case["tol"], norm=case["norm"], **case["kw"])
        assert list(res.grid.shape) == c["final_shape"], case
        assert res.iterations == c["iterations"] and res.converged == c["converged"], case
        assert _close(res.global_error, c["global_error"], case["norm"]), (case, res.global_error, c["global_error"])
        for k, v in c["errors"].items():
            assert _close(res.errors[int(k)], v, case["norm"]), case
        assert len(res.history) == len(c["history"])
        for h, hr in zip(res.history, c["history"]):
            assert list(h["shape"]) == hr["shape"]
            assert _close(h["global_error"], hr["global_error"], case["norm"])
            assert sorted(int(k.rsplit("_", 1)[1]) for k in h if k.startswith("refined_axis_")) == hr["refined"]
            for k, v in hr["axis_errors"].items():
                assert _close(h["axis_errors"][int(k)], v, case["norm"])
        f = res.f.cpu().numpy() if device_func else res.f
        assert float(np.sum(f)) == c["f_checksum"]
        est = ng.estimate_error(grid, func, norm=case["norm"])
        assert _close(est["global"], c["estimate"]["global"], case["norm"])
        for k, v in c["estimate"]["per_axis"].items():
            assert _close(est["per_axis"][int(k)], v, case["norm"])
        e = ng.ErrorEstimator(grid, func, norm=case["norm"])
        worst = e.axis_needing_refinement()
        ref_axis = max(c["estimate"]["per_axis"], key=c["estimate"]["per_axis"].get)
        assert worst == int(ref_axis)


@pytest.mark.gpu
def test_diff_norm_kinds_against_numpy(gpu_ready):
    import torch
    from numgrids_b200.amr import _diff_norm
    rng = np.random.default_rng(2)
    for n in (1, 7, 64, 1001, 1 << 20, (1 << 22) + 3):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        d = a - b
        assert _diff_norm(da, db, 0) == float(np.max(np.abs(d)))
        assert abs(_diff_norm(da, db, 1) - float(np.sqrt(np.mean(d ** 2)))) <= 1e-13 * float(np.sqrt(np.mean(d ** 2)))
        assert abs(_diff_norm(da, db, 2) - float(np.mean(np.abs(d)))) <= 1e-13 * float(np.mean(np.abs(d)))
    a[5] = np.nan
    assert np.isnan(_diff_norm(torch.from_numpy(a).cuda(), db, 0))
