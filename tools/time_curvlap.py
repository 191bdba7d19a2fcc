"""Timing of CurvilinearGrid.laplacian on all-finite-difference grids: single-pass kernel (csrc/curvlap3d.cu)
against the six-pass path, unit-scale (cfg5 in the reference's second idiom) and a general metric."""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib  # noqa: E402

A = ng.AxisType
ONE = lambda c: np.ones_like(c[0])
CUSTOM = (lambda c: 1.0 + c[0] ** 2, lambda c: c[0] * np.exp(c[1]), lambda c: 2.0 + np.sin(2 * np.pi * c[2]) * c[1])
PEAK = 6550.4


def timed(fn, steps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


cases = [("unit", (ONE, ONE, ONE), (256, 256, 256)), ("unit", (ONE, ONE, ONE), (512, 512, 512)),
         ("custom", CUSTOM, (256, 256, 256)), ("cylindrical-like", (ONE, lambda c: c[0], ONE), (512, 512, 512))]
if "--big" in sys.argv:
    cases.append(("unit", (ONE, ONE, ONE), (1024, 1024, 1024)))
for name, fns, shape in cases:
    axes = [ng.create_axis(A.EQUIDISTANT, n, 0.5, 1.5) for n in shape]
    g = ng.CurvilinearGrid(*axes, scale_factors=fns)
    if shape[0] >= 1024:            # the reference API evaluates the scale factors on the full meshgrid (24 GiB here):
        g._h_arrays = tuple(np.broadcast_to(1.0, shape) for _ in range(3))   # unit metric, same tables
        g._jacobian = np.broadcast_to(1.0, shape)
    f = torch.rand(*shape, dtype=torch.float64, device="cuda")
    pts = f.numel()
    for mode, cfg, xc in (("1", "4", "0"), ("1", "8", "0"), ("1", "2", "0"), ("1", "0", "0"), ("0", "0", "0")):
        os.environ["NG_CURV_FUSED"], os.environ["NG_CURV_CFG"], os.environ["NG_CURV_XCHUNK"] = mode, cfg, xc
        _lib.tuning_reload()
        med, mn = timed(lambda: g.laplacian(f))
        print(f"{name:16s} {shape}  fused={mode} cfg={cfg} xchunk={xc:>3s} [{_lib.last_kernel():16s}]  median {med:7.3f} ms  min {mn:7.3f} ms  "
              f"{pts / med / 1e6:7.1f} Gpts/s  {16 * pts / med / 1e6 / PEAK:5.3f} of HBM peak at 16 B/pt")
    for k in ("NG_CURV_FUSED", "NG_CURV_CFG", "NG_CURV_XCHUNK"):
        del os.environ[k]
    _lib.tuning_reload()
    del f, g
    torch.cuda.empty_cache()
