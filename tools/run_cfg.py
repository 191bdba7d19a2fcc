"""Run one BASELINE config a few times on device tensors (for ncu): python tools/run_cfg.py cfg2|cfg3|cfg4 [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

A = ng.AxisType
which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which == "cfg2":
    ax = ng.create_axis(A.CHEBYSHEV, 256, 0.0, 1.0)
    g = ng.Grid(ax, ax, ax)
    f = torch.randn(256, 256, 256, dtype=torch.float64, device="cuda")
    ops = [ng.Diff(g, 2, 0), ng.Diff(g, 2, 2)]
    fn = lambda: [op(f) for op in ops]
elif which == "cfg3":
    e = ng.create_axis(A.EQUIDISTANT, 512, 0.0, 1.0)
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, 1024, 0.0, 2 * np.pi)
    g = ng.Grid(e, e, p)
    f = torch.randn(512, 512, 1024, dtype=torch.float64, device="cuda")
    op = ng.Diff(g, 1, 2)
    fn = lambda: op(f)
else:
    r = ng.create_axis(A.CHEBYSHEV, 128, 0.5, 2.0)
    th = ng.create_axis(A.CHEBYSHEV, 128, 0.3, np.pi - 0.3)
    ph = ng.create_axis(A.EQUIDISTANT_PERIODIC, 256, 0.0, 2 * np.pi)
    g = ng.SphericalGrid(r, th, ph)
    f = torch.randn(128, 128, 256, dtype=torch.float64, device="cuda")
    fn = lambda: (g.laplacian(f), g.gradient(f), g.divergence(f, f, f))
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("ok")
