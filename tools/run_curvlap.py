"""A few applies of the unit-scale CurvilinearGrid.laplacian (single-pass kernel) for ncu: python tools/run_curvlap.py [n] [metric]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
metric = sys.argv[2] if len(sys.argv) > 2 else "unit"
ONE = lambda c: np.ones_like(c[0])
fns = (ONE, ONE, ONE) if metric == "unit" else (ONE, lambda c: c[0], ONE)
ax = ng.create_axis(ng.AxisType.EQUIDISTANT, n, 0.5, 1.5)
g = ng.CurvilinearGrid(ax, ax, ax, scale_factors=fns)
f = torch.rand(n, n, n, dtype=torch.float64, device="cuda")
for _ in range(4):
    out = g.laplacian(f)
torch.cuda.synchronize()
print("done", float(out[3, 3, 3]))
