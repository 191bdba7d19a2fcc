"""cfg4 gradient a few times (for ncu: the phi pass is the fused FFT apply with a post table + finite select)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

A = ng.AxisType
r = ng.create_axis(A.CHEBYSHEV, 128, 0.5, 2.0)
th = ng.create_axis(A.CHEBYSHEV, 128, 0.3, np.pi - 0.3)
ph = ng.create_axis(A.EQUIDISTANT_PERIODIC, 256, 0.0, 2 * np.pi)
g = ng.SphericalGrid(r, th, ph)
f = torch.randn(128, 128, 256, dtype=torch.float64, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    g.gradient(f)
torch.cuda.synchronize()
print("ok")
