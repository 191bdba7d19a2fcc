"""A few Chebyshev applies along axis 0 at cfg4's shape (for ncu): python tools/run_cheb.py [n0 n1 n2]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 128, 256)
A = ng.AxisType
axes = [ng.create_axis(A.CHEBYSHEV if a == 0 else A.EQUIDISTANT, m, 0.0, 1.0) for a, m in enumerate(shape)]
g = ng.Grid(*axes)
f = torch.rand(shape, dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
op = ng.Diff(g, 1, 0).operator.operator
for _ in range(4):
    op.apply_ptr(f.data_ptr(), o.data_ptr())
torch.cuda.synchronize()
print("done")
