"""Run one finite-difference Diff a few times on a device array (for ncu): python tools/run_fd.py n0 n1 n2 axis [order] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

shape = tuple(int(x) for x in sys.argv[1:4])
axis = int(sys.argv[4])
order = int(sys.argv[5]) if len(sys.argv) > 5 else 1
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
g = ng.Grid(*(ng.create_axis(ng.AxisType.EQUIDISTANT, n, 0.0, 1.0) for n in shape))
f = torch.rand(shape, dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
op = ng.Diff(g, order, axis).operator.operator
for _ in range(reps):
    op.apply_ptr(f.data_ptr(), o.data_ptr())
torch.cuda.synchronize()
print("ok")
