"""A few applies of the periodic derivative for ncu: python tools/run_fft_any.py n axis batch"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib  # noqa: E402
from numgrids_b200.diff import FFTDiff  # noqa: E402

n, axis, batch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4:
    FFTDiff.NONPOW2 = sys.argv[4]
shape = [batch, batch, batch]
shape[axis] = n
A = ng.AxisType
axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == axis else A.EQUIDISTANT, m, 0.0, 2 * np.pi if a == axis else 1.0)
        for a, m in enumerate(shape)]
g = ng.Grid(*axes)
f = torch.rand(shape, dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
op = ng.Diff(g, 1, axis).operator.operator
for _ in range(3):
    op.apply_ptr(f.data_ptr(), o.data_ptr())
torch.cuda.synchronize()
print("done", _lib.last_kernel(), float(o[3, 3, 3]))
