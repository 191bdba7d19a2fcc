"""A few applies of the periodic derivative along the MIDDLE axis (TMA tile kernel, csrc/fft_tile.cu) for ncu:
python tools/run_fft_tile.py [n0 n n2]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 1024, 512)
A = ng.AxisType
axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == 1 else A.EQUIDISTANT, m, 0.0, 2 * np.pi if a == 1 else 1.0)
        for a, m in enumerate(shape)]
g = ng.Grid(*axes)
f = torch.rand(shape, dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
op = ng.Diff(g, 1, 1).operator.operator
for _ in range(4):
    op.apply_ptr(f.data_ptr(), o.data_ptr())
torch.cuda.synchronize()
print("done", float(o[3, 3, 3]))
