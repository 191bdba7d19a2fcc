"""Run one fused finite-difference pass a few times (for ncu): python tools/run_fd_fused.py n axis post|addend|div"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib  # noqa: E402

n, axis, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
lib = _lib.load()
g = ng.Grid(*(ng.create_axis(ng.AxisType.EQUIDISTANT, n, 0.0, 1.0) for _ in range(3)))
f = torch.rand(n, n, n, dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
add = torch.rand_like(f)
one = torch.ones(1, dtype=torch.float64, device="cuda")
fu = _lib.NgFuse()
c = _lib.NgCoef(); c.ptr = one.data_ptr()
if kind == "post":
    fu.post = c
else:
    fu.addend = add.data_ptr()
    if kind == "div":
        fu.divisor = c
        fu.finite = 1
op = ng.Diff(g, 1, axis).operator.operator
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.ng_apply_fused(op.plan.handle, f.data_ptr(), o.data_ptr(), C.byref(fu), st), "apply")
torch.cuda.synchronize()
print("ok")
