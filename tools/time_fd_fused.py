"""Fused finite-difference applies (the passes of a curvilinear composite) at 512^3: ms per pass and fraction of the
measured HBM peak for the bytes each pass really moves."""
import ctypes as C
import os
import sys
import statistics

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib  # noqa: E402

PEAK = 6550.4
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = _lib.load()
g = ng.Grid(*(ng.create_axis(ng.AxisType.EQUIDISTANT, n, 0.0, 1.0) for _ in range(3)))
f = torch.rand(n, n, n, dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
add = torch.rand_like(f)
one = torch.ones(1, dtype=torch.float64, device="cuda")
tab2 = torch.rand(n, n, dtype=torch.float64, device="cuda") + 0.5      # varies along axes 0 and 1


def coef(t, strides):
    c = _lib.NgCoef()
    c.ptr = t.data_ptr()
    for a in range(4):
        c.stride[a] = strides[a] if a < len(strides) else 0
    return c


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


cases = {}
fu = _lib.NgFuse(); fu.post = coef(one, (0, 0, 0)); cases["post scalar"] = (fu, 16)
fu = _lib.NgFuse(); fu.post = coef(tab2, (n, 1, 0)); cases["post [n0][n1] table"] = (fu, 16)
fu = _lib.NgFuse(); fu.pre = coef(tab2, (n, 1, 0)); cases["pre [n0][n1] table"] = (fu, 16)
fu = _lib.NgFuse(); fu.addend = add.data_ptr(); cases["addend"] = (fu, 24)
fu = _lib.NgFuse(); fu.addend = add.data_ptr(); fu.divisor = coef(one, (0, 0, 0)); fu.finite = 1; cases["addend + div scalar + finite"] = (fu, 24)
fu = _lib.NgFuse(); fu.addend = add.data_ptr(); fu.divisor = coef(tab2, (n, 1, 0)); fu.finite = 1; cases["addend + div table + finite"] = (fu, 24)
for axis in range(3):
    op = ng.Diff(g, 1, axis).operator.operator
    st = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: lib.ng_apply(op.plan.handle, f.data_ptr(), o.data_ptr(), st))
    print(f"axis {axis} plain: {ms:.3f} ms  {16 * f.numel() / ms / 1e6 / PEAK:.2f}")
    for name, (fu, bpp) in cases.items():
        ms = timed(lambda: lib.ng_apply_fused(op.plan.handle, f.data_ptr(), o.data_ptr(), C.byref(fu), st))
        print(f"axis {axis} {name}: {ms:.3f} ms  {bpp * f.numel() / ms / 1e6 / PEAK:.2f} of HBM peak at {bpp} B/pt", flush=True)
