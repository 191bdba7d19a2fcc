"""Time cfg4-size Chebyshev applies / the spherical Laplacian with one library variant (NUMGRIDS_B200_LIB)."""
import os
import sys
import statistics

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
o = torch.empty_like(f)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
d0 = ng.Diff(g, 1, 0).operator.operator
d1 = ng.Diff(g, 1, 1).operator.operator
ax = ng.create_axis(A.CHEBYSHEV, 256, 0.0, 1.0)
g2 = ng.Grid(ax, ax, ax)
f2 = torch.randn(256, 256, 256, dtype=torch.float64, device="cuda")
o2 = torch.empty_like(f2)
c0 = ng.Diff(g2, 2, 0).operator.operator
c2 = ng.Diff(g2, 2, 2).operator.operator


def timed(fn, steps=15, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


a0 = timed(lambda: d0.apply_ptr(f.data_ptr(), o.data_ptr()))
a1 = timed(lambda: d1.apply_ptr(f.data_ptr(), o.data_ptr()))
lap = timed(lambda: g.laplacian(f))
b0 = timed(lambda: c0.apply_ptr(f2.data_ptr(), o2.data_ptr()))
b2 = timed(lambda: c2.apply_ptr(f2.data_ptr(), o2.data_ptr()))
chk = float(g.laplacian(f).double().sum().item())
print(f"{os.environ.get('NUMGRIDS_B200_LIB', 'default')[-24:]:>24s} NCP={os.environ.get('NG_DENSE_NCP', '-')}: cfg4 Diff axis0 {a0 * 1e3:.1f} us  axis1 {a1 * 1e3:.1f} us  "
      f"laplacian {lap:.3f} ms | cfg2 axis0 {b0:.3f} ms axis2 {b2:.3f} ms | checksum {chk!r}", flush=True)
