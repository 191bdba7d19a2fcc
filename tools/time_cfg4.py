"""cfg4 (SphericalGrid 128x128x256): per-axis Diff and composite timings on device tensors (L2 flushed between calls)."""
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


d2 = ng.Diff(g, 1, 2).operator.operator
a0 = timed(lambda: d0.apply_ptr(f.data_ptr(), o.data_ptr()))
a1 = timed(lambda: d1.apply_ptr(f.data_ptr(), o.data_ptr()))
a2 = timed(lambda: d2.apply_ptr(f.data_ptr(), o.data_ptr()))
lap = timed(lambda: g.laplacian(f))
grad = timed(lambda: g.gradient(f))
div = timed(lambda: g.divergence(f, f, f))
curl = timed(lambda: g.curl(f, f, f))
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("NG_"))
print(f"[{tag}] Diff axis0 {a0 * 1e3:.1f} us  axis1 {a1 * 1e3:.1f} us  axis2 {a2 * 1e3:.1f} us  laplacian {lap:.3f} ms  "
      f"gradient {grad:.3f}  divergence {div:.3f}  curl {curl:.3f}", flush=True)
