"""cfg4 (SphericalGrid 128x128x256) composite timings on device tensors, optionally sweeping env knobs."""
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
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")


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


ref = None
for ncp in ("4", "1"):
    os.environ["NG_DENSE_NCP"] = ncp
    lap = timed(lambda: g.laplacian(f))
    grad = timed(lambda: g.gradient(f))
    div = timed(lambda: g.divergence(f, f, f))
    curl = timed(lambda: g.curl(f, f, f))
    out = g.laplacian(f)
    if ref is None:
        ref = out.clone()
    print(f"NCP={ncp}: laplacian {lap:.3f} ms  gradient {grad:.3f}  divergence {div:.3f}  curl {curl:.3f}  "
          f"bit-equal={torch.equal(out, ref)}")
