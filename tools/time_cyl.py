"""CylindricalGrid composites (phi in the MIDDLE: strided periodic axis) on device tensors: per-call times."""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib  # noqa: E402

A = ng.AxisType
shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 256, 256)
g = ng.CylindricalGrid(ng.create_axis(A.CHEBYSHEV, shape[0], 0.2, 2.0), ng.create_axis(A.EQUIDISTANT_PERIODIC, shape[1], 0.0, 2 * np.pi),
                       ng.create_axis(A.EQUIDISTANT, shape[2], 0.0, 1.0))
f = torch.randn(shape, dtype=torch.float64, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device="cuda")


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


d1 = ng.Diff(g, 1, 1).operator.operator
o = torch.empty_like(f)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("NG_"))
print(f"[{tag}] {shape}: Diff phi {timed(lambda: d1.apply_ptr(f.data_ptr(), o.data_ptr())) * 1e3:.1f} us  laplacian {timed(lambda: g.laplacian(f)):.3f} ms  "
      f"gradient {timed(lambda: g.gradient(f)):.3f}  divergence {timed(lambda: g.divergence(f, f, f)):.3f}  curl {timed(lambda: g.curl(f, f, f)):.3f}  "
      f"(last kernel of the Laplacian's phi pass: see NG traces)", flush=True)
