"""Compare the FFT kernel variants (one-shot / ring 6-6 / ring 8-4 / ring 16-8 where built) and check them against each other."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

b0, b1, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A = ng.AxisType
g = ng.Grid(ng.create_axis(A.EQUIDISTANT, b0, 0.0, 1.0), ng.create_axis(A.EQUIDISTANT, b1, 0.0, 1.0),
            ng.create_axis(A.EQUIDISTANT_PERIODIC, n, 0.0, 2 * np.pi))
f = torch.randn(b0, b1, n, dtype=torch.float64, device="cuda")
o = torch.empty_like(f)
op = ng.Diff(g, 1, 2).operator.operator
ref = None
for pf, var in (("0", "0"), ("1", "1"), ("1", "2"), ("1", "3")):
    os.environ["NG_FFT_PREFETCH"] = pf
    os.environ["NG_FFT_VARIANT"] = var
    for _ in range(3):
        op.apply_ptr(f.data_ptr(), o.data_ptr())
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        op.apply_ptr(f.data_ptr(), o.data_ptr())
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    if ref is None:
        ref = o.clone()
    same = float((o - ref).abs().max() / ref.abs().max())
    ms = float(np.median(ts))
    print(f"n={n} batch={b0}x{b1} prefetch={pf} variant={var}: {ms:.4f} ms  {f.numel() / ms / 1e6:.1f} Gpts/s  {16 * f.numel() / ms / 1e6:.0f} GB/s  maxdiff {same:.1e}")
