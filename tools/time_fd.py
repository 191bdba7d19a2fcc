"""Per-axis finite-difference Diff (fd_apply_kernel) on large device arrays: ms, Gpts/s, fraction of the measured HBM peak."""
import os
import sys
import statistics

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

PEAK = 6550.4
A = ng.AxisType


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


for shape in ((1024, 1024, 1024), (512, 512, 1024), (8192, 8192), (256, 256, 256)):
    axes = [ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0) for n in shape]
    g = ng.Grid(*axes)
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    o = torch.empty_like(f)
    for order in (1, 2):
        for axis in range(len(shape)):
            op = ng.Diff(g, order, axis).operator.operator
            ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()))
            print(f"fd {shape} order {order} axis {axis}: {ms:.4f} ms  {f.numel() / ms / 1e6:.1f} Gpts/s  "
                  f"{16 * f.numel() / ms / 1e6:.0f} GB/s = {16 * f.numel() / ms / 1e6 / PEAK:.2f} of measured HBM peak", flush=True)
    lg = ng.Grid(ng.create_axis(A.LOGARITHMIC, shape[0], 0.1, 10.0), *axes[1:])
    op = ng.Diff(lg, 1, 0).operator.operator
    ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()))
    print(f"log {shape} order 1 axis 0: {ms:.4f} ms  {16 * f.numel() / ms / 1e6 / PEAK:.2f} of measured HBM peak", flush=True)
    del f, o
