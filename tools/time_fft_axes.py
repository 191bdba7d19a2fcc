"""Periodic spectral Diff along each axis position (last axis: two-pass register kernel; other axes: fallback)."""
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
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


SHAPES = {(256, 256, 256): (0, 1, 2), (512, 1024, 512): (0, 1, 2), (128, 256, 128): (0, 1, 2),
          (1024, 256, 1024): (1,), (2048, 128, 1024): (1,), (4096, 64, 1024): (1,)}
for shape, which in SHAPES.items():
    for axis in which:
        axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == axis else A.EQUIDISTANT, n, 0.0, 2 * np.pi if a == axis else 1.0)
                for a, n in enumerate(shape)]
        g = ng.Grid(*axes)
        f = torch.rand(shape, dtype=torch.float64, device="cuda")
        o = torch.empty_like(f)
        op = ng.Diff(g, 1, axis).operator.operator
        ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()))
        print(f"fft {shape} n={shape[axis]} axis {axis}: {ms:.4f} ms  {f.numel() / ms / 1e6:.1f} Gpts/s  "
              f"{16 * f.numel() / ms / 1e6 / PEAK:.2f} of measured HBM peak", flush=True)
        del f, o
