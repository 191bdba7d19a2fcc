"""Chebyshev Diff at sizes off the 128-multiples (fast-kernel eligibility): ms and fraction of the FP64 pipe at 1.965 GHz."""
import os
import sys
import statistics

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

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


for n in (64, 96, 100, 128, 160, 192, 200, 256):
    for axis in (0, 2):
        shape = [256, 256, 256]
        shape[axis] = n
        axes = [ng.create_axis(A.CHEBYSHEV if a == axis else A.EQUIDISTANT, m, 0.0, 1.0) for a, m in enumerate(shape)]
        g = ng.Grid(*axes)
        f = torch.rand(shape, dtype=torch.float64, device="cuda")
        o = torch.empty_like(f)
        op = ng.Diff(g, 1, axis).operator.operator
        ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()))
        instr = 2.0 * n * f.numel()
        print(f"cheb n={n} axis {axis} shape {tuple(shape)}: {ms:.4f} ms  {f.numel() / ms / 1e6:.1f} Gpts/s  "
              f"{instr / ms / 1e9 / (148 * 64 * 1.965e-3):.2f} of the FP64 pipe", flush=True)
