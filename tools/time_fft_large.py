"""Periodic derivative on long power-of-two axes (n > 1024: the shared-memory kernel of fft.cu) and n = 1000 / 1536 (zero-padded
2048 / 4096-point transforms): error against numpy.fft and timing."""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib, _tables  # noqa: E402

A = ng.AxisType
PEAK = 6550.4


def timed(fn, steps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


for n, batch in ((2048, 256), (4096, 256), (8192, 128), (1000, 512), (1536, 256), (16, 4096), (32, 4096)):
    for axis, shape in ((2, (batch, batch, n)), (1, (batch, n, batch))):
        axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == axis else A.EQUIDISTANT, m, 0.0, 2 * np.pi if a == axis else 1.0) for a, m in enumerate(shape)]
        g = ng.Grid(*axes)
        P = torch.from_numpy(axes[axis].coords).cuda()
        shp = [1, 1, 1]; shp[axis] = n
        f = (torch.exp(torch.sin(P)) * torch.cos(2 * P)).reshape(shp) * (1.0 + torch.rand(shape, dtype=torch.float64, device="cuda") * 1e-3)
        f = f.contiguous()
        op = ng.Diff(g, 1, axis)
        out = op(f)
        k = _lib.last_kernel()
        W = torch.from_numpy(_tables.fft_multiplier(n, axes[axis].spacing, 1)).cuda().reshape(shp)
        ref = torch.fft.ifft(torch.fft.fft(f, dim=axis) * W, dim=axis).real
        err = float((out - ref).abs().max() / ref.abs().max())
        o = torch.empty_like(f)
        dev = op.operator.operator
        ms = timed(lambda: dev.apply_ptr(f.data_ptr(), o.data_ptr()))
        print(f"n={n:5d} axis {axis} {shape}: [{k:24s}] err {err:.1e}  {ms:7.3f} ms  {f.numel() / ms / 1e6:6.1f} Gpts/s  {16 * f.numel() / ms / 1e6 / PEAK:4.2f} of HBM", flush=True)
        del f, o, out, ref
        torch.cuda.empty_cache()
