"""Periodic axes whose length is not a power of two: the native mixed-radix transform (csrc/fft_mixed.cu) against
torch.fft's n-point transform, and its time next to the zero-padded power-of-two transform and the dense contraction."""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib, _tables  # noqa: E402
from numgrids_b200.diff import FFTDiff  # noqa: E402

A = ng.AxisType


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


def make(n, axis, shape, order, mode):
    FFTDiff.NONPOW2 = "padded" if mode == "dense" else mode
    FFTDiff.PADDED_ORDERS = () if mode == "dense" else (1, 2)
    FFTDiff.MIXED_MIN_POINTS = 3
    axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == axis else A.EQUIDISTANT, m, 0.0, 2 * np.pi if a == axis else 1.0) for a, m in enumerate(shape)]
    g = ng.Grid(*axes)
    return g, axes, ng.Diff(g, order, axis)


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
cases = [(6, 2048), (12, 2048), (24, 1024), (48, 1024), (60, 1024)] if quick else [(96, 512), (100, 512), (360, 512), (500, 512), (1000, 512), (1536, 256), (3000, 128), (45, 1024), (182, 512), (6, 2048), (81, 512), (2 * 3 * 5 * 7 * 11, 128)]
worst = 0.0
for n, batch in cases:
    for axis, shape in ((2, (batch, batch, n)), (1, (batch, n, batch)), (0, (n, batch, batch))):
        if quick and axis == 0:
            continue
        for order in (1, 2, 3):
            g, axes, op = make(n, axis, shape, order, "native")
            P = torch.from_numpy(axes[axis].coords).cuda()
            shp = [1, 1, 1]; shp[axis] = n
            f = (torch.exp(torch.sin(P)) * torch.cos(2 * P)).reshape(shp) * (1.0 + torch.rand(shape, dtype=torch.float64, device="cuda") * 1e-3)
            f = f.contiguous()
            out = op(f)
            k = _lib.last_kernel()
            W = torch.from_numpy(_tables.fft_multiplier(n, axes[axis].spacing, order)).cuda().reshape(shp)
            ref = torch.fft.ifft(torch.fft.fft(f, dim=axis) * W, dim=axis).real
            err = float((out - ref).abs().max() / ref.abs().max())
            worst = max(worst, err)
            line = f"n={n:5d} axis {axis} order {order} {str(shape):18s}: [{k:12s}] err {err:.1e}"
            if order == 1 or (order == 3 and axis == 2):
                o = torch.empty_like(f)
                dev = op.operator.operator
                ms = timed(lambda: dev.apply_ptr(f.data_ptr(), o.data_ptr()))
                line += f"  {ms:7.3f} ms {f.numel() / ms / 1e6:6.1f} Gpts/s"
                g2, _, op2 = make(n, axis, shape, order, "padded")
                out2 = op2(f)
                k2 = _lib.last_kernel()
                err2 = float((out2 - ref).abs().max() / ref.abs().max())
                dev2 = op2.operator.operator
                ms2 = timed(lambda: dev2.apply_ptr(f.data_ptr(), o.data_ptr()))
                line += f"   | [{k2:22s}] err {err2:.1e} {ms2:7.3f} ms {f.numel() / ms2 / 1e6:6.1f} Gpts/s"
                if n <= 200:
                    _, _, op3 = make(n, axis, shape, order, "dense")
                    op3(f)
                    k3 = _lib.last_kernel()
                    dev3 = op3.operator.operator
                    ms3 = timed(lambda: dev3.apply_ptr(f.data_ptr(), o.data_ptr()))
                    line += f"   | [{k3:22s}] {ms3:7.3f} ms {f.numel() / ms3 / 1e6:6.1f} Gpts/s"
                del o, out2
            print(line, flush=True)
            del f, out, ref
            torch.cuda.empty_cache()
# a pencil with a non-finite sample comes out all-NaN and leaves its partner alone
FFTDiff.NONPOW2 = "native"
g, axes, op = make(360, 2, (8, 9, 360), 1, "native")
f = torch.rand((8, 9, 360), dtype=torch.float64, device="cuda")
clean = op(f)
f2 = f.clone(); f2[3, 4, 17] = float("inf"); f2[7, 8, 0] = float("nan")
bad = op(f2)
m = torch.ones((8, 9), dtype=torch.bool, device="cuda"); m[3, 4] = False; m[7, 8] = False
assert float((bad[m] - clean[m]).abs().max()) <= 1e-12 * float(clean.abs().max()) and bool(torch.isnan(bad[3, 4]).all()) and bool(torch.isnan(bad[7, 8]).all())
print(f"non-finite pencils: ok; worst rel err vs torch.fft {worst:.2e}")
