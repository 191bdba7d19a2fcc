"""Periodic axes whose length is not a power of two: zero-padded transform (default for first derivatives) against the
dense O(n^2) contraction, Diff(g, 1, 2) on (b, b, n) grids: python tools/time_fft_anylen.py [b]"""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib  # noqa: E402
from numgrids_b200.diff import FFTDiff  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 512
PEAK = 6550.4


def timed(fn, steps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    return statistics.median(ts)


e = ng.create_axis(ng.AxisType.EQUIDISTANT, b, 0.0, 1.0)
for n in (100, 200, 360, 500, 1000, 1536):
    p = ng.create_axis(ng.AxisType.EQUIDISTANT_PERIODIC, n, 0.0, 2 * np.pi)
    g = ng.Grid(e, e, p)
    f = torch.randn(b, b, n, dtype=torch.float64, device="cuda")
    o = torch.empty_like(f)
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for label, min_pts in (("padded", 65), ("dense", 10 ** 9)):
        FFTDiff.PADDED_MIN_POINTS = min_pts
        op = ng.Diff(g, 1, 2).operator.operator
        ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr(), stream=st))
        res.append(f"{label} [{_lib.last_kernel():18s}] {ms:8.3f} ms {f.numel() / ms / 1e6:7.1f} Gpts/s {16 * f.numel() / ms / 1e6 / PEAK:5.2f} of HBM")
    FFTDiff.PADDED_MIN_POINTS = 65
    print(f"{b}x{b}x{n:5d}: " + "   ".join(res), flush=True)
    del f, o
    torch.cuda.empty_cache()
