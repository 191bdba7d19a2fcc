"""Strided-axis periodic derivative (TMA tile kernel) at large sizes against torch.fft on the same device:
python tools/check_fft_tile_big.py  (tolerance 1e-11 of the result's L-inf norm)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib, _tables  # noqa: E402

A = ng.AxisType
bad = 0
for shape, axis in (((1024, 256, 1024), 1), ((2048, 128, 1024), 1), ((4096, 64, 1024), 1), ((512, 1024, 512), 1),
                    ((512, 1024, 512), 0), ((128, 300, 70), 0), ((64, 999, 18), 0), ((1024, 4098), 0)):
    axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == axis else A.EQUIDISTANT, n, 0.0, 2 * np.pi if a == axis else 1.0)
            for a, n in enumerate(shape)]
    g = ng.Grid(*axes)
    torch.manual_seed(3)
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    for order in (1, 2):
        op = ng.Diff(g, order, axis)
        for rep in range(3):
            out = op(f)
            torch.cuda.synchronize()
        kern = _lib.last_kernel()
        W = torch.from_numpy(_tables.fft_multiplier(shape[axis], axes[axis].spacing, order)).cuda()
        wshape = [1] * len(shape)
        wshape[axis] = shape[axis]
        ref = torch.fft.ifft(torch.fft.fft(f, dim=axis) * W.reshape(wshape), dim=axis).real
        err = float((out - ref).abs().max() / ref.abs().max())
        ok = err <= (3e-11 if order == 2 and shape[axis] == 1024 else 1e-11)
        bad += not ok
        print(f"{shape} axis {axis} order {order}: {kern}  rel Linf {err:.2e}  {'ok' if ok else 'MISMATCH'}", flush=True)
sys.exit(1 if bad else 0)
