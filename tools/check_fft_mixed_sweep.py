"""Random sweep of the native mixed-radix transform: random lengths n = 2^a 3^b 5^c 7^d 11^e 13^f <= 8192, random axis position and
batch extents (odd pencil counts, ragged last CTA), first derivative against torch.fft's n-point transform.
python tools/check_fft_mixed_sweep.py [cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib, _tables  # noqa: E402
from numgrids_b200.diff import FFTDiff  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
FFTDiff.NONPOW2 = "native"
FFTDiff.MIXED_MIN_POINTS = 3
A = ng.AxisType
smooth = [n for n in range(3, 8193) if _tables.fft_mixed_supported(n)]
worst, done = 0.0, 0
while done < cases:
    n = int(smooth[int(rng.integers(0, len(smooth)))]) if rng.random() < 0.7 else int(smooth[int(rng.integers(0, 200))])
    ndim = int(rng.integers(1, 4))
    axis = int(rng.integers(0, ndim))
    budget = max(1, (1 << 21) // n)
    shape = []
    for a in range(ndim):
        if a == axis:
            shape.append(n)
        else:
            e = int(rng.integers(1, max(2, min(40, budget) + 1)))
            budget = max(1, budget // e)
            shape.append(e)
    axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == axis else A.EQUIDISTANT, m, 0.0, 2 * np.pi if a == axis else 1.0) if m > 1 or a == axis
            else ng.create_axis(A.EQUIDISTANT, 2, 0.0, 1.0) for a, m in enumerate(shape)]
    shape = [len(ax) for ax in axes]
    g = ng.Grid(*axes)
    f = torch.rand(shape, dtype=torch.float64, device="cuda") - 0.5
    out = ng.Diff(g, 1, axis)(f)
    assert _lib.last_kernel() == "fft_mixed", (_lib.last_kernel(), n)
    W = torch.from_numpy(_tables.fft_multiplier(n, axes[axis].spacing, 1)).cuda().reshape([n if a == axis else 1 for a in range(ndim)])
    ref = torch.fft.ifft(torch.fft.fft(f, dim=axis) * W, dim=axis).real
    err = float((out - ref).abs().max() / ref.abs().max())
    worst = max(worst, err)
    assert err <= 1e-11, (n, shape, axis, err)
    done += 1
print(f"{cases} cases, worst rel. L-inf vs torch.fft {worst:.2e}")
