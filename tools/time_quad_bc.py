"""Time the quadrature pass (ng_quad_apply, 8 B/point) and the boundary kernels on device arrays."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

PEAK = 6550.4   # GB/s, MEASURED_PEAKS.json
A = ng.AxisType


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


for shape, kinds in (((1024, 1024, 1024), (A.EQUIDISTANT,) * 3), ((512, 512, 1024), (A.EQUIDISTANT, A.EQUIDISTANT, A.EQUIDISTANT_PERIODIC)),
                     ((256, 256, 256), (A.CHEBYSHEV,) * 3), ((128, 128, 256), (A.CHEBYSHEV, A.CHEBYSHEV, A.EQUIDISTANT_PERIODIC)),
                     ((4096, 4096), (A.EQUIDISTANT,) * 2), ((512, 512), (A.EQUIDISTANT,) * 2)):
    grid = ng.Grid(*(ng.create_axis(k, n, 0.0, 1.0) for k, n in zip(kinds, shape)))
    I = ng.Integral(grid)
    f = torch.rand(shape, dtype=torch.float64, device="cuda")
    med, mn = timed(lambda: I(f))
    gb = 8 * f.numel() / 1e9
    print(f"quad {shape}: {med:.4f} ms (min {mn:.4f})  {f.numel() / med / 1e6:.1f} Gpts/s  {gb / med * 1e3:.0f} GB/s = {gb / med * 1e3 / PEAK:.2f} of measured HBM peak  value {I(f).item():.15g}")
    w = [torch.from_numpy(x).cuda() for x in I.weights]
    chk = f
    for x in w:
        chk = torch.tensordot(x, chk, dims=([0], [0]))
    print(f"   torch tensordot check: rel diff {abs(chk.item() - I(f).item()) / abs(chk.item()):.2e}")
    if len(shape) == 3 and shape[0] >= 256:
        for axis in range(3):
            for cls, name in ((ng.DirichletBC, "dirichlet"), (ng.NeumannBC, "neumann")):
                bc = cls(grid.faces[f"{axis}_low"], 0.5) if not grid.axes[axis].periodic else None
                if bc is None:
                    continue
                med, mn = timed(lambda: bc.apply(f))
                face = f.numel() // shape[axis]
                print(f"   bc {name} axis {axis}: {med * 1e3:.1f} us  ({face} face points)")
    del f
