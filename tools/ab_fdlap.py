"""A/B of two fused-Laplacian details on one GPU, interleaved so that clock drift hits both arms:
NG_FDLAP_ZSINGLE (1 = boundary lane computes its 2R one-sided z sums alone; 0 = spread over lanes)
and NG_FDLAP_HALO3D (1 = slab halos as {2, ROWS, 1} boxes; 0 = one contiguous row per plane)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

shape = tuple(int(s) for s in (sys.argv[1:4] if len(sys.argv) >= 4 else (2048, 2048, 256)))
lap = SlabLaplacian(shape, [1.0 / (n - 1) for n in shape])
f = torch.randn(*shape, dtype=torch.float64, device="cuda")
out = torch.empty_like(f)
rows = shape[0] * shape[1]
lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")


def run(fn, steps=6, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


cases = {
    "none": lambda: lap.apply_local(f, out),
    "hi": lambda: lap.apply_local(f, out, None, hi),
    "both": lambda: lap.apply_local(f, out, lo, hi),
    "part2(both)": lambda: lap.apply_part(f, out, 2, lo, hi),
}
for rep in range(3):
    for name, fn in cases.items():
        line = f"rep {rep} {name:12s}"
        for zs, h3 in ((0, 0), (1, 0), (0, 1), (1, 1)):
            os.environ["NG_FDLAP_ZSINGLE"], os.environ["NG_FDLAP_HALO3D"] = str(zs), str(h3)
            line += f"  zsingle={zs} halo3d={h3}: {run(fn):6.3f}"
        print(line, flush=True)
