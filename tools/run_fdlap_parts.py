"""Boundary-tile launches of the slab Laplacian for ncu: python tools/run_fdlap_parts.py nx ny nz mode
mode: halo (both neighbour planes), onesided (global boundary on both sides), interior (tiles 1 and nt-2)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

shape = tuple(int(s) for s in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "halo"
lap = SlabLaplacian(shape, [1.0 / (n - 1) for n in shape])
f = torch.randn(*shape, dtype=torch.float64, device="cuda")
out = torch.empty_like(f)
rows = shape[0] * shape[1]
lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
for _ in range(3):
    if mode == "halo":
        lap.apply_part(f, out, 2, lo, hi)
    elif mode == "onesided":
        lap.apply_part(f, out, 2)
    else:
        lap.apply_part(f, out, 3)
torch.cuda.synchronize()
print("ok")
