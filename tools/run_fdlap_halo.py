"""Run the fused Laplacian with both slab halos a few times (for ncu): python tools/run_fdlap_halo.py nx ny nz"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

shape = tuple(int(s) for s in sys.argv[1:4])
lap = SlabLaplacian(shape, [1.0 / (n - 1) for n in shape])
f = torch.randn(*shape, dtype=torch.float64, device="cuda")
out = torch.empty_like(f)
rows = shape[0] * shape[1]
lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
for _ in range(3):
    lap.apply_local(f, out, lo, hi)
torch.cuda.synchronize()
print("ok")
