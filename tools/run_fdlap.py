"""Run the fused Laplacian a few times at a given shape (for ncu): python tools/run_fdlap.py nx ny nz [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

shape = tuple(int(s) for s in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
lap = SlabLaplacian(shape, [1.0 / (n - 1) for n in shape])
f = torch.randn(*shape, dtype=torch.float64, device="cuda")
out = torch.empty_like(f)
for _ in range(reps):
    lap.apply_local(f, out)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
