"""Launch the fused Laplacian once per shape (for ncu): python tools/run_lap_shapes.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for sh in ((1024, 1024, 1024), (2048, 2048, 256)):
    lap = SlabLaplacian(sh, [1.0 / (n - 1) for n in sh])
    f = torch.randn(*sh, dtype=torch.float64, device="cuda")
    out = torch.empty_like(f)
    for _ in range(reps):
        lap.apply_local(f, out)
    torch.cuda.synchronize()
    del f, out
print("ok")
