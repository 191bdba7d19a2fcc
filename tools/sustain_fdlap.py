"""Per-launch timings of the fused Laplacian over a sustained run (detect clock/power drift)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

shape = (1024, 1024, 1024)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
lap = SlabLaplacian(shape, [1.0 / (s - 1) for s in shape])
xs = [torch.linspace(0, 1, s, dtype=torch.float64, device="cuda") for s in shape]
f = (torch.sin(2 * np.pi * xs[0])[:, None, None] * torch.cos(2 * np.pi * xs[1])[None, :, None]
     * torch.exp(xs[2])[None, None, :]).contiguous()
out = torch.empty_like(f)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(n):
    lap.apply_local(f, out)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)])
for lo in range(0, n, max(1, n // 15)):
    seg = ts[lo:lo + max(1, n // 15)]
    print(f"launch {lo:4d}..: mean {seg.mean():.3f} ms  min {seg.min():.3f}  max {seg.max():.3f}")
print(f"overall: median {np.median(ts):.3f} ms, min {ts.min():.3f}, mean {ts.mean():.3f}")
