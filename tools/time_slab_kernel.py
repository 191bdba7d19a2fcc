"""Time the fused Laplacian kernel with and without slab halos on one GPU (same shape, same field)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

shape = tuple(int(s) for s in (sys.argv[1:4] if len(sys.argv) >= 4 else (1024, 1024, 1024)))
hs = [1.0 / (n - 1) for n in shape]
lap = SlabLaplacian(shape, hs)
xs = [torch.linspace(0, 1, n, dtype=torch.float64, device="cuda") for n in shape]
f = (torch.sin(2 * np.pi * xs[0])[:, None, None] * torch.cos(2 * np.pi * xs[1])[None, :, None]
     * torch.exp(xs[2])[None, None, :]).contiguous()
out = torch.empty_like(f)
rows = shape[0] * shape[1]
lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
pts = f.numel()


def run(halos, steps=10, warm=3):
    for _ in range(warm):
        lap.apply_local(f, out, *halos)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        lap.apply_local(f, out, *halos)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for rep in range(2):
    for name, halos in (("none", (None, None)), ("hi", (None, hi)), ("lo", (lo, None)), ("both", (lo, hi))):
        ms = run(halos)
        print(f"halos={name:5s}: {ms:7.3f} ms  {16 * pts / ms / 1e6:8.1f} GB/s")
# pack kernel
bufs = lap._buffers(f)
for _ in range(3):
    lap.pack(f)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    lap.pack(f)
b.record()
torch.cuda.synchronize()
print(f"pack kernel: {a.elapsed_time(b) / 10:7.3f} ms")


def run_parts(steps=10, warm=3):
    def once():
        lap.apply_part(f, out, 1, lo, hi)
        lap.apply_part(f, out, 2, lo, hi)
    for _ in range(warm):
        once()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        once()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def run_one_part(part, steps=10, warm=3):
    for _ in range(warm):
        lap.apply_part(f, out, part, lo, hi)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        lap.apply_part(f, out, part, lo, hi)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


ms = run_parts()
print(f"part 1 + part 2 (both halos): {ms:7.3f} ms  {16 * pts / ms / 1e6:8.1f} GB/s")
print(f"part 1 alone: {run_one_part(1):7.3f} ms   part 2 alone: {run_one_part(2):7.3f} ms")
