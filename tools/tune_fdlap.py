"""Sweep the fused-Laplacian kernel's knobs on one GPU (device-resident 1024^3, CUDA-event timing)."""
import os
import sys
import itertools

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

shape = tuple(int(s) for s in (sys.argv[1:4] if len(sys.argv) >= 4 else (1024, 1024, 1024)))
hs = [1.0 / (n - 1) for n in shape]
lap = SlabLaplacian(shape, hs)
f = torch.randn(*shape, dtype=torch.float64, device="cuda")
out = torch.empty_like(f)
pts = f.numel()


def run(steps=8, warm=2):
    for _ in range(warm):
        lap.apply_local(f, out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        lap.apply_local(f, out)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


os.environ["NG_FDLAP_GENERIC"] = "1"
ms = run(3, 1)
ref = out.clone()
print(f"generic kernel           : {ms:8.3f} ms  {16 * pts / ms / 1e6:8.1f} GB/s")
os.environ["NG_FDLAP_GENERIC"] = "0"
os.environ["NG_FDLAP_V"] = os.environ.get("TUNE_V", "4")
for (R, WY, S), xc in itertools.product(((2, 8, 6), (2, 4, 6), (2, 4, 8), (1, 8, 6), (4, 4, 6)), (8, 16, 24, 32, 48, 64)):
    os.environ["NG_FDLAP_R"] = str(R)
    os.environ["NG_FDLAP_WY"] = str(WY)
    os.environ["NG_FDLAP_S"] = str(S)
    os.environ["NG_FDLAP_XCHUNK"] = str(xc)
    ms = run()
    ok = torch.equal(out, ref)
    print(f"v{os.environ['NG_FDLAP_V']} R={R} WY={WY} S={S} xchunk={xc:5d}    : {ms:8.3f} ms  {16 * pts / ms / 1e6:8.1f} GB/s  bit-equal-to-generic={ok}")
