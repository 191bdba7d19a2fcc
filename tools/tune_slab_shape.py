"""Sweep x-chunk length / tile configuration of the fused-Laplacian kernel on the 8-GPU slab shape
(2048 x 2048 x 256) against the 1-GPU shape (1024^3), interleaved so that clock drift hits both."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

shapes = [(1024, 1024, 1024), (2048, 2048, 256)]
laps, fs, outs = {}, {}, {}
for sh in shapes:
    laps[sh] = SlabLaplacian(sh, [1.0 / (n - 1) for n in sh])
    fs[sh] = torch.randn(*sh, dtype=torch.float64, device="cuda")
    outs[sh] = torch.empty_like(fs[sh])


def run(sh, steps=6, warm=2):
    for _ in range(warm):
        laps[sh].apply_local(fs[sh], outs[sh])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        laps[sh].apply_local(fs[sh], outs[sh])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for rep in range(3):
    for (R, WY, S), xc in itertools.product(((2, 4, 8),), (32, 48, 64, 96, 128)):
        os.environ["NG_FDLAP_R"], os.environ["NG_FDLAP_WY"], os.environ["NG_FDLAP_S"] = str(R), str(WY), str(S)
        os.environ["NG_FDLAP_XCHUNK"] = str(xc)
        t = [run(sh) for sh in shapes]
        print(f"rep {rep} R={R} WY={WY} S={S} xchunk={xc:5d}: 1024^3 {t[0]:6.3f} ms   2048x2048x256 {t[1]:6.3f} ms   ratio {t[1] / t[0]:.3f}", flush=True)
