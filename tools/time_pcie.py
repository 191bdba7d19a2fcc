"""PCIe ceiling of the host-buffer path: pinned H2D alone, D2H alone, both at once (two streams), 4 GiB each."""
import time

import torch

n = 1 << 29
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True); h_out.fill_(0.0)
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 8 / 1e9


def run(h2d, d2h, chunks=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = n // chunks
    for c in range(chunks):
        sl = slice(c * step, (c + 1) * step)
        if h2d:
            with torch.cuda.stream(s1):
                d_in[sl].copy_(h_in[sl], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for name, a, b, ch in (("H2D alone", 1, 0, 1), ("D2H alone", 0, 1, 1), ("both", 1, 1, 1), ("both, 16 chunks", 1, 1, 16), ("both, 128 chunks", 1, 1, 128)):
    run(a, b, ch)
    t = min(run(a, b, ch) for _ in range(3))
    print(f"{name:18s} {t * 1e3:7.1f} ms  {gb / t:6.1f} GB/s per direction")
