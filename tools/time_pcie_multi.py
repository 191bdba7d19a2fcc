"""PCIe / host-memory ceiling with every GPU of the box copying at once (the e2e path at N > 1):
torchrun --nproc-per-node N tools/time_pcie_multi.py -- pinned 2 GiB per direction per rank, both directions busy."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
numa = bind_to_gpu_numa_node(local)
dist.init_process_group("gloo")
n = 1 << 28
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True); h_out.fill_(0.0)
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 8 / 1e9


def run(h2d, d2h, chunks=16, active=True):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    step = n // chunks
    if active:
        for c in range(chunks):
            sl = slice(c * step, (c + 1) * step)
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    dist.barrier()
    return t


for name, a, b, who in (("H2D, all ranks", 1, 0, None), ("both, all ranks", 1, 1, None), ("both, rank 0 alone", 1, 1, 0),
                        ("both, ranks 0+1", 1, 1, (0, 1)), ("both, even ranks", 1, 1, tuple(range(0, world, 2)))):
    act = who is None or (rank == who if isinstance(who, int) else rank in who)
    run(a, b, active=act)
    t = min(run(a, b, active=act) for _ in range(3))
    ts = [None] * world
    dist.all_gather_object(ts, (t if act else None, numa))
    if rank == 0:
        rates = [f"{gb / x[0]:5.1f}" if x[0] else "    -" for x in ts]
        agg = sum(gb / x[0] for x in ts if x[0])
        print(f"{name:20s} per-rank GB/s per direction: {' '.join(rates)}   aggregate {agg:6.1f} GB/s per direction", flush=True)
if rank == 0:
    print("numa:", [x[1] for x in ts])
dist.destroy_process_group()
