"""Multi-GPU parity of the slab Laplacian (run under torchrun): every rank's slab -- device path (halo exchange
through NVLink peer memory, or NCCL) and host path (SlabLaplacian.apply_host: chunked H2D / per-chunk halo
exchange / stencil / D2H) -- must equal the same slab of a single-GPU run of the global grid (bit-exact)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
ok_all = True
for gshape in ((96, 80, 64 * world), (130, 64, 192 * world), (33, 47, 70 * world)):
    hs = [float(np.linspace(0, 1, n)[1]) for n in gshape]
    torch.manual_seed(7)
    fg = torch.randn(*gshape, dtype=torch.float64, device="cuda")          # same on every rank (same seed)
    ref = SlabLaplacian(gshape, hs).apply_local(fg, torch.empty_like(fg))
    s = SlabLaplacian(gshape, hs, rank=rank, nranks=world)
    local = fg[..., s.z0:s.z1].contiguous()
    ok = True
    for step in range(3):                          # several steps: the two halo buffer sets alternate
        out = s(local)
        torch.cuda.synchronize()
        ok = ok and torch.equal(out, ref[..., s.z0:s.z1])
    # host buffers in / out: chunked pipeline with per-chunk halo exchange (SlabLaplacian.apply_host)
    h_in = torch.empty(local.shape, dtype=torch.float64, pin_memory=True).copy_(local)
    h_out = torch.empty(local.shape, dtype=torch.float64, pin_memory=True)
    for step in range(2):
        h_out.fill_(float("nan"))
        s.apply_host(h_in, h_out, chunk_mib=1)     # small chunks: several per slab even at these sizes
        ok_host = torch.equal(h_out.cuda(), ref[..., s.z0:s.z1])
        if not ok_host:
            print(f"rank {rank}: apply_host mismatch on {gshape} (step {step})")
        ok = ok and ok_host
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        path = "cabi" if getattr(s, "_cabi_plan", None) is not None else ("symm" if getattr(s, "_symm", None) is not None else "nccl")
        print(f"global {gshape} on {world} ranks: bit-exact={bool(t.item())} path={path} {getattr(s, '_cabi_error', '')}")
    ok_all &= bool(t.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
