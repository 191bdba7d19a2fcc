"""Multi-GPU parity of the slab Laplacian (run under torchrun): every rank's slab, computed with the
NCCL halo exchange, must equal the same slab of a single-GPU run of the global grid (bit-exact)."""
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
    out = s(local)
    torch.cuda.synchronize()
    ok = torch.equal(out, ref[..., s.z0:s.z1])
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"global {gshape} on {world} ranks: bit-exact={bool(t.item())}")
    ok_all &= bool(t.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
