"""Multi-GPU parity of the general slab layer (run under torchrun): SlabDiff on finite-difference,
Chebyshev and FFT axes (sharded axis included: halo exchange / pencil redistribution over NVLink peer
memory, or NCCL with NG_SLAB_EXCHANGE=nccl) and SlabCurvilinear's composites must equal the same
slab of a single-GPU run of the global grid bit for bit.  Also times the redistribution path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import slab  # noqa: E402

A = ng.AxisType
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
ok_all = True


def agree(ok, what):
    global ok_all
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{what}: match={bool(t.item())}", flush=True)
    ok_all &= bool(t.item())


def same(a, b, exact=True):
    """Bit for bit, or (FFT axis involved: pencil pairing follows the local layout) rel. L-inf <= 1e-13."""
    if exact:
        return torch.equal(a, b)
    return float((a - b).abs().max() / b.abs().max()) <= 1e-13


def field(shape, seed):
    torch.manual_seed(seed)                      # same on every rank
    return torch.randn(*shape, dtype=torch.float64, device="cuda")


grids = {
    "fd": [(A.EQUIDISTANT, 40, 0.0, 1.0), (A.EQUIDISTANT, 33, 0.0, 2.0), (A.EQUIDISTANT, 48 * world, -1.0, 1.0)],
    "cheb": [(A.CHEBYSHEV, 32, 0.0, 1.0), (A.EQUIDISTANT, 17, 0.0, 1.0), (A.CHEBYSHEV, 24 * world + 1, 0.5, 2.0)],
    "fft": [(A.EQUIDISTANT, 20, 0.0, 1.0), (A.CHEBYSHEV, 16, 0.0, 1.0), (A.EQUIDISTANT_PERIODIC, 256, 0.0, 2 * np.pi)],
}
for name, specs in grids.items():
    g = ng.Grid(*[ng.create_axis(*s) for s in specs])
    fg = field(g.shape, 11)
    for order in (1, 2):
        for axis in range(g.ndims):
            ref = ng.Diff(g, order, axis)(fg)
            d = slab.SlabDiff(g, order, axis, 4, rank, world)
            loc = fg[..., d.z0:d.z1].contiguous()
            out = d(loc)
            out = d(loc, out)
            torch.cuda.synchronize()
            agree(same(out, ref[..., d.z0:d.z1], not g.axes[axis].periodic), f"{name} order {order} axis {axis} x{world}")
            if getattr(d, "_symm_failed", False) and rank == 0:
                print("  symmetric memory unavailable:", d._symm_error)

sph = ng.SphericalGrid(ng.create_axis(A.CHEBYSHEV, 32, 0.5, 2.0), ng.create_axis(A.CHEBYSHEV, 24, 0.3, np.pi - 0.3),
                       ng.create_axis(A.EQUIDISTANT_PERIODIC, 128, 0.0, 2 * np.pi))
unit = ng.CurvilinearGrid(*[ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0) for n in (24, 20, 32 * world)],
                          scale_factors=(lambda c: np.ones_like(c[0]),) * 3)
for name, g in (("spherical", sph), ("unit3", unit)):
    fs = [field(g.shape, 20 + i) for i in range(3)]
    sc = slab.SlabCurvilinear(g, rank, world)
    loc = [x[..., sc.z0:sc.z1].contiguous() for x in fs]
    sl = (Ellipsis, slice(sc.z0, sc.z1))
    ex = name != "spherical"
    agree(same(sc.laplacian(loc[0]), g.laplacian(fs[0])[sl], ex), f"{name} laplacian x{world}")
    agree(all(same(a, b[sl], ex or i < 2) for i, (a, b) in enumerate(zip(sc.gradient(loc[0]), g.gradient(fs[0])))),
          f"{name} gradient x{world}")
    agree(same(sc.divergence(*loc), g.divergence(*fs)[sl], ex), f"{name} divergence x{world}")
    agree(all(same(a, b[sl], ex) for a, b in zip(sc.curl(*loc), g.curl(*fs))), f"{name} curl x{world}")

# timing: FFT along the sharded axis (cfg3-like, weak: 256 x 512 x 1024 points per GPU)
if os.environ.get("NG_CHECK_TIME", "1") == "1":
    e0 = ng.create_axis(A.EQUIDISTANT, 256 * world, 0.0, 1.0)
    e1 = ng.create_axis(A.EQUIDISTANT, 512, 0.0, 1.0)
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, 1024, 0.0, 2 * np.pi)
    g = ng.Grid(e0, e1, p)
    d = slab.SlabDiff(g, 1, 2, 4, rank, world)
    loc = torch.randn(*d.local_shape, dtype=torch.float64, device="cuda")
    out = torch.empty_like(loc)
    for mode in ("symm", "nccl"):
        os.environ["NG_SLAB_EXCHANGE"] = mode
        for _ in range(3):
            d(loc, out)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            d(loc, out)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / 10], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            pts = float(np.prod(g.shape))
            print(f"FFT along the sharded axis, global {g.shape} on {world} GPUs, exchange={mode}: "
                  f"{ms.item():.3f} ms/apply, {pts / ms.item() / 1e6:.1f} Gpts/s", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
