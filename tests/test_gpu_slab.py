"""Slab layer on the GPU: every rank of an N-way last-axis decomposition is run inside one process
(slab.EmulatedWorld: same pack / block-copy / apply kernels in the same order as on N GPUs, peer
buffers are local tensors, signals are thread barriers) and each slab must equal the same slab of
the single-GPU result BIT FOR BIT -- for finite-difference, log, Chebyshev and FFT axes, sharded
or not, and for the curvilinear composites built on them.  tools/check_slab_nccl.py runs the same
comparison across real GPUs (torchrun)."""
import numpy as np
import pytest

import numgrids_b200 as ng
from numgrids_b200 import slab
from tests_support import build_curv_grid, smooth_field

pytestmark = pytest.mark.gpu
A = ng.AxisType


def _slabs(world, fn):
    return world.run(fn)


def _check(parts, ref, nranks, what, exact=True):
    """exact: bit for bit.  Otherwise (an FFT axis is involved) rel. L-inf <= 1e-13: the FFT kernels
    transform two real pencils as one complex one, so a pencil's rounding depends on which
    neighbour it is paired with, and the pairing follows the local layout (tolerance of the FFT
    path itself: 1e-11, BASELINE.json)."""
    z = 0
    scale = float(np.max(np.abs(ref))) or 1.0
    for r, p in enumerate(parts):
        zl = p.shape[-1]
        got = p.cpu().numpy()
        want = ref[..., z:z + zl]
        if exact:
            assert np.array_equal(got, want), (what, "rank", r, "of", nranks)
        else:
            assert float(np.max(np.abs(got - want))) / scale <= 1e-13, (what, "rank", r, "of", nranks)
        z += zl
    assert z == ref.shape[-1]


GRIDS = {
    "fd": [(A.EQUIDISTANT, 14, 0.0, 1.0), (A.EQUIDISTANT, 9, 0.0, 2.0), (A.EQUIDISTANT, 26, -1.0, 1.0)],
    "cheb": [(A.CHEBYSHEV, 12, 0.0, 1.0), (A.EQUIDISTANT, 7, 0.0, 1.0), (A.CHEBYSHEV, 23, 0.5, 2.0)],
    "fft": [(A.EQUIDISTANT, 10, 0.0, 1.0), (A.CHEBYSHEV, 6, 0.0, 1.0), (A.EQUIDISTANT_PERIODIC, 64, 0.0, 2 * np.pi)],
    "fft_odd": [(A.CHEBYSHEV, 9, 0.0, 1.0), (A.EQUIDISTANT_PERIODIC, 30, 0.0, 2 * np.pi)],
    "log": [(A.EQUIDISTANT, 8, 0.0, 1.0), (A.LOGARITHMIC, 25, 0.1, 10.0)],
}


@pytest.mark.parametrize("name", sorted(GRIDS))
@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_diff_every_axis_bit_exact(gpu_ready, name, nranks):
    import torch
    axes = [ng.create_axis(*spec) for spec in GRIDS[name]]
    g = ng.Grid(*axes)
    f = smooth_field([a.coords for a in axes], 0.3)
    fd = torch.from_numpy(f).cuda()
    for order in (1, 2):
        for axis in range(g.ndims):
            ref = ng.Diff(g, order, axis)(fd).cpu().numpy()
            world = slab.EmulatedWorld(nranks)

            def rank_fn(r):
                d = slab.SlabDiff(g, order, axis, 4, r, nranks, world)
                local = fd[..., d.z0:d.z1].contiguous()
                out = d(local)
                return d(local, out) if axis == g.ndims - 1 else out     # second call: buffer reuse
            _check(_slabs(world, rank_fn), ref, nranks, (name, order, axis), exact=not axes[axis].periodic)


CURV = {
    "spherical": [(A.CHEBYSHEV, 10, 0.5, 2.0), (A.CHEBYSHEV, 9, 0.3, np.pi - 0.3),
                  (A.EQUIDISTANT_PERIODIC, 32, 0.0, 2 * np.pi)],
    "cylindrical": [(A.CHEBYSHEV, 9, 0.0, 1.5), (A.EQUIDISTANT_PERIODIC, 16, 0.0, 2 * np.pi),
                    (A.EQUIDISTANT, 22, -1.0, 1.0)],
    "custom3": [(A.EQUIDISTANT, 12, 0.1, 1.0), (A.CHEBYSHEV, 8, 0.0, 1.0), (A.CHEBYSHEV, 17, 0.0, 1.0)],
    "unit3": [(A.EQUIDISTANT, 11, 0.0, 1.0), (A.EQUIDISTANT, 10, 0.0, 1.0), (A.EQUIDISTANT, 24, 0.0, 1.0)],
    "polar": [(A.CHEBYSHEV, 12, 0.0, 1.0), (A.EQUIDISTANT_PERIODIC, 32, 0.0, 2 * np.pi)],
}


@pytest.mark.parametrize("name", sorted(CURV))
@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_curvilinear_bit_exact(gpu_ready, name, nranks):
    """laplacian / gradient / divergence / curl on slabs == the single-GPU CurvilinearGrid methods
    (incl. the r = 0 singular line of the cylindrical / polar grids; unit3 is cfg5's idiom (ii))."""
    import torch
    axes = [ng.create_axis(*spec) for spec in CURV[name]]
    g = build_curv_grid(ng, name, axes)
    coords = [a.coords for a in axes]
    fs = [torch.from_numpy(smooth_field(coords, 0.1 * (i + 1))).cuda() for i in range(g.ndims)]
    ref_lap = g.laplacian(fs[0]).cpu().numpy()
    ref_grad = [x.cpu().numpy() for x in g.gradient(fs[0])]
    ref_div = g.divergence(*fs).cpu().numpy()
    rc = g.curl(*fs)
    ref_curl = [x.cpu().numpy() for x in (rc if isinstance(rc, tuple) else (rc,))]
    world = slab.EmulatedWorld(nranks)

    def rank_fn(r):
        sc = slab.SlabCurvilinear(g, r, nranks, world)
        loc = [x[..., sc.z0:sc.z1].contiguous() for x in fs]
        lap = sc.laplacian(loc[0])
        lap2 = sc.laplacian(loc[0])
        assert torch.equal(lap, lap2)
        grad = sc.gradient(loc[0])
        div = sc.divergence(*loc)
        curl = sc.curl(*loc)
        return lap, grad, div, curl if isinstance(curl, tuple) else (curl,)
    res = _slabs(world, rank_fn)
    fft_axes = [i for i, a in enumerate(axes) if a.periodic]
    exact = not fft_axes
    _check([x[0] for x in res], ref_lap, nranks, (name, "laplacian"), exact)
    for i in range(g.ndims):
        _check([x[1][i] for x in res], ref_grad[i], nranks, (name, "gradient", i), i not in fft_axes)
    _check([x[2] for x in res], ref_div, nranks, (name, "divergence"), exact)
    for i in range(len(ref_curl)):
        _check([x[3][i] for x in res], ref_curl[i], nranks, (name, "curl", i), exact)


def test_block_copy_argument_checks(gpu_ready):
    import ctypes as C
    import torch
    from numgrids_b200 import _lib
    lib = _lib.load()
    a = torch.arange(24, dtype=torch.float64, device="cuda")
    b = torch.zeros(24, dtype=torch.float64, device="cuda")
    rd = slab.Redistribution((4, 3, 6), 0, 2)
    blk = rd._blk(a.data_ptr() + 8, b.data_ptr() + 8, 3, 5, 7, 7, 0)       # odd, misaligned: scalar path
    slab.run_blocks([blk])
    exp = torch.zeros(24, dtype=torch.float64)
    for r in range(3):
        exp[1 + 7 * r:6 + 7 * r] = torch.arange(1 + 7 * r, 6 + 7 * r, dtype=torch.float64)
    assert torch.equal(b.cpu(), exp)
    bad = rd._blk(a.data_ptr(), b.data_ptr(), 2, 8, 4, 8, 0)                # rows overlap
    arr = (_lib.NgCopyBlock * 1)(bad)
    assert lib.ng_slab_block_copy(1, arr, 0, None, None, 0, None) != 0
    assert lib.ng_slab_block_copy(17, arr, 0, None, None, 0, None) != 0
    assert lib.ng_slab_block_copy(1, arr, 0, None, None, 1, None) != 0      # fused without fuse record


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_step_behind_the_c_abi(gpu_ready, nranks):
    """ng_slab_plan_create / ng_slab_lap_apply (include/numgrids_b200.h): the whole multi-GPU step -- pack into the
    neighbours' halo buffers, flags, interior || boundary tiles, join -- with every rank's plan living in THIS process:
    the library takes plain device pointers, so the "peer" buffers are local tensors and the flags travel between the
    ranks' side streams of the one GPU.  Several steps (the halo sets alternate); every slab bit for bit."""
    import ctypes as C
    import torch
    from numgrids_b200 import _lib
    from numgrids_b200.slab import SlabLaplacian, make_fd_plan
    lib = _lib.load()
    gshape = (40, 33, 48 * nranks + 5)
    hs = [float(np.linspace(0, 1, n)[1]) for n in gshape]
    torch.manual_seed(11)
    fg = torch.randn(*gshape, dtype=torch.float64, device="cuda")
    ref = SlabLaplacian(gshape, hs).apply_local(fg, torch.empty_like(fg))
    ranks = [SlabLaplacian(gshape, hs, rank=r, nranks=nranks) for r in range(nranks)]
    rows, nb = ranks[0].rows, ranks[0].nb
    halo = [torch.zeros(2, 2, rows, nb, dtype=torch.float64, device="cuda") for _ in range(nranks)]
    flags = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(nranks)]
    plans, keep = [], []
    for r, s in enumerate(ranks):
        c = _lib.NgSlabComm()
        c.rank, c.nranks = r, nranks
        for q in range(2):
            for side in range(2):
                c.my_halo[q][side] = halo[r][q, side].data_ptr()
            c.peer_lo_halo[q] = halo[r - 1][q, 1].data_ptr() if r > 0 else None
            c.peer_hi_halo[q] = halo[r + 1][q, 0].data_ptr() if r < nranks - 1 else None
        c.my_flags = flags[r].data_ptr()
        c.peer_lo_flag = flags[r - 1].data_ptr() + 8 if r > 0 else None
        c.peer_hi_flag = flags[r + 1].data_ptr() if r < nranks - 1 else None
        children = [make_fd_plan(s.local_shape, a, 2, 4, hs[a]) for a in range(3)]
        arr = _lib.ptr_array([p.handle.value for p in children])
        shp, shp_p = _lib.as_i64(list(s.local_shape))
        out = C.c_void_p()
        _lib.check(lib.ng_slab_plan_create(C.byref(c), 3, shp_p, arr, C.byref(out)), "ng_slab_plan_create")
        plans.append(_lib.Plan(out.value, keepalive=children))
        keep.append(c)
    streams = [torch.cuda.Stream() for _ in range(nranks)]
    local = [fg[..., s.z0:s.z1].contiguous() for s in ranks]
    outs = [torch.empty_like(x) for x in local]
    # every kernel instance a step launches is loaded before the first flag wait spins: CUDA loads kernels lazily, and
    # loading one synchronises the context -- across the ranks' streams when all ranks share ONE context as they do here
    for r, s in enumerate(ranks):
        b = s.pack(local[r])
        s.apply_part(local[r], outs[r], 1)
        s.apply_part(local[r], outs[r], 2, b["lo_halo"] if r > 0 else None, b["hi_halo"] if r < nranks - 1 else None)
    torch.cuda.synchronize()
    for step in range(4):
        for o in outs:
            o.fill_(float("nan"))
        torch.cuda.synchronize()
        for r in range(nranks):
            _lib.check(lib.ng_slab_lap_apply(plans[r].handle, local[r].data_ptr(), outs[r].data_ptr(),
                                             streams[r].cuda_stream), "ng_slab_lap_apply")
        torch.cuda.synchronize()
        for r, s in enumerate(ranks):
            assert lib.ng_slab_timed_out(plans[r].handle) == 0
            assert torch.equal(outs[r], ref[..., s.z0:s.z1]), ("step", step, "rank", r)
    # argument checks: a rank with a neighbour but no buffers is refused
    bad = _lib.NgSlabComm()
    bad.rank, bad.nranks = 0, 2
    out = C.c_void_p()
    shp, shp_p = _lib.as_i64(list(ranks[0].local_shape))
    children = [make_fd_plan(ranks[0].local_shape, a, 2, 4, hs[a]) for a in range(3)]
    arr = _lib.ptr_array([p.handle.value for p in children])
    assert lib.ng_slab_plan_create(C.byref(bad), 3, shp_p, arr, C.byref(out)) != 0
