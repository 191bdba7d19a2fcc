"""GPU parity at (or near) BASELINE.json config sizes against the pencil oracle, plus
size-independent properties at full size.  Inputs are the analytic fields of SURVEY.md 8(d)."""
import numpy as np
import pytest

import numgrids_b200 as ng
from conftest import TOL_FFT, rel_linf
from oracle import pencil

pytestmark = pytest.mark.gpu
A = ng.AxisType


def test_cfg1_fd_512x512_bit_exact(gpu_ready):
    ax = ng.create_axis(A.EQUIDISTANT, 512, 0.0, 1.0)
    g = ng.Grid(ax, ax)
    X, Y = g.meshed_coords
    f = np.sin(2 * np.pi * X) * np.cos(3 * Y) + 0.3 * np.exp(-X * Y)
    for order, axis in ((1, 0), (1, 1), (2, 0), (2, 1)):
        out = ng.Diff(g, order, axis)(f)
        ref = pencil.fd_apply(f, axis, ax.spacing, order, 4)
        assert np.array_equal(out, ref), (order, axis, rel_linf(out, ref))


@pytest.mark.parametrize("n", [48, 128])
def test_cfg2_chebyshev_cube_bit_exact(gpu_ready, n):
    """Diff(g,2,0) on Chebyshev^3 (descending-j order) and the other axes/orders; 256^3 runs in bench."""
    ax = ng.create_axis(A.CHEBYSHEV, n, 0.0, 1.0)
    g = ng.Grid(ax, ax, ax)
    x = ax.coords
    f = (np.sin(3 * x)[:, None, None] * np.cos(2 * x)[None, :, None] * np.exp(-x)[None, None, :])
    pax = [pencil.make_axis("chebyshev", n, 0.0, 1.0)] * 3
    for order, axis in ((2, 0), (2, 1), (2, 2), (1, 0), (1, 2)):
        out = ng.Diff(g, order, axis)(f)
        ref = pencil.make_diff(pax, order, axis)(f)
        assert np.array_equal(out, ref), (n, order, axis, rel_linf(out, ref))


def test_cfg2_chebyshev_256_cube_bit_exact_bench_instance(gpu_ready):
    """BASELINE cfg2 at its stated size: Diff(g, 2, 0) on Chebyshev 256^3 and the other axes / order 1, bit for
    bit against the oracle.  At this size the dispatcher picks the 4-column-pair tile variant (the instance
    bench.py times), which the 48^3 / 128^3 cases never reach -- asserted through ng_last_kernel()."""
    from numgrids_b200 import _lib
    from oracle import pencil_c
    n = 256
    ax = ng.create_axis(A.CHEBYSHEV, n, 0.0, 1.0)
    g = ng.Grid(ax, ax, ax)
    x = ax.coords
    f = (np.sin(3 * x)[:, None, None] * np.cos(2 * x)[None, :, None] * np.exp(-x)[None, None, :])
    pax = [pencil.make_axis("chebyshev", n, 0.0, 1.0)] * 3
    seen = set()
    for order, axis in ((2, 0), (2, 1), (2, 2), (1, 0), (1, 1), (1, 2)):
        out = ng.Diff(g, order, axis)(f)
        seen.add(_lib.last_kernel())
        M = pencil.chebyshev_matrix(n, 0.0, 1.0, order, last_axis_of_nd=(axis == 2))
        ref = pencil_c.dense_apply(f, axis, M, pencil.chebyshev_descending(order, axis, 3))   # C oracle (pinned to the
        assert np.array_equal(out, ref), (order, axis, rel_linf(out, ref))                    # NumPy one, test_oracle_c)
        if (order, axis) == (2, 0):          # the BASELINE case once more against the NumPy oracle itself
            assert np.array_equal(out, pencil.make_diff(pax, 2, 0)(f))
    assert seen == {"dense_fast<ncp4>", "dense_fast<ncp4,lastax>"}, seen


@pytest.mark.parametrize("n", [1024, 512, 256, 128])
def test_cfg3_fft_ring_kernel_vs_oracle(gpu_ready, n):
    """The persistent ring kernel -- the instance bench.py times at cfg3 -- needs >= 1776 pencils: 64 x 32 x n
    has 2048.  Orders 1 and 2 against numpy.fft (oracle), tolerance 1e-11 (order 2 at n = 1024: 3e-11, see
    test_cfg3_fft_axis_1024)."""
    from numgrids_b200 import _lib
    e0 = ng.create_axis(A.EQUIDISTANT, 64, 0.0, 1.0)
    e1 = ng.create_axis(A.EQUIDISTANT, 32, 0.0, 1.0)
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, n, 0.0, 2 * np.pi)
    g = ng.Grid(e0, e1, p)
    X, Y, P = g.meshed_coords
    f = np.exp(np.sin(P)) * np.cos(2 * X) + Y * np.sin(3 * P)
    import torch
    fd = torch.from_numpy(f).cuda()
    for order in (1, 2):
        out = ng.Diff(g, order, 2)(fd)
        assert _lib.last_kernel() == "fft2pass_ring", _lib.last_kernel()
        ref = pencil.fft_apply(f, 2, pencil.fft_multiplier(n, p.spacing, order))
        err = rel_linf(out.cpu().numpy(), ref)
        print(f"ring kernel n={n} order {order}: rel Linf {err:.3e}")
        assert err <= (3e-11 if (order == 2 and n == 1024) else TOL_FFT), (n, order, err)


@pytest.mark.parametrize("n", [1024, 512, 256, 128, 64, 100, 360])
@pytest.mark.parametrize("shape_kind", ["middle", "first", "ragged"])
def test_fft_strided_axis_tile_kernel_vs_oracle(gpu_ready, n, shape_kind):
    """Periodic axis that is not the contiguous one (cylindrical phi in the middle, doubly periodic boxes): the TMA
    tile kernel (csrc/fft_tile.cu), orders 1 and 2 against numpy.fft (oracle), 1e-11.  `ragged`: an inner extent
    that is no multiple of the tile width (the last tile hangs over the array: zero-filled loads, clipped stores)
    and more tiles than slots; one pencil carries an inf (the reference's singular rows): it must come out all-NaN
    and leave its pair partner and every other pencil untouched.  n = 100, 360: zero-padded 256 / 1024-point transforms
    (the rows past n are the TMA's out-of-bounds zero fill)."""
    import torch
    from numgrids_b200 import _lib
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, n, 0.0, 2 * np.pi)
    want_kernel = "fft_tile" if n & (n - 1) == 0 else "fft_tile<padded>"
    if shape_kind == "middle":
        axes, ax = [ng.create_axis(A.EQUIDISTANT, 5, 0.0, 1.0), p, ng.create_axis(A.CHEBYSHEV, 48, 0.0, 1.0)], 1
    elif shape_kind == "first":
        axes, ax = [p, ng.create_axis(A.EQUIDISTANT, 160, 0.0, 1.0)], 0
    else:
        axes, ax = [ng.create_axis(A.EQUIDISTANT, 3, 0.0, 1.0), p, ng.create_axis(A.EQUIDISTANT, 7, 0.0, 1.0),
                    ng.create_axis(A.EQUIDISTANT, 6, 0.0, 1.0)], 1           # inner = 42
    g = ng.Grid(*axes)
    M = g.meshed_coords
    P = M[ax]
    others = [m for i, m in enumerate(M) if i != ax]
    f = np.exp(np.sin(P)) * np.cos(2 * others[0]) + others[-1] * np.sin(3 * P)
    if shape_kind == "ragged":
        f[1, 5, 2, 3] = np.inf
    fd = torch.from_numpy(f).cuda()
    for order in (1, 2):
        out = ng.Diff(g, order, ax)(fd)
        assert _lib.last_kernel() == want_kernel, _lib.last_kernel()
        got = out.cpu().numpy()
        with np.errstate(invalid="ignore"):
            ref = pencil.fft_apply(f, ax, pencil.fft_multiplier(n, p.spacing, order))
        if shape_kind == "ragged":
            assert np.all(np.isnan(got[1, :, 2, 3])) and np.all(np.isnan(ref[1, :, 2, 3]))
            got[1, :, 2, 3] = ref[1, :, 2, 3] = 0.0
        assert np.all(np.isfinite(got))
        err = rel_linf(got, ref)
        print(f"tile kernel n={n} {shape_kind} order {order}: rel Linf {err:.3e}")
        assert err <= (3e-11 if (order == 2 and n == 1024) else TOL_FFT), (n, shape_kind, order, err)


def test_cfg3_full_size_512x512x1024(gpu_ready):
    """BASELINE cfg3 at its stated size (2 GiB array, 262 144 pencils): pencils are independent, so three
    8-plane blocks of the full-size result are compared with numpy.fft on the same blocks."""
    import torch
    from numgrids_b200 import _lib
    e = ng.create_axis(A.EQUIDISTANT, 512, 0.0, 1.0)
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, 1024, 0.0, 2 * np.pi)
    g = ng.Grid(e, e, p)
    x = torch.from_numpy(e.coords).cuda()
    ph = torch.from_numpy(p.coords).cuda()
    f = (torch.exp(torch.sin(ph))[None, None, :] * torch.cos(2 * x)[:, None, None]
         + x[None, :, None] * torch.sin(3 * ph)[None, None, :]).contiguous()
    out = ng.Diff(g, 1, 2)(f)
    assert _lib.last_kernel() == "fft2pass_ring"
    W = pencil.fft_multiplier(1024, p.spacing, 1)
    worst = 0.0
    for a in (0, 252, 504):
        blk = f[a:a + 8].cpu().numpy()
        ref = pencil.fft_apply(blk, 2, W)
        worst = max(worst, rel_linf(out[a:a + 8].cpu().numpy(), ref))
    print(f"cfg3 full size: rel Linf {worst:.3e}")
    assert worst <= TOL_FFT


def test_cfg3_fft_axis_1024(gpu_ready):
    """Periodic 1024-point last axis, reduced batch (32x32), Nyquist zeroed; tolerance 1e-11."""
    e = ng.create_axis(A.EQUIDISTANT, 32, 0.0, 1.0)
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, 1024, 0.0, 2 * np.pi)
    g = ng.Grid(e, e, p)
    X, Y, P = g.meshed_coords
    f = np.exp(np.sin(P)) * np.cos(2 * X) + Y * np.sin(3 * P)
    for order in (1, 2):
        out = ng.Diff(g, order, 2)(f)
        W = pencil.fft_multiplier(1024, p.spacing, order)
        ref = pencil.fft_apply(f, 2, W)
        err = rel_linf(out, ref)
        print(f"cfg3 order {order}: rel Linf {err:.3e}")
        # Order 1 (the BASELINE config) must meet the 1e-11 contract.  Order 2 at n=1024 amplifies
        # rounding noise by (n/2)^2: numpy.fft itself is 1.15e-11 away from a long-double DFT and
        # 1.34e-11 away from scipy.fft on this field ("FFT tolerance"), so no independent
        # FFT can be closer to it than ~1.5e-11; the bound here is 3e-11.
        assert err <= (TOL_FFT if order == 1 else 3e-11)
    # analytic check (reference tests/test_diff.py:44-60 style)
    np.testing.assert_allclose(ng.Diff(g, 1, 2)(np.exp(np.sin(P))), np.cos(P) * np.exp(np.sin(P)), atol=1e-10)


def test_fft_non_last_axis_and_odd_sizes(gpu_ready):
    for shape_axis in (((64, 12), 0), ((6, 128, 10), 1), ((21, 9), 0), ((7, 30), 1), ((100,), 0)):
        shape, axis = shape_axis
        axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC, n, 0.0, 2 * np.pi) for n in shape]
        g = ng.Grid(*axes)
        rng = np.random.default_rng(1)
        f = rng.standard_normal(shape)
        for order in (1, 2, 3):
            out = ng.Diff(g, order, axis)(f)
            ref = pencil.fft_apply(f, axis, pencil.fft_multiplier(shape[axis], axes[axis].spacing, order))
            assert rel_linf(out, ref) <= TOL_FFT, (shape, axis, order, rel_linf(out, ref))


def test_cfg4_spherical_full_size(gpu_ready):
    """SphericalGrid(Cheb 128, Cheb 128, periodic 256): laplacian + gradient + divergence (+ curl)."""
    r = ng.create_axis(A.CHEBYSHEV, 128, 0.5, 2.0)
    th = ng.create_axis(A.CHEBYSHEV, 128, 0.3, np.pi - 0.3)
    ph = ng.create_axis(A.EQUIDISTANT_PERIODIC, 256, 0.0, 2 * np.pi)
    g = ng.SphericalGrid(r, th, ph)
    R, T, P = g.meshed_coords
    f = R ** 2 * np.sin(T) ** 2 * np.cos(2 * P) + np.exp(-R) * np.cos(T)
    pax = [pencil.make_axis("chebyshev", 128, 0.5, 2.0), pencil.make_axis("chebyshev", 128, 0.3, np.pi - 0.3),
           pencil.make_axis("periodic", 256, 0.0, 2 * np.pi)]
    D = [pencil.make_diff(pax, 1, i) for i in range(3)]
    h = (np.ones_like(R), R, R * np.sin(T))
    lap = g.laplacian(f)
    ref = pencil.laplacian(f, h, D)
    print(f"cfg4 laplacian rel Linf {rel_linf(lap, ref):.3e}")
    assert rel_linf(lap, ref) <= 1e-12
    grad = g.gradient(f)
    gref = pencil.gradient(f, h, D)
    assert np.array_equal(grad[0], gref[0]) and np.array_equal(grad[1], gref[1])      # Chebyshev axes: bit-exact
    assert rel_linf(grad[2], gref[2]) <= TOL_FFT
    v = (R, np.sin(T), np.cos(P))
    div = g.divergence(*v)
    assert rel_linf(div, pencil.divergence(v, h, D)) <= 1e-12
    curl = g.curl(*v)
    cref = pencil.curl(v, h, D)
    for i in range(3):
        assert rel_linf(curl[i], cref[i]) <= TOL_FFT


def test_cfg5_cartesian_laplacian_bit_exact_and_slabs(gpu_ready):
    """Fused Laplacian vs the oracle idiom at 96x80x128; then the last axis split into 2 and 4
    slabs with packed halos (one GPU emulating all ranks): every slab equals the global result."""
    import torch
    from numgrids_b200.slab import SlabLaplacian
    shape = (96, 80, 128)
    axes = [ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0 + i) for i, n in enumerate(shape)]
    g = ng.Grid(*axes)
    x, y, zc = (a.coords for a in axes)
    f = (np.sin(2 * np.pi * x)[:, None, None] * np.cos(2 * np.pi * y)[None, :, None] * np.exp(zc)[None, None, :])
    ref = pencil.cartesian_laplacian(f, [a.spacing for a in axes], 4)
    out = ng.cartesian_laplacian(g, f)
    assert np.array_equal(out, ref), rel_linf(out, ref)
    for nranks in (2, 4):
        parts = SlabLaplacian.emulate_on_one_gpu(g, torch.from_numpy(f).cuda(), nranks)
        got = torch.cat(parts, dim=2).cpu().numpy()
        assert np.array_equal(got, ref), nranks


def test_linearity_and_constants_full_size(gpu_ready):
    """Size-independent properties at 256^3-class sizes on the device path: D(const)=0 for FD,
    D(a f + b g) ~ a D f + b D g, fused Laplacian == three-call idiom."""
    import torch
    n = 256
    ax = ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0)
    g = ng.Grid(ax, ax, ax)
    torch.manual_seed(0)
    f = torch.randn(n, n, n, dtype=torch.float64, device="cuda")
    h = torch.randn(n, n, n, dtype=torch.float64, device="cuda")
    lap = ng.CartesianLaplacian(g)
    idiom = ng.Diff(g, 2, 0)(f) + ng.Diff(g, 2, 1)(f) + ng.Diff(g, 2, 2)(f)
    assert torch.equal(lap(f), idiom)
    d = ng.Diff(g, 1, 1)
    lhs = d(2.0 * f + 0.5 * h)
    rhs = 2.0 * d(f) + 0.5 * d(h)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-12
    const = torch.full((n, n, n), 3.25, dtype=torch.float64, device="cuda")
    assert float(d(const).abs().max()) < 1e-9


def test_cfg5_full_size_identities(gpu_ready, monkeypatch):
    """BASELINE size (1024^3 points): the fused Laplacian equals the reference idiom
    Diff(g,2,0)(f)+Diff(g,2,1)(f)+Diff(g,2,2)(f) bit for bit, and every per-axis streaming kernel
    equals the gather kernel (itself pinned to the oracle at small sizes) bit for bit."""
    import torch
    from numgrids_b200 import _lib
    n = 1024
    ax = ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0)
    g = ng.Grid(ax, ax, ax)
    x = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device="cuda")
    f = (torch.sin(2 * np.pi * x)[:, None, None] * torch.cos(2 * np.pi * x)[None, :, None]
         * torch.exp(x)[None, None, :]).contiguous()
    idiom = ng.Diff(g, 2, 0)(f)
    idiom += ng.Diff(g, 2, 1)(f)
    idiom += ng.Diff(g, 2, 2)(f)
    lap = ng.CartesianLaplacian(g)(f)
    assert torch.equal(lap, idiom)
    del lap, idiom
    for order, axis in ((1, 0), (2, 1), (1, 2), (2, 2)):
        op = ng.Diff(g, order, axis)
        monkeypatch.setenv("NG_FD_STREAM", "1")
        _lib.tuning_reload()
        fast = op(f)
        monkeypatch.setenv("NG_FD_STREAM", "0")
        _lib.tuning_reload()
        slow = op(f)
        monkeypatch.delenv("NG_FD_STREAM")
        _lib.tuning_reload()
        assert torch.equal(fast.view(torch.int64), slow.view(torch.int64)), (order, axis)
        del fast, slow


def _lap_blocks_vs_oracle(f_dev, out_dev, hs, blocks, margin=8):
    """Compare output planes [a, b) of axis 0 with the C oracle run on input planes [a-margin, b+margin)
    (clipped to the array: a real boundary stays a real boundary; an artificial cut is >= 6 planes away
    from every compared plane, further than the widest one-sided stencil reaches)."""
    from oracle import pencil_c
    nx = f_dev.shape[0]
    for a, b in blocks:
        lo, hi = max(a - margin, 0), min(b + margin, nx)
        ref = pencil_c.cartesian_laplacian3(f_dev[lo:hi].cpu().numpy(), hs, 4)[a - lo:b - lo]
        got = out_dev[a:b].cpu().numpy()
        assert np.array_equal(got, ref), ((a, b), rel_linf(got, ref))


def test_cfg5_1024_cube_vs_oracle(gpu_ready):
    """BASELINE cfg5 at its stated per-GPU size: the fused Laplacian of a 1024^3 field against the C/OpenMP
    oracle (oracle/pencil_c.c, pinned bit for bit to the NumPy restatement), array_equal -- on the whole array
    when the host has the memory (3 x 8 GiB), else on x blocks at both boundaries and in the middle."""
    import psutil
    import torch
    from numgrids_b200 import _lib
    from oracle import pencil_c
    n = 1024
    ax = ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0)
    g = ng.Grid(ax, ax, ax)
    x = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device="cuda")
    f = (torch.sin(2 * np.pi * x)[:, None, None] * torch.cos(2 * np.pi * x)[None, :, None]
         * torch.exp(x)[None, None, :]).contiguous()
    out = ng.CartesianLaplacian(g)(f)
    assert _lib.last_kernel() == "fdlap3d_ring2", _lib.last_kernel()
    hs = [ax.spacing] * 3
    _lap_blocks_vs_oracle(f, out, hs, [(0, 24), (500, 524), (1000, 1024)])
    if psutil.virtual_memory().available > 40 * 2 ** 30:
        pencil_c.use_all_cores()
        ref = pencil_c.cartesian_laplacian3(f.cpu().numpy(), hs, 4)
        got = out.cpu().numpy()
        assert np.array_equal(got, ref), rel_linf(got, ref)
        print("cfg5 1024^3: whole array equals the oracle bit for bit")


def test_cfg5_slab_shapes_vs_oracle(gpu_ready):
    """The per-GPU slab shapes of the 2/4/8-GPU weak-scaling runs (halos from random neighbour planes): the
    interior part (plain instance) + boundary part (slab instance) against the oracle applied to the array
    extended by the halo columns."""
    import torch
    from numgrids_b200 import _lib
    from numgrids_b200.slab import SlabLaplacian
    from oracle import pencil_c
    torch.manual_seed(7)
    for shape in ((96, 72, 256), (40, 200, 128), (64, 48, 64)):
        hs = [0.01, 0.02, 0.03]
        lap = SlabLaplacian(shape, hs)
        f = torch.randn(*shape, dtype=torch.float64, device="cuda")
        rows = shape[0] * shape[1]
        lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
        hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
        for use_lo, use_hi in ((True, True), (True, False), (False, True)):
            ext = torch.cat(([lo.view(shape[0], shape[1], 2)] if use_lo else []) + [f]
                            + ([hi.view(shape[0], shape[1], 2)] if use_hi else []), dim=2)
            ref = pencil_c.cartesian_laplacian3(ext.cpu().numpy(), hs, 4)
            ref = ref[:, :, (2 if use_lo else 0):(2 if use_lo else 0) + shape[2]]
            out = torch.full_like(f, float("nan"))
            lap.apply_part(f, out, 1, lo if use_lo else None, hi if use_hi else None)
            if shape[2] > 128:
                assert _lib.last_kernel() == "fdlap3d_ring2"            # interior tiles: plain instance
            lap.apply_part(f, out, 2, lo if use_lo else None, hi if use_hi else None)
            assert _lib.last_kernel() == "fdlap3d_ring2<slab>"
            assert np.array_equal(out.cpu().numpy(), ref), (shape, use_lo, use_hi)


def test_host_pipeline_chunked_paths(gpu_ready):
    """Arrays >= 64 MiB go through the chunked H2D / compute / D2H pipeline of the *_host entry
    points; results must equal the device-resident path bit for bit."""
    import os
    import torch
    from numgrids_b200 import _lib
    os.environ["NG_HOST_CHUNK_MB"] = "8"               # force chunking (default chunk is 256 MiB)
    _lib.tuning_reload()
    n = 224                                            # 224^3 doubles = 86 MiB -> 11 axis-0 chunks
    ax = ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0)
    g = ng.Grid(ax, ax, ax)
    rng = np.random.default_rng(5)
    f = rng.standard_normal((n, n, n))
    fd = torch.from_numpy(f).cuda()
    lap = ng.CartesianLaplacian(g)
    assert np.array_equal(lap(f), lap(fd).cpu().numpy())
    for axis in (0, 1, 2):                             # axis 0 runs unchunked, 1 and 2 chunked
        d = ng.Diff(g, 1, axis)
        assert np.array_equal(d(f), d(fd).cpu().numpy()), axis
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, 256, 0.0, 2 * np.pi)
    g2 = ng.Grid(ax, ax, p)
    f2 = rng.standard_normal(g2.shape)
    d = ng.Diff(g2, 1, 2)
    assert np.array_equal(d(f2), d(torch.from_numpy(f2).cuda()).cpu().numpy())
    del os.environ["NG_HOST_CHUNK_MB"]
    _lib.tuning_reload()


def test_laplacian_plane_ranges(gpu_ready):
    """ng_fdlap_apply_slab_range: any split of axis 0 reproduces the one-shot result."""
    import torch
    from numgrids_b200.slab import SlabLaplacian
    shape = (70, 48, 128)
    lap = SlabLaplacian(shape, [0.01, 0.02, 0.03])
    torch.manual_seed(2)
    f = torch.randn(*shape, dtype=torch.float64, device="cuda")
    ref = lap.apply_local(f, torch.empty_like(f))
    for cuts in ((0, 9, 70), (0, 1, 2, 35, 68, 69, 70), (0, 64, 70)):
        out = torch.full_like(f, float("nan"))
        for a, b in zip(cuts[:-1], cuts[1:]):
            lap.apply_range(f, out, a, b)
        assert torch.equal(out, ref), cuts


def test_laplacian_z_parts(gpu_ready):
    """Interior + boundary z parts (the halo/compute overlap split) cover the slab exactly once."""
    import torch
    from numgrids_b200.slab import SlabLaplacian
    for shape in ((24, 40, 256), (20, 16, 64), (18, 24, 128), (9, 11, 70)):
        lap = SlabLaplacian(shape, [0.01, 0.02, 0.03])
        torch.manual_seed(3)
        f = torch.randn(*shape, dtype=torch.float64, device="cuda")
        rows = shape[0] * shape[1]
        lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
        hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
        for halos in ((None, None), (lo, hi), (lo, None)):
            ref = lap.apply_local(f, torch.empty_like(f), *halos)
            out = torch.full_like(f, float("nan"))
            lap.apply_part(f, out, 1, *halos)
            lap.apply_part(f, out, 2, *halos)
            assert torch.equal(out, ref), (shape, [h is not None for h in halos])


def test_apply_batch_equals_single_applies(gpu_ready):
    """ng_apply_batch: K stacked fields in one launch are bit-identical to K single applies (FD, Chebyshev, FFT
    axes; contiguous and strided), and a stack of cfg1-size fields takes the streaming kernel."""
    import torch
    from numgrids_b200 import _lib
    torch.manual_seed(11)
    ax = ng.create_axis(A.EQUIDISTANT, 512, 0.0, 1.0)
    g = ng.Grid(ax, ax)
    fs = torch.randn(12, 512, 512, dtype=torch.float64, device="cuda")
    for order, axis in ((1, 0), (2, 1)):
        d = ng.Diff(g, order, axis)
        out = d.apply_batch(fs)
        assert _lib.last_kernel() in ("fd_march", "fd_row<u4>"), _lib.last_kernel()
        for k in range(12):
            assert torch.equal(out[k], d(fs[k])), (order, axis, k)
    axes = [ng.create_axis(A.CHEBYSHEV, 24, 0.0, 1.0), ng.create_axis(A.EQUIDISTANT_PERIODIC, 64, 0.0, 2 * np.pi),
            ng.create_axis(A.LOGARITHMIC, 20, 0.1, 10.0)]
    g3 = ng.Grid(*axes)
    fs = torch.randn(5, *g3.shape, dtype=torch.float64, device="cuda")
    for axis in range(3):
        d = ng.Diff(g3, 1, axis)
        out = d.apply_batch(fs)
        for k in range(5):
            assert torch.equal(out[k], d(fs[k])), (axis, k)
    with pytest.raises(ValueError):
        ng.Diff(g3, 1, 0).apply_batch(fs[:, :, :, :10])


def test_matrix_free_linear_operator(gpu_ready):
    """GridDiff.as_linear_operator(): matvec on the device equals the (lazily assembled) sparse matrix product."""
    axes = [ng.create_axis(A.CHEBYSHEV, 12, 0.0, 1.0), ng.create_axis(A.EQUIDISTANT, 16, 0.0, 2.0)]
    g = ng.Grid(*axes)
    rng = np.random.default_rng(4)
    v = rng.standard_normal(g.size)
    for order, axis in ((1, 0), (2, 1)):
        d = ng.Diff(g, order, axis)
        L = d.as_linear_operator()
        assert L.shape == (g.size, g.size)
        ref = d.as_matrix() @ v
        assert rel_linf(L @ v, ref) <= 1e-12
        assert rel_linf(L.rmatvec(v), d.as_matrix().T @ v) <= 1e-12


def _smooth(n):
    for q in (2, 3, 5, 7, 11, 13):
        while n % q == 0:
            n //= q
    return n == 1


def _exact_pencil(f1d, W):
    """real(ifft(W fft(f))) of one pencil through a long-double DFT: the exact answer to ~1e-19."""
    n = len(f1d)
    k = np.arange(n, dtype=np.longdouble)
    E = np.exp(-2j * np.pi * np.outer(k, k).astype(np.longdouble) / n)
    return np.real((np.conj(E) @ (W.astype(np.clongdouble) * (E @ f1d.astype(np.longdouble)))) / n).astype(np.float64)


@pytest.mark.parametrize("n", [65, 96, 100, 118, 150, 177, 200, 360, 500, 1000, 1536, 3000])
def test_periodic_axis_of_any_length_is_n_log_n(gpu_ready, n):
    """Periodic axes whose length is not a power of two never take the O(n^2) dense contraction for first and second
    derivatives: up to n = 512 a zero-padded power-of-two transform on the two-pass register kernels, beyond that the
    native mixed-radix transform (n = 2^a 3^b 5^c 7^d 11^e 13^f; csrc/fft_mixed.cu), which also serves every higher order;
    <= 1e-11 against numpy.fft (oracle) on the last axis and on a strided one, plain and fused (polar / spherical
    composites, which use the native transform wherever it exists)."""
    import torch
    from numgrids_b200 import _lib
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, n, 0.0, 2 * np.pi)
    e = ng.create_axis(A.EQUIDISTANT, 10, 0.0, 1.0)
    rng = np.random.default_rng(n)

    def check_kernel(order):
        kernel = _lib.last_kernel()
        if _smooth(n) and (n > 512 or order > 2 or n % 2 == 1 or 128 < n < 182):
            assert kernel == "fft_mixed", (n, order, kernel)
        elif order <= 2:
            assert "padded" in kernel, (n, order, kernel)
            assert ("fft2pass" in kernel) == (n <= 512), (n, kernel)
        else:
            assert "dense" in kernel, (n, order, kernel)

    for axes, axis in (((e, p), 1), ((p, e), 0)):
        g = ng.Grid(*axes)
        X = g.meshed_coords[axis]
        f = np.exp(np.sin(X)) * np.cos(2 * X) + 0.05 * rng.standard_normal(g.shape)
        out = ng.Diff(g, 1, axis)(torch.from_numpy(f).cuda())
        check_kernel(1)
        ref = pencil.fft_apply(f, axis, pencil.fft_multiplier(n, p.spacing, 1))
        err = rel_linf(out.cpu().numpy(), ref)
        assert err <= TOL_FFT, (n, axis, err)
    # orders 2 and 3.  Order k amplifies the rounding of ANY transform by (n/2)^k -- numpy.fft itself is 9e-11 from a
    # long-double DFT at n = 1000, order 2, and 1.6e-11 at n = 200, order 3 -- so the distance from numpy.fft is bounded by
    # the 1e-11 of the contract or by numpy.fft's own distance e from the exact answer: 0.5 e for order 2 (the dense
    # contraction measures the same: 9.7e-12 vs 1.2e-11 at n = 1000, tools/check_fft_order2_nonpow2.py), 1.5 e for
    # order 3 (ours + numpy's); and against the exact answer itself the result is no worse than 1.5 e.
    g = ng.Grid(e, p)
    P = g.meshed_coords[1]
    f = np.exp(np.sin(P)) * np.cos(2 * P) * (1.0 + g.meshed_coords[0])
    for order in (2, 3):
        out = ng.Diff(g, order, 1)(torch.from_numpy(f).cuda()).cpu().numpy()
        check_kernel(order)
        Wk = pencil.fft_multiplier(n, p.spacing, order)
        ref = pencil.fft_apply(f, 1, Wk)
        exact = _exact_pencil(f[3], Wk)
        e_numpy = rel_linf(ref[3], exact)
        bound = max(TOL_FFT, (0.5 if order == 2 else 1.5) * e_numpy)
        err = rel_linf(out, ref)
        assert err <= bound, (n, order, err, bound)
        assert rel_linf(out[3], exact) <= max(TOL_FFT, 1.5 * e_numpy), (n, order, rel_linf(out[3], exact), e_numpy)
    # fused: polar Laplacian with n points in phi (Chebyshev r)
    if n in (100, 118, 360):
        r = ng.create_axis(A.CHEBYSHEV, 12, 0.5, 2.0)
        pg = ng.PolarGrid(r, p)
        R, P = pg.meshed_coords
        f = R ** 2 * np.cos(2 * P) + np.exp(-R) * np.sin(P)
        pax = [pencil.make_axis("chebyshev", 12, 0.5, 2.0), pencil.make_axis("periodic", n, 0.0, 2 * np.pi)]
        D = [pencil.make_diff(pax, 1, i) for i in range(2)]
        ref = pencil.laplacian(f, (np.ones_like(R), R), D)
        assert rel_linf(pg.laplacian(f), ref) <= TOL_FFT
        # enough pencils for the persistent ring kernel (>= 1776): spherical composites with n points in phi --
        # fused prologue / double apply / tails on zero-padded transforms
        sg = ng.SphericalGrid(ng.create_axis(A.CHEBYSHEV, 48, 0.5, 2.0), ng.create_axis(A.CHEBYSHEV, 40, 0.3, np.pi - 0.3), p)
        R, T, P = sg.meshed_coords
        f = R ** 2 * np.sin(T) ** 2 * np.cos(2 * P) + np.exp(-R) * np.cos(T)
        pax = [pencil.make_axis("chebyshev", 48, 0.5, 2.0), pencil.make_axis("chebyshev", 40, 0.3, np.pi - 0.3),
               pencil.make_axis("periodic", n, 0.0, 2 * np.pi)]
        D = [pencil.make_diff(pax, 1, i) for i in range(3)]
        h = (np.ones_like(R), R, R * np.sin(T))
        # (this field is hard on a Laplacian: r^2 sin^2(theta) cos(2 phi) = x^2 - y^2 is harmonic, so its three terms cancel, and
        # the phi term carries 1 / (r sin theta)^2 <= 46: two correct n-point transforms differ by more than 1e-11 of the
        # result from n ~ 300 on.  Measured against the oracle at n = 360: native mixed-radix 2.3e-11, zero-padded 3.7e-11;
        # n = 500: 2.4e-11 / 1.1e-10; up to n = 128 both are within the 1e-11 of the contract)
        assert rel_linf(sg.laplacian(f), pencil.laplacian(f, h, D)) <= (TOL_FFT if n <= 128 else 5e-11)
        assert _lib.last_kernel() == ("fft_mixed" if _smooth(n) else "fft2pass_ring<padded>"), _lib.last_kernel()
        for got, want in zip(sg.gradient(f), pencil.gradient(f, h, D)):
            assert rel_linf(got, want) <= TOL_FFT
        v = (R, np.sin(T), np.cos(P))
        assert rel_linf(sg.divergence(*v), pencil.divergence(v, h, D)) <= TOL_FFT


@pytest.mark.parametrize("n", [6, 12, 45, 77, 143, 182, 210, 364, 1155, 2187, 6000])
def test_mixed_radix_transform_vs_oracle(gpu_ready, n, monkeypatch):
    """The native mixed-radix kernel on its own (every radix: 2 4 8 16 3 5 7 9 11 13; one to seven passes), forced for
    every length and order: <= 1e-11 against numpy.fft on the contiguous axis and on strided ones (odd extents, a ragged
    last CTA), with a fused prologue / tail, and with non-finite pencils, which must come out all-NaN without touching
    the pencil they share a complex transform with (numgrids/diff.py:133-143; curvilinear.py:236-244 relies on it)."""
    import torch
    from numgrids_b200 import _lib
    from numgrids_b200.diff import FFTDiff
    monkeypatch.setattr(FFTDiff, "NONPOW2", "native")
    monkeypatch.setattr(FFTDiff, "MIXED_MIN_POINTS", 3)
    p = ng.create_axis(A.EQUIDISTANT_PERIODIC, n, 0.0, 2 * np.pi)
    rng = np.random.default_rng(n)
    small = n <= 400
    for shape, axis in (((7, 9, n), 2), ((5, n, 11), 1), ((n, 3, 7), 0)) if small else (((3, 5, n), 2), ((n, 9), 0)):
        axes = [p if a == axis else ng.create_axis(A.EQUIDISTANT, m, 0.0, 1.0) for a, m in enumerate(shape)]
        g = ng.Grid(*axes)
        X = g.meshed_coords[axis]
        f = np.exp(np.sin(X)) * np.cos(2 * X) + 0.05 * rng.standard_normal(g.shape)
        for order in (1, 2, 3) if small else (1,):
            out = ng.Diff(g, order, axis)(torch.from_numpy(f).cuda()).cpu().numpy()
            assert _lib.last_kernel() == "fft_mixed", _lib.last_kernel()
            Wk = pencil.fft_multiplier(n, p.spacing, order)
            ref = pencil.fft_apply(f, axis, Wk)
            line = tuple(0 if a != axis else slice(None) for a in range(len(shape)))
            bound = max(TOL_FFT, 1.5 * rel_linf(ref[line], _exact_pencil(f[line], Wk))) if order > 1 and n <= 1200 else TOL_FFT
            assert rel_linf(out, ref) <= bound, (n, shape, axis, order)
        # non-finite samples
        fb = f.copy()
        idx = [tuple(int(rng.integers(0, m)) for m in shape) for _ in range(3)]
        for k, ix in enumerate(idx):
            fb[ix] = (np.inf, -np.inf, np.nan)[k]
        out = ng.Diff(g, 1, axis)(torch.from_numpy(fb).cuda()).cpu().numpy()
        ref = pencil.fft_apply(f, axis, pencil.fft_multiplier(n, p.spacing, 1))
        bad = np.zeros(shape, dtype=bool)
        for ix in idx:
            sl = list(ix); sl[axis] = slice(None)
            bad[tuple(sl)] = True
        assert np.isnan(out[bad]).all()
        assert np.max(np.abs(out[~bad] - ref[~bad])) <= TOL_FFT * np.max(np.abs(ref))
    # fused prologue / tails: polar composites with n points in phi against the oracle
    if n in (45, 182, 364):
        r = ng.create_axis(A.CHEBYSHEV, 14, 0.5, 2.0)
        pg = ng.PolarGrid(r, p)
        R, P = pg.meshed_coords
        f = R ** 2 * np.cos(2 * P) + np.exp(-R) * np.sin(P)
        pax = [pencil.make_axis("chebyshev", 14, 0.5, 2.0), pencil.make_axis("periodic", n, 0.0, 2 * np.pi)]
        D = [pencil.make_diff(pax, 1, i) for i in range(2)]
        h = (np.ones_like(R), R)
        assert rel_linf(pg.laplacian(f), pencil.laplacian(f, h, D)) <= TOL_FFT
        assert _lib.last_kernel() == "fft_mixed", _lib.last_kernel()
        for got, want in zip(pg.gradient(f), pencil.gradient(f, h, D)):
            assert rel_linf(got, want) <= TOL_FFT
        v = (R * np.cos(P), np.sin(2 * P) / R)
        assert rel_linf(pg.divergence(*v), pencil.divergence(v, h, D)) <= TOL_FFT
