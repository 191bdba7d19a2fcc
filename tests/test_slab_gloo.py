"""CPU, world_size 2 and 3 over gloo: the slab layer's host logic (partition, neighbour pattern,
halo exchange, slab <-> pencil redistribution with its all-to-all splits and block descriptors)
reproduces the single-process oracle slab by slab.

The CUDA pack/stencil kernels are replaced by torch slicing and the oracle stencil here (no GPU in
this tier); the GPU tier checks the same decomposition with the real kernels
(tests/test_gpu_configs.py::test_cfg5_*).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from numgrids_b200 import slab
from oracle import pencil

SHAPE = (10, 9, 26)
SPACINGS = (0.1, 0.125, 0.04)


def _field():
    x, y, z = (np.arange(n) * h for n, h in zip(SHAPE, SPACINGS))
    return np.sin(2 * x)[:, None, None] * np.cos(3 * y)[None, :, None] * np.exp(z)[None, None, :]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nb = 2
        f = _field()
        ref = pencil.cartesian_laplacian(f, SPACINGS, 4)
        z0, z1 = slab.partition(SHAPE[-1], world, rank)
        local = np.ascontiguousarray(f[..., z0:z1])
        rows = local.shape[0] * local.shape[1]
        lo_planes = torch.from_numpy(local[..., :nb].reshape(rows, nb).copy())
        hi_planes = torch.from_numpy(local[..., -nb:].reshape(rows, nb).copy())
        lo_halo = torch.full((rows, nb), np.nan, dtype=torch.float64)
        hi_halo = torch.full((rows, nb), np.nan, dtype=torch.float64)
        slab.exchange_halos(lo_planes, hi_planes, lo_halo, hi_halo, rank, world)
        lo_n, hi_n = slab.neighbours(rank, world)
        pieces = []
        if lo_n is not None:
            assert np.array_equal(lo_halo.numpy().reshape(*local.shape[:2], nb), f[..., z0 - nb:z0])
            pieces.append(lo_halo.numpy().reshape(*local.shape[:2], nb))
        pieces.append(local)
        if hi_n is not None:
            assert np.array_equal(hi_halo.numpy().reshape(*local.shape[:2], nb), f[..., z1:z1 + nb])
            pieces.append(hi_halo.numpy().reshape(*local.shape[:2], nb))
        padded = np.concatenate(pieces, axis=2)
        out = pencil.cartesian_laplacian(padded, SPACINGS, 4)
        a = nb if lo_n is not None else 0
        out = out[..., a:a + (z1 - z0)]
        q.put((rank, bool(np.array_equal(out, ref[..., z0:z1]))))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_matches_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(world))
    assert got == {r: True for r in range(world)}


def test_partition_and_neighbours():
    assert [slab.partition(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [slab.partition(2048, 8, r) for r in (0, 7)] == [(0, 256), (1792, 2048)]
    assert slab.neighbours(0, 1) == (None, None)
    assert slab.neighbours(0, 4) == (None, 1) and slab.neighbours(3, 4) == (2, None)
    with pytest.raises(ValueError):
        slab.partition(10, 2, 2)
    s = slab.SlabLaplacian((64, 64, 64), (0.1, 0.1, 0.1), rank=1, nranks=4)
    assert s.local_shape == (64, 64, 16) and s.nb == 2 and s.rows == 4096
    with pytest.raises(ValueError):
        slab.SlabLaplacian((64, 64, 16), (0.1, 0.1, 0.1), rank=0, nranks=4)


# ---------------------------------------------------------------------------------------------
# slab <-> pencil redistribution (Chebyshev / FFT along the sharded axis)
# ---------------------------------------------------------------------------------------------
RSHAPE = (7, 5, 11)            # uneven in both partitioned axes for world = 2 and 3


def _exec_blocks(blocks):
    """CPU stand-in of ng_slab_block_copy (plain copy): run the descriptors with memmove."""
    import ctypes as C
    for b in blocks:
        for r in range(b.rows):
            C.memmove(b.dst + 8 * r * b.dst_stride, b.src + 8 * r * b.src_stride, 8 * b.row_len)


def _rfield():
    return np.arange(np.prod(RSHAPE), dtype=np.float64).reshape(RSHAPE) * 0.25 + 1.0


def _redist_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        f = _rfield()
        n = RSHAPE[-1]
        M = pencil.chebyshev_matrix(n, 0.0, 1.0, 2, True)
        ref = pencil.chebyshev_apply(f, 2, M, False)
        rd = slab.Redistribution(RSHAPE, rank, world)
        z0, z1 = rd.z[rank]
        a0, a1 = rd.a[rank]
        local = np.ascontiguousarray(f[..., z0:z1])
        assert local.shape == rd.slab_shape
        # forward all-to-all: slab-side chunks are contiguous row ranges of the slab itself
        chunks = torch.empty(int(np.prod(rd.pencil_shape)), dtype=torch.float64)
        dist.all_to_all_single(chunks, torch.from_numpy(local).reshape(-1), rd.pencil_splits(), rd.slab_splits())
        pin = np.full(rd.pencil_shape, np.nan)
        _exec_blocks(rd.unpack_blocks(chunks.data_ptr(), pin.ctypes.data))
        ok_fwd = bool(np.array_equal(pin, f[a0:a1]))
        pout = np.ascontiguousarray(pencil.chebyshev_apply(pin, 2, M, False))
        _exec_blocks(rd.pack_blocks(pout.ctypes.data, chunks.data_ptr()))
        out = torch.full((int(np.prod(rd.slab_shape)),), float("nan"), dtype=torch.float64)
        dist.all_to_all_single(out, chunks, rd.slab_splits(), rd.pencil_splits())
        ok_bwd = bool(np.array_equal(out.numpy().reshape(rd.slab_shape), ref[..., z0:z1]))
        q.put((rank, ok_fwd and ok_bwd))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pencil_redistribution_matches_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_redist_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(world))
    assert got == {r: True for r in range(world)}


def test_push_pull_descriptors_cover_everything_once():
    """Peer-memory form (push into the peers' pencil buffers, pull out of their results), all ranks
    emulated in one process with NumPy buffers."""
    for world in (1, 2, 3, 4):
        f = _rfield()
        rds = [slab.Redistribution(RSHAPE, r, world) for r in range(world)]
        slabs = [np.ascontiguousarray(f[..., rd.z[rd.rank][0]:rd.z[rd.rank][1]]) for rd in rds]
        pens = [np.full(rd.pencil_elems_max, np.nan) for rd in rds]
        for rd, s in zip(rds, slabs):
            _exec_blocks(rd.push_blocks(s.ctypes.data, [p.ctypes.data for p in pens]))
        for rd, p in zip(rds, pens):
            a0, a1 = rd.a[rd.rank]
            n = int(np.prod(rd.pencil_shape))
            assert np.array_equal(p[:n].reshape(rd.pencil_shape), f[a0:a1])
        outs = [np.full(rd.slab_shape, np.nan) for rd in rds]
        for rd, o in zip(rds, outs):
            _exec_blocks(rd.pull_blocks([p.ctypes.data for p in pens], o.ctypes.data))
        for s, o in zip(slabs, outs):
            assert np.array_equal(s, o)
        offs = [b.local_offset for b in rds[0].push_blocks(0, [0] * world)]
        assert offs == [a0 * rds[0].B * rds[0].zl for a0, _ in rds[0].a]
    with pytest.raises(ValueError):
        slab.Redistribution((2, 5, 11), 0, 3)
    with pytest.raises(ValueError):
        slab.Redistribution((11,), 0, 2)
