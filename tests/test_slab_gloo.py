"""CPU, world_size 2 and 3 over gloo: the slab layer's host logic (partition, neighbour pattern,
halo exchange) reproduces the single-process oracle slab by slab.

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
