"""Last-axis slab decomposition of the fused Cartesian Laplacian across the GPUs of one node.

The reference is single-process (no distributed code), so this layer has no
counterpart there; it follows BASELINE.json's north_star: one process per GPU, the grid sharded
along the LAST axis, derivatives along unsharded axes local, the sharded finite-difference axis
served by a halo exchange of ``nb = len(center)//2`` planes per side (``torch.distributed`` P2P,
NCCL over NVLink on the GPU box, gloo in the CPU tests).  Ranks on the global boundary keep the
one-sided stencils, so every slab is bit-identical to the same slab of a single-GPU run.

Per step and rank: pack kernel (first/last nb planes -> two contiguous [rows][nb] buffers),
grouped isend/irecv with the +-1 neighbours, then ONE fused stencil launch that reads the received
halos for the nb boundary columns.  Halo volume is rows*nb*8 B per side (67 MB at 2048x2048x256)
against 17 GB of local stencil traffic.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _device, _lib, _tables


def partition(n_global: int, nranks: int, rank: int) -> tuple[int, int]:
    """[start, stop) of `rank`'s slab along the sharded axis (remainder spread over the low ranks)."""
    if not (0 <= rank < nranks):
        raise ValueError(f"rank {rank} out of range for {nranks} ranks")
    base, rem = divmod(n_global, nranks)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def neighbours(rank: int, nranks: int) -> tuple[int | None, int | None]:
    """(lower, upper) neighbour ranks; None on the global (non-periodic) boundary."""
    return (rank - 1 if rank > 0 else None, rank + 1 if rank < nranks - 1 else None)


def exchange_halos(lo_planes, hi_planes, lo_halo, hi_halo, rank, nranks, group=None):
    """Send my first planes down / last planes up; receive the neighbours' into lo_halo / hi_halo.

    Works on any torch.distributed backend (tensors on the backend's device).  Returns the list of
    pending work handles (already waited).
    """
    import torch.distributed as dist
    lo_n, hi_n = neighbours(rank, nranks)
    ops = []
    if lo_n is not None:
        ops.append(dist.P2POp(dist.isend, lo_planes, lo_n, group))
        ops.append(dist.P2POp(dist.irecv, lo_halo, lo_n, group))
    if hi_n is not None:
        ops.append(dist.P2POp(dist.isend, hi_planes, hi_n, group))
        ops.append(dist.P2POp(dist.irecv, hi_halo, hi_n, group))
    if not ops:
        return []
    works = dist.batch_isend_irecv(ops)
    for w in works:
        w.wait()
    return works


def make_fd_plan(shape, axis_index, order, acc, h) -> _lib.Plan:
    """FD plan for an explicit local shape and a GLOBAL spacing h (the slab keeps the global axis' bits)."""
    lib = _lib.load()
    s = _tables.fd_schemes(order, acc)
    (oc, wc), (of, wf), (ob, wb) = s["center"], s["forward"], s["backward"]
    w = [_lib.as_f64(_tables.snap_unit_weights(x)) for x in (wc, wf, wb)]
    o = [_lib.as_i32(x) for x in (oc, of, ob)]
    shp, shp_p = _lib.as_i64(list(shape))
    out = C.c_void_p()
    _lib.check(lib.ng_fd_plan_create(len(shape), shp_p, axis_index, len(oc), w[0][1], o[0][1], len(of),
                                     w[1][1], o[1][1], w[2][1], o[2][1], float(1.0 / h ** order),
                                     None, C.byref(out)), "ng_fd_plan_create")
    return _lib.Plan(out.value)


class SymmetricHalo:
    """Halo buffers in NVLink peer memory (torch symmetric memory): a rank's pack kernel stores its
    boundary planes DIRECTLY into the neighbours' halo buffers (P2P stores over NVLink/NVSwitch),
    followed by a per-neighbour signal -- pack and exchange are one kernel, no NCCL send/recv
    kernels compete with the stencil for SMs.  Two buffer sets alternate per step; the signal chain
    orders a rank's overwrite of set q after the neighbour's last read of it (the neighbour signals
    step s only after its stencil of step s-1 was enqueued ahead on the same stream dependency).
    """

    def __init__(self, rows, nb, device, rank, nranks, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.rank, self.nranks = rank, nranks
        self.shape = (2, 2, rows, nb)                     # [set][side: 0 = lo halo, 1 = hi halo][rows][nb]
        self.buf = symm.empty(self.shape, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.lo_n, self.hi_n = neighbours(rank, nranks)
        self.peer_lo = self.hdl.get_buffer(self.lo_n, self.shape, torch.float64) if self.lo_n is not None else None
        self.peer_hi = self.hdl.get_buffer(self.hi_n, self.shape, torch.float64) if self.hi_n is not None else None
        self.step = 0
        torch.cuda.synchronize()
        self.hdl.barrier(channel=0)

    def targets(self, q):
        """Where my first / last planes go: the lower neighbour's hi halo, the upper neighbour's lo halo."""
        return (self.peer_lo[q, 1] if self.peer_lo is not None else None,
                self.peer_hi[q, 0] if self.peer_hi is not None else None)

    def mine(self, q):
        return (self.buf[q, 0] if self.lo_n is not None else None,
                self.buf[q, 1] if self.hi_n is not None else None)

    def signal_and_wait(self, q, timeout_ms=20000):
        """After the pack kernel on the current stream: tell both neighbours their halo set q is
        complete, then wait for theirs (channels 2q / 2q+1 keep the two directions apart)."""
        if self.hi_n is not None:
            self.hdl.put_signal(self.hi_n, 2 * q, timeout_ms)          # upward traffic
        if self.lo_n is not None:
            self.hdl.put_signal(self.lo_n, 2 * q + 1, timeout_ms)      # downward traffic
        if self.lo_n is not None:
            self.hdl.wait_signal(self.lo_n, 2 * q, timeout_ms)
        if self.hi_n is not None:
            self.hdl.wait_signal(self.hi_n, 2 * q + 1, timeout_ms)


class SlabLaplacian:
    """Fused Cartesian Laplacian on this rank's last-axis slab of a global equidistant grid."""

    def __init__(self, global_shape, spacings, rank=0, nranks=1, acc=4, group=None):
        self.global_shape = tuple(int(s) for s in global_shape)
        self.spacings = tuple(spacings)
        self.rank, self.nranks, self.group, self.acc = rank, nranks, group, acc
        self.z0, self.z1 = partition(self.global_shape[-1], nranks, rank)
        self.local_shape = self.global_shape[:-1] + (self.z1 - self.z0,)
        self.nb = len(_tables.fd_schemes(2, acc)["center"][0]) // 2
        if self.local_shape[-1] < max(2 * self.nb, _tables.fd_min_points(2, acc)):
            raise ValueError("slab too thin for the stencil")
        self.rows = int(np.prod(self.local_shape[:-1]))
        self._plan = None
        self._bufs = None

    @classmethod
    def for_grid(cls, grid, rank=0, nranks=1, acc=4, group=None):
        from .axes import EquidistantAxis
        for a in grid.axes:
            if not isinstance(a, EquidistantAxis) or a.periodic:
                raise TypeError("SlabLaplacian needs equidistant, non-periodic axes")
        return cls(grid.shape, [a.spacing for a in grid.axes], rank, nranks, acc, group)

    @property
    def plan(self) -> _lib.Plan:
        if self._plan is None:
            lib = _lib.load()
            _lib.require_gpu()
            nd = len(self.local_shape)
            children = [make_fd_plan(self.local_shape, a, 2, self.acc, self.spacings[a]) for a in range(nd)]
            arr = _lib.ptr_array([p.handle.value for p in children])
            shp, shp_p = _lib.as_i64(list(self.local_shape))
            out = C.c_void_p()
            _lib.check(lib.ng_fdlap_plan_create(nd, shp_p, arr, C.byref(out)), "ng_fdlap_plan_create")
            self._plan = _lib.Plan(out.value, keepalive=children)
        return self._plan

    def _buffers(self, like):
        if self._bufs is None:
            t = _device.torch()
            mk = lambda: t.empty((self.rows, self.nb), dtype=t.float64, device=like.device)
            self._bufs = {k: mk() for k in ("lo_planes", "hi_planes", "lo_halo", "hi_halo")}
        return self._bufs

    def pack(self, f):
        b = self._buffers(f)
        _lib.check(_lib.load().ng_fd_pack_halo(self.plan.handle, f.data_ptr(), b["lo_planes"].data_ptr(),
                                               b["hi_planes"].data_ptr(), _device.current_stream_ptr()),
                   "ng_fd_pack_halo")
        return b

    def apply_local(self, f, out, lo_halo=None, hi_halo=None):
        lo = lo_halo.data_ptr() if lo_halo is not None else None
        hi = hi_halo.data_ptr() if hi_halo is not None else None
        _lib.check(_lib.load().ng_fdlap_apply_slab(self.plan.handle, f.data_ptr(), out.data_ptr(), lo, hi,
                                                   _device.current_stream_ptr()), "ng_fdlap_apply_slab")
        return out

    def apply_range(self, f, out, x_begin, x_end, lo_halo=None, hi_halo=None):
        """Output planes [x_begin, x_end) of axis 0 only (halos are needed for those planes only)."""
        lo = lo_halo.data_ptr() if lo_halo is not None else None
        hi = hi_halo.data_ptr() if hi_halo is not None else None
        _lib.check(_lib.load().ng_fdlap_apply_slab_range(self.plan.handle, f.data_ptr(), out.data_ptr(), lo, hi,
                                                         int(x_begin), int(x_end), _device.current_stream_ptr()),
                   "ng_fdlap_apply_slab_range")
        return out

    def apply_part(self, f, out, z_part, lo_halo=None, hi_halo=None):
        """z_part 1: the columns whose stencils never read a neighbour rank (no halo needed);
        z_part 2: the remaining boundary columns; 0: everything."""
        lo = lo_halo.data_ptr() if lo_halo is not None else None
        hi = hi_halo.data_ptr() if hi_halo is not None else None
        _lib.check(_lib.load().ng_fdlap_apply_slab_part(self.plan.handle, f.data_ptr(), out.data_ptr(), lo, hi,
                                                        int(z_part), _device.current_stream_ptr()),
                   "ng_fdlap_apply_slab_part")
        return out

    def __call__(self, f, out=None, overlap=True):
        """One distributed apply.  `f` is the local CUDA slab.

        pack (one kernel) -> halo exchange -> fused stencil.  With `overlap` the exchange runs on a
        side stream while the main stream computes the interior z tiles (which never read a halo);
        the two boundary z tiles follow once the neighbours' planes have arrived.
        """
        t = _device.torch()
        f = _device.as_device_tensor(f, self.local_shape)
        out = t.empty_like(f) if out is None else out
        if self.nranks == 1:
            return self.apply_local(f, out)
        if self._use_symm(f):
            return self._call_symm(f, out)
        b = self.pack(f)
        lo_n, hi_n = neighbours(self.rank, self.nranks)
        lo_h = b["lo_halo"] if lo_n is not None else None
        hi_h = b["hi_halo"] if hi_n is not None else None
        tm = getattr(self, "stencil_events", None)        # optional [(e0, e1, e2, e3), ...] sink for bench.py
        if not overlap or len(self.local_shape) != 3:
            exchange_halos(b["lo_planes"], b["hi_planes"], b["lo_halo"], b["hi_halo"], self.rank, self.nranks,
                           self.group)
            return self.apply_local(f, out, lo_h, hi_h)
        main = t.cuda.current_stream()
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = t.cuda.Stream()
        comm = self._comm_stream
        comm.wait_stream(main)                            # the pack kernel
        with t.cuda.stream(comm):
            exchange_halos(b["lo_planes"], b["hi_planes"], b["lo_halo"], b["hi_halo"], self.rank, self.nranks,
                           self.group)
            ev = comm.record_event()
        if tm is not None:
            e = [t.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
        self.apply_part(f, out, 1, lo_h, hi_h)            # interior z tiles: no halo dependence
        if tm is not None:
            e[1].record()
        main.wait_event(ev)
        if tm is not None:
            e[2].record()
        self.apply_part(f, out, 2, lo_h, hi_h)            # first / last z tile (or everything on fallbacks)
        if tm is not None:
            e[3].record()
            tm.append(tuple(e))
        return out

    # -- NVLink peer-memory exchange (pack kernel stores straight into the neighbours' halos) -------
    def _use_symm(self, f):
        import os
        mode = os.environ.get("NG_SLAB_EXCHANGE", "symm")
        if mode != "symm" or len(self.local_shape) != 3 or getattr(self, "_symm_failed", False):
            return False
        if getattr(self, "_symm", None) is None:
            try:
                self._symm = SymmetricHalo(self.rows, self.nb, f.device, self.rank, self.nranks, self.group)
            except Exception as e:          # symmetric memory not available: NCCL send/recv path
                self._symm_failed = True
                self._symm_error = repr(e)
                return False
        return True

    def _call_symm(self, f, out):
        t = _device.torch()
        sh = self._symm
        q = sh.step & 1
        sh.step += 1
        main = t.cuda.current_stream()
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = t.cuda.Stream()
        comm = self._comm_stream
        comm.wait_stream(main)            # f is ready; my previous stencil (reader of set q) is ordered before
        with t.cuda.stream(comm):
            to_lo, to_hi = sh.targets(q)
            _lib.check(_lib.load().ng_fd_pack_halo(
                self.plan.handle, f.data_ptr(), to_lo.data_ptr() if to_lo is not None else None,
                to_hi.data_ptr() if to_hi is not None else None, _device.current_stream_ptr()), "ng_fd_pack_halo")
            sh.signal_and_wait(q)
            ev = comm.record_event()
        lo_h, hi_h = sh.mine(q)
        tm = getattr(self, "stencil_events", None)
        if tm is not None:
            e = [t.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
        self.apply_part(f, out, 1, lo_h, hi_h)            # interior z tiles under the exchange
        if tm is not None:
            e[1].record()
        main.wait_event(ev)
        if tm is not None:
            e[2].record()
        self.apply_part(f, out, 2, lo_h, hi_h)
        if tm is not None:
            e[3].record()
            tm.append(tuple(e))
        return out

    # -- single-GPU emulation of all ranks (tests; no inter-process waits, see B200_PROFILING.md) ----
    @staticmethod
    def emulate_on_one_gpu(grid, f_global, nranks, acc=4):
        """Split `f_global` (CUDA tensor) into nranks slabs, exchange halos by copy, apply each slab."""
        slabs = [SlabLaplacian.for_grid(grid, r, nranks, acc) for r in range(nranks)]
        parts = [f_global[..., s.z0:s.z1].contiguous() for s in slabs]
        packed = []
        for s, p in zip(slabs, parts):
            b = s.pack(p)
            packed.append((b["lo_planes"].clone(), b["hi_planes"].clone()))
        outs = []
        for r, (s, p) in enumerate(zip(slabs, parts)):
            lo = packed[r - 1][1] if r > 0 else None              # lower neighbour's LAST planes
            hi = packed[r + 1][0] if r < nranks - 1 else None     # upper neighbour's FIRST planes
            outs.append(s.apply_local(p, _device.torch().empty_like(p), lo, hi))
        return outs
