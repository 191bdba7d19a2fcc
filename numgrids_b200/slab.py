"""Last-axis slab decomposition across the GPUs of one node.

The reference is single-process (no distributed code), so this layer has no
counterpart there; it follows BASELINE.json's north_star and SURVEY.md 8(e): one process per GPU,
the grid sharded along the LAST axis.

* ``SlabLaplacian`` -- the fused Cartesian Laplacian (cfg5): halo exchange of ``nb`` planes per
  side, interior z tiles computed under the exchange.
* ``SlabDiff`` -- any ``Diff(grid, order, axis)`` on a slab.  Unsharded axis: local.  Sharded
  finite-difference / log axis: halo exchange.  Sharded Chebyshev / FFT axis (whole pencils
  needed): slab <-> pencil redistribution (all-to-all), apply, and back -- ``Redistribution``.
* ``SlabCurvilinear`` -- gradient / divergence / laplacian / curl of a ``CurvilinearGrid`` on
  slabs, the reference's expressions (numgrids/curvilinear.py:169-317) with the elementwise work
  fused into the applies, the halo pack and the redistribution kernels.

Exchanges go through NVLink peer memory when torch symmetric memory is available (the pack /
redistribution kernels store straight into, or load straight from, the peers' buffers; only
signals are exchanged besides), else through ``torch.distributed`` (NCCL on the GPU box, gloo in
the CPU tests).  Ranks on the global boundary keep the one-sided stencils and every table is the
global axis' table, so each slab is bit-identical to the same slab of a single-GPU run.
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


class SymmetricSlabComm:
    """The peer-memory record of ``ng_slab_plan_create`` (include/numgrids_b200.h) from ONE torch symmetric-memory
    allocation per rank: two halo sets [set][side][rows][nb] followed by the two flag words.  torch is used to
    allocate and map the buffers only; pack, signals, stencil launches, streams and events are the C library's."""

    def __init__(self, rows, nb, device, rank, nranks, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.rank, self.nranks = rank, nranks
        halo = rows * nb                                   # doubles per halo buffer
        self.elems = 4 * halo + 2                          # + 2 flag words (8 bytes each, start at 0)
        self.buf = symm.empty((self.elems,), dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        lo_n, hi_n = neighbours(rank, nranks)
        base = self.buf.data_ptr()
        peer = lambda r: self.hdl.get_buffer(r, (self.elems,), torch.float64).data_ptr()
        off = lambda q, side: 8 * ((2 * q + side) * halo)
        flags = 8 * 4 * halo
        c = _lib.NgSlabComm()
        c.rank, c.nranks = rank, nranks
        for q in range(2):
            for side in range(2):
                c.my_halo[q][side] = base + off(q, side)
            c.peer_lo_halo[q] = peer(lo_n) + off(q, 1) if lo_n is not None else None
            c.peer_hi_halo[q] = peer(hi_n) + off(q, 0) if hi_n is not None else None
        c.my_flags = base + flags
        c.peer_lo_flag = peer(lo_n) + flags + 8 if lo_n is not None else None
        c.peer_hi_flag = peer(hi_n) + flags if hi_n is not None else None
        self.record = c
        torch.cuda.synchronize()
        self.hdl.barrier(channel=0)


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
        if self._use_cabi(f):
            _lib.check(_lib.load().ng_slab_lap_apply(self._cabi_plan.handle, f.data_ptr(), out.data_ptr(),
                                                     _device.current_stream_ptr()), "ng_slab_lap_apply")
            return out
        if self._use_symm(f):
            return self._call_symm(f, out)
        b = self.pack(f)
        lo_n, hi_n = neighbours(self.rank, self.nranks)
        lo_h = b["lo_halo"] if lo_n is not None else None
        hi_h = b["hi_halo"] if hi_n is not None else None
        tm = getattr(self, "stencil_events", None)        # optional sink for bench.py: one dict of events per step
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
            arrived = comm.record_event()
        ev = {k: t.cuda.Event(enable_timing=True) for k in ("int0", "int1", "edge0", "edge1", "join")} \
            if tm is not None else None
        if ev: ev["int0"].record()
        self.apply_part(f, out, 1)                        # interior z tiles: no halo dependence
        if ev: ev["int1"].record()
        main.wait_event(arrived)
        if ev: ev["edge0"].record()
        self.apply_part(f, out, 2, lo_h, hi_h)            # first / last z tile (or everything on fallbacks)
        if ev:
            ev["edge1"].record()
            ev["join"].record()
            tm.append(ev)
        return out

    # -- the whole step behind the C ABI (ng_slab_plan_create / ng_slab_lap_apply) ---------------------
    def _use_cabi(self, f):
        """Default multi-GPU path: torch symmetric memory only allocates and maps the halo / flag buffers; the step
        itself (pack, flags, the two stencil launches, side stream, events) is ``ng_slab_lap_apply``.
        NG_SLAB_EXCHANGE = cabi (default) | symm (the same sequence driven from Python, with per-part events) | nccl."""
        import os
        if os.environ.get("NG_SLAB_EXCHANGE", "cabi") != "cabi" or len(self.local_shape) != 3 \
                or getattr(self, "_cabi_failed", False) or getattr(self, "stencil_events", None) is not None:
            return False
        if getattr(self, "_cabi_plan", None) is None:
            try:
                self._cabi_comm = SymmetricSlabComm(self.rows, self.nb, f.device, self.rank, self.nranks, self.group)
                lib = _lib.load()
                children = [make_fd_plan(self.local_shape, a, 2, self.acc, self.spacings[a]) for a in range(3)]
                arr = _lib.ptr_array([p.handle.value for p in children])
                shp, shp_p = _lib.as_i64(list(self.local_shape))
                out = C.c_void_p()
                _lib.check(lib.ng_slab_plan_create(C.byref(self._cabi_comm.record), 3, shp_p, arr, C.byref(out)),
                           "ng_slab_plan_create")
                self._cabi_plan = _lib.Plan(out.value, keepalive=children)
            except Exception as e:          # no symmetric memory: the torch.distributed paths below
                self._cabi_failed = True
                self._cabi_error = repr(e)
                return False
        return True

    # -- NVLink peer-memory exchange (pack kernel stores straight into the neighbours' halos) -------
    def _use_symm(self, f):
        import os
        mode = os.environ.get("NG_SLAB_EXCHANGE", "cabi")
        if mode not in ("symm", "cabi") or len(self.local_shape) != 3 or getattr(self, "_symm_failed", False):
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
        """main stream: pack kernel (stores this rank's edge columns straight into the neighbours' halo
        buffers over NVLink), then the interior z tiles.  side stream, behind the pack: signals, then the two
        boundary z tiles.  The pack runs BEFORE the stencil rather than next to it: its 16-byte reads (one per
        row and side) are DRAM-activate bound, 0.06-0.2 ms alone but 2.5 ms when they compete with the
        stencil's streams (measured, profiles/r02_slab_step_parts.txt), which would delay the boundary launch.
        The boundary launch needs no SM until the interior launch has handed out its last CTAs, so it fills
        that launch's tail; the main stream joins the side stream at the end of the step.  Reuse of halo set
        q: my signal of step s+1 follows my boundary launch of step s on the side stream, my pack of step s+2
        follows the join of step s+1 -- behind my wait for the neighbours' signals of step s+1, which they
        sent after their boundary launch of step s (the last reader of set q)."""
        t = _device.torch()
        sh = self._symm
        q = sh.step & 1
        sh.step += 1
        main = t.cuda.current_stream()
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = t.cuda.Stream()
        comm = self._comm_stream
        tm = getattr(self, "stencil_events", None)        # optional sink for bench.py: one dict of events per step
        ev = {k: t.cuda.Event(enable_timing=True) for k in
              ("pack0", "pack1", "edge0", "edge1", "int0", "int1", "join")} if tm is not None else None
        lo_h, hi_h = sh.mine(q)
        to_lo, to_hi = sh.targets(q)
        if ev: ev["pack0"].record()
        _lib.check(_lib.load().ng_fd_pack_halo(
            self.plan.handle, f.data_ptr(), to_lo.data_ptr() if to_lo is not None else None,
            to_hi.data_ptr() if to_hi is not None else None, _device.current_stream_ptr()), "ng_fd_pack_halo")
        packed = main.record_event()
        if ev: ev["pack1"].record()
        with t.cuda.stream(comm):
            comm.wait_event(packed)
            sh.signal_and_wait(q)
            if ev: ev["edge0"].record()
            self.apply_part(f, out, 2, lo_h, hi_h)        # first / last z tile, neighbour planes staged by TMA
            if ev: ev["edge1"].record()
            done = comm.record_event()
        if ev: ev["int0"].record()
        self.apply_part(f, out, 1)                        # interior z tiles: never read a halo
        if ev: ev["int1"].record()
        main.wait_event(done)
        if ev:
            ev["join"].record()
            tm.append(ev)
        return out

    # -- host buffers in, host buffers out: chunked pipeline with per-chunk halo exchange ------------
    def apply_host(self, h_in, h_out, chunk_mib=256):
        """One distributed apply from this rank's slab in (pinned) HOST memory into `h_out` (host), the
        multi-GPU counterpart of ``ng_fdlap_apply_host``.  The slab is cut into chunks of axis-0 planes;
        per chunk k: H2D on a copy stream -> pack kernel stores the chunk's first/last z columns straight into
        the neighbours' halo buffers + signals (side stream) -> stencil of chunk k-1 (it needs chunk k's first
        two planes along x and its own halo rows) -> D2H of chunk k-1 on a second copy stream.  Both PCIe
        directions, the NVLink halo traffic and the kernels overlap; the call returns when h_out is complete."""
        t = _device.torch()
        if tuple(h_in.shape) != self.local_shape or tuple(h_out.shape) != self.local_shape:
            raise ValueError(f"host slabs must have the local shape {self.local_shape}")
        dev = t.device("cuda", t.cuda.current_device())
        st = getattr(self, "_host_state", None)
        if st is None:
            st = self._host_state = {
                "d_in": t.empty(self.local_shape, dtype=t.float64, device=dev),
                "d_out": t.empty(self.local_shape, dtype=t.float64, device=dev),
                "s_in": t.cuda.Stream(), "s_cmp": t.cuda.Stream(), "s_out": t.cuda.Stream(), "s_comm": t.cuda.Stream()}
        d_in, d_out = st["d_in"], st["d_out"]
        nx = self.local_shape[0]
        # the same chunking on every rank (the per-chunk signals pair up): sized from the widest slab
        zl_max = -(-self.global_shape[-1] // self.nranks)
        plane = int(np.prod(self.local_shape[1:-1])) * zl_max
        cp = max(8, (chunk_mib << 17) // plane)
        if self.nranks == 1 or not self._use_symm(d_in):
            # no peer memory: one exchange for the whole slab (NCCL / gloo), copies not overlapped with it
            d_in.copy_(h_in, non_blocking=True)
            self(d_in, out=d_out)
            h_out.copy_(d_out, non_blocking=True)
            t.cuda.current_stream().synchronize()
            return h_out
        chunks = [(a, min(a + cp, nx)) for a in range(0, nx, cp)]
        sh = self._symm
        q = sh.step & 1
        sh.step += 1
        to_lo, to_hi = sh.targets(q)
        lo_h, hi_h = sh.mine(q)
        lib = _lib.load()
        in_done, halo_done = [], []
        main = t.cuda.current_stream()
        for s_ in (st["s_in"], st["s_comm"], st["s_cmp"], st["s_out"]):
            s_.wait_stream(main)

        def run(c):
            a, b = chunks[c]
            with t.cuda.stream(st["s_cmp"]):
                st["s_cmp"].wait_event(in_done[min(c + 1, len(chunks) - 1)])
                st["s_cmp"].wait_event(halo_done[c])
                self.apply_range(d_in, d_out, a, b, lo_h, hi_h)
                done = st["s_cmp"].record_event()
            with t.cuda.stream(st["s_out"]):
                st["s_out"].wait_event(done)
                h_out[a:b].copy_(d_out[a:b], non_blocking=True)

        for k, (a, b) in enumerate(chunks):
            with t.cuda.stream(st["s_in"]):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
                in_done.append(st["s_in"].record_event())
            with t.cuda.stream(st["s_comm"]):
                st["s_comm"].wait_event(in_done[k])
                _lib.check(lib.ng_fdlap_pack_halo_range(
                    self.plan.handle, d_in.data_ptr(), to_lo.data_ptr() if to_lo is not None else None,
                    to_hi.data_ptr() if to_hi is not None else None, a, b, _device.current_stream_ptr()),
                    "ng_fdlap_pack_halo_range")
                sh.signal_and_wait(q)
                halo_done.append(st["s_comm"].record_event())
            if k >= 1:
                run(k - 1)
        run(len(chunks) - 1)
        for s_ in (st["s_in"], st["s_comm"], st["s_cmp"], st["s_out"]):
            s_.synchronize()
        return h_out

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


# =================================================================================================
# general per-axis applies on slabs
# =================================================================================================
def make_coef(tensor, strides):
    """ng_coef for a device table (torch tensor) with per-axis element strides (0 = broadcast)."""
    c = _lib.NgCoef()
    c.ptr = tensor.data_ptr() if tensor is not None else None
    for a in range(_lib.NG_MAX_DIM):
        c.stride[a] = int(strides[a]) if tensor is not None and a < len(strides) else 0
    return c


def make_fuse(pre=None, minuend=None, post=None, addend=None, add_zero=False, divisor=None, finite=False):
    """ng_fuse record; pre/post/divisor are (tensor, strides) pairs, minuend/addend CUDA tensors."""
    f = _lib.NgFuse()
    none = (None, ())
    f.pre = make_coef(*(pre or none))
    f.post = make_coef(*(post or none))
    f.divisor = make_coef(*(divisor or none))
    f.minuend = minuend.data_ptr() if minuend is not None else None
    f.addend = addend.data_ptr() if addend is not None else None
    f.add_zero = 1 if add_zero else 0
    f.finite = 1 if finite else 0
    f._keep = (pre, post, divisor, minuend, addend)
    return f


class Redistribution:
    """Host description of slab [A][B][zl] <-> pencil [A_r][B][Z] (axis 0 partitioned instead of the
    last axis; B = product of the middle extents).  Both directions are lists of strided 2-D blocks
    whose rows are contiguous z segments; `rank`'s slab block for peer s is the contiguous row range
    ``a_s0*B .. a_s1*B`` of its slab, so the all-to-all splits need no packing on the slab side."""

    def __init__(self, global_shape, rank, nranks):
        gs = tuple(int(x) for x in global_shape)
        if len(gs) < 2:
            raise ValueError("a 1-D grid cannot be redistributed: shard a grid with >= 2 axes")
        if gs[0] < nranks or gs[-1] < nranks:
            raise ValueError(f"axes 0 and -1 of {gs} need at least {nranks} points each")
        self.global_shape, self.rank, self.nranks = gs, rank, nranks
        self.A, self.Z = gs[0], gs[-1]
        self.B = int(np.prod(gs[1:-1])) if len(gs) > 2 else 1
        self.a = [partition(self.A, nranks, r) for r in range(nranks)]
        self.z = [partition(self.Z, nranks, r) for r in range(nranks)]
        self.zl = self.z[rank][1] - self.z[rank][0]
        self.al = self.a[rank][1] - self.a[rank][0]
        self.slab_shape = gs[:-1] + (self.zl,)
        self.pencil_shape = (self.al,) + gs[1:]
        self.pencil_elems_max = max(a1 - a0 for a0, a1 in self.a) * self.B * self.Z

    # all_to_all_single split sizes (elements)
    def slab_splits(self):      # slab-side chunks: rows a_s0..a_s1 of my slab, one per peer
        return [(a1 - a0) * self.B * self.zl for a0, a1 in self.a]

    def pencil_splits(self):    # pencil-side chunks: [A_r][B][zl_s], one per peer
        return [self.al * self.B * (z1 - z0) for z0, z1 in self.z]

    def _blk(self, src, dst, rows, row_len, ss, ds, off):
        b = _lib.NgCopyBlock()
        b.src, b.dst, b.rows, b.row_len = int(src), int(dst), int(rows), int(row_len)
        b.src_stride, b.dst_stride, b.local_offset = int(ss), int(ds), int(off)
        return b

    def push_blocks(self, slab_ptr, pencil_ptrs):
        """My slab rows for peer s -> columns z0_r.. of peer s's pencil buffer (pencil_ptrs[s])."""
        z0 = self.z[self.rank][0]
        out = []
        for s, (a0, a1) in enumerate(self.a):
            off = a0 * self.B * self.zl
            out.append(self._blk(slab_ptr + 8 * off, pencil_ptrs[s] + 8 * z0, (a1 - a0) * self.B, self.zl,
                                 self.zl, self.Z, off))
        return out

    def pull_blocks(self, pencil_ptrs, slab_ptr):
        """Columns z0_r.. of peer s's pencil result -> rows a_s0..a_s1 of my slab."""
        z0 = self.z[self.rank][0]
        out = []
        for s, (a0, a1) in enumerate(self.a):
            off = a0 * self.B * self.zl
            out.append(self._blk(pencil_ptrs[s] + 8 * z0, slab_ptr + 8 * off, (a1 - a0) * self.B, self.zl,
                                 self.Z, self.zl, off))
        return out

    def unpack_blocks(self, chunks_ptr, pencil_ptr):
        """all_to_all receive buffer (chunk s = [A_r][B][zl_s], contiguous) -> my pencil buffer."""
        out, pos = [], 0
        for (z0, z1) in self.z:
            out.append(self._blk(chunks_ptr + 8 * pos, pencil_ptr + 8 * z0, self.al * self.B, z1 - z0,
                                 z1 - z0, self.Z, 0))
            pos += self.al * self.B * (z1 - z0)
        return out

    def pack_blocks(self, pencil_ptr, chunks_ptr):
        """My pencil result -> all_to_all send buffer (chunk s = [A_r][B][zl_s])."""
        out, pos = [], 0
        for (z0, z1) in self.z:
            out.append(self._blk(pencil_ptr + 8 * z0, chunks_ptr + 8 * pos, self.al * self.B, z1 - z0,
                                 self.Z, z1 - z0, 0))
            pos += self.al * self.B * (z1 - z0)
        return out


def run_blocks(blocks, local_shape=None, fuse=None, fuse_side=0, stream=None):
    """One ng_slab_block_copy launch (<= 16 blocks per launch)."""
    lib = _lib.load()
    st = _device.current_stream_ptr() if stream is None else stream
    shp_p, nd = None, 0
    if fuse_side:
        shp, shp_p = _lib.as_i64(list(local_shape))
        nd = len(local_shape)
    for i in range(0, len(blocks), 16):
        part = blocks[i:i + 16]
        arr = (_lib.NgCopyBlock * len(part))(*part)
        _lib.check(lib.ng_slab_block_copy(len(part), arr, nd, shp_p, C.byref(fuse) if fuse_side else None,
                                          fuse_side, st), "ng_slab_block_copy")


class SymmetricPencils:
    """Pencil-layout input and result buffers in NVLink peer memory (torch symmetric memory).

    forward : every rank's block-copy kernel stores its slab rows straight into the peers' input
              buffers (P2P stores), then signals each peer (channel 8);
    backward: after the apply each rank signals "result ready" (channel 9) and every rank's kernel
              loads its columns straight out of the peers' result buffers (P2P loads).
    Reuse is ordered by the same signals: a peer's "result ready" of call k is sent after its apply
    read the input buffer, so call k+1 may overwrite it; a peer's channel-8 signal of call k+1 is
    sent after its pull of call k, so my apply of call k+1 may overwrite my result buffer."""

    def __init__(self, elems, device, rank, nranks, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.rank, self.nranks = rank, nranks
        self.buf = symm.empty((2, elems), dtype=torch.float64, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.peers = [self.hdl.get_buffer(r, (2, elems), torch.float64) for r in range(nranks)]
        torch.cuda.synchronize()
        self.hdl.barrier(channel=0)

    def in_ptrs(self):
        return [p[0].data_ptr() for p in self.peers]

    def out_ptrs(self):
        return [p[1].data_ptr() for p in self.peers]

    def all_signal_wait(self, channel, timeout_ms=20000):
        for r in range(self.nranks):
            if r != self.rank:
                self.hdl.put_signal(r, channel, timeout_ms)
        for r in range(self.nranks):
            if r != self.rank:
                self.hdl.wait_signal(r, channel, timeout_ms)


class EmulatedWorld:
    """All ranks of a slab decomposition inside ONE process on ONE GPU (GPU tests; inter-process
    waits cannot be tested without a multi-GPU box).  Pass it as `group` and run every rank's calls
    from its own Python thread: the "peer" buffers are ordinary tensors shared through this object
    and every signal exchange is a thread barrier, so the same pack / block-copy / apply kernels run
    in the same order as on N GPUs (all threads launch into the one default stream)."""

    def __init__(self, nranks):
        import threading
        self.nranks = nranks
        self.barrier = threading.Barrier(nranks)
        self.lock = threading.Lock()
        self.shared = {}
        self.counters = [0] * nranks

    def next_key(self, rank):
        self.counters[rank] += 1
        return self.counters[rank]

    def buffer(self, key, rank, shape, device):
        t = _device.torch()
        with self.lock:
            k = (key, rank)
            if k not in self.shared:
                self.shared[k] = t.zeros(shape, dtype=t.float64, device=device)
            return self.shared[k]

    def run(self, fn):
        """fn(rank) on one thread per rank; returns the list of results (re-raises the first error)."""
        import threading
        res, err = [None] * self.nranks, []

        def work(r):
            try:
                res[r] = fn(r)
            except BaseException as e:      # noqa: BLE001 -- reported to the caller below
                err.append(e)
                self.barrier.abort()
        th = [threading.Thread(target=work, args=(r,)) for r in range(self.nranks)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if err:
            raise err[0]
        return res


class _EmulatedHalo:
    def __init__(self, world, key, rows, nb, device, rank, nranks):
        self.world, self.rank, self.nranks, self.step = world, rank, nranks, 0
        self.lo_n, self.hi_n = neighbours(rank, nranks)
        shape = (2, 2, rows, nb)
        self.buf = world.buffer(key, rank, shape, device)
        self.peer_lo = world.buffer(key, self.lo_n, shape, device) if self.lo_n is not None else None
        self.peer_hi = world.buffer(key, self.hi_n, shape, device) if self.hi_n is not None else None

    targets = SymmetricHalo.targets
    mine = SymmetricHalo.mine

    def signal_and_wait(self, q, timeout_ms=0):
        self.world.barrier.wait()


class _EmulatedPencils:
    def __init__(self, world, key, elems, device, rank, nranks):
        self.world, self.rank, self.nranks = world, rank, nranks
        self.peers = [world.buffer(key, r, (2, elems), device) for r in range(nranks)]

    in_ptrs = SymmetricPencils.in_ptrs
    out_ptrs = SymmetricPencils.out_ptrs

    def all_signal_wait(self, channel, timeout_ms=0):
        self.world.barrier.wait()


class SlabDiff:
    """``Diff(grid, order, axis_index, acc)`` applied to this rank's last-axis slab of the GLOBAL grid."""

    def __init__(self, grid, order, axis_index=0, acc=4, rank=0, nranks=1, group=None):
        from .api import Diff
        from .diff import ChebyshevDiff, FFTDiff
        self.grid, self.rank, self.nranks, self.group = grid, rank, nranks, group
        self.diff = Diff(grid, order, axis_index, acc)             # host tables of the global axis
        self.op = self.diff.operator
        self.axis_index = axis_index
        gs = tuple(grid.shape)
        self.z0, self.z1 = partition(gs[-1], nranks, rank)
        self.local_shape = gs[:-1] + (self.z1 - self.z0,)
        self.sharded = nranks > 1 and axis_index == grid.ndims - 1
        self.dense = isinstance(self.op, (ChebyshevDiff, FFTDiff))
        self.redist = Redistribution(gs, rank, nranks) if self.sharded and self.dense else None
        self._plan = None
        self._halo = None
        self._symm = None
        self._symm_failed = False
        self._tmp = {}
        self._key = group.next_key(rank) if isinstance(group, EmulatedWorld) else None

    # -- plans -----------------------------------------------------------------------------------
    @property
    def plan(self) -> _lib.Plan:
        """Local plan: slab shape, or pencil shape when the sharded axis needs whole pencils."""
        if self._plan is None:
            _lib.require_gpu()
            if self.redist is not None:
                self._plan = self.op._make_plan(shape=self.redist.pencil_shape)
            elif hasattr(self.op, "_x_div") and self.op._x_div is not None and self.sharded:
                self._plan = self.op._make_plan(shape=self.local_shape, x_div=self.op._x_div[self.z0:self.z1])
            else:
                self._plan = self.op._make_plan(shape=self.local_shape)
        return self._plan

    @property
    def nb(self):
        return self.plan.halo_width

    def _scratch(self, name, elems, like):
        t = _device.torch()
        b = self._tmp.get(name)
        if b is None or b.numel() < elems or b.device != like.device:
            b = t.empty(int(elems), dtype=t.float64, device=like.device)
            self._tmp[name] = b
        return b

    # -- apply -----------------------------------------------------------------------------------
    def __call__(self, f, out=None, fuse=None):
        t = _device.torch()
        f = _device.as_device_tensor(f, self.local_shape)
        out = t.empty_like(f) if out is None else out
        lib = _lib.load()
        st = _device.current_stream_ptr()
        if not self.sharded:
            if fuse is None:
                _lib.check(lib.ng_apply(self.plan.handle, f.data_ptr(), out.data_ptr(), st), "ng_apply")
            else:
                _lib.check(lib.ng_apply_fused(self.plan.handle, f.data_ptr(), out.data_ptr(), C.byref(fuse), st),
                           "ng_apply_fused")
            return out
        if self.dense:
            return self._call_pencils(f, out, fuse)
        return self._call_halo(f, out, fuse)

    # sharded finite-difference axis: halo exchange
    def _call_halo(self, f, out, fuse):
        t = _device.torch()
        lib = _lib.load()
        nb, rows = self.nb, int(np.prod(self.local_shape[:-1]))
        if self.local_shape[-1] < 2 * nb:
            raise ValueError("slab too thin for the stencil")
        lo_n, hi_n = neighbours(self.rank, self.nranks)
        pre = C.byref(fuse.pre) if fuse is not None and fuse.pre.ptr else None
        use_symm = self._want_symm(f)
        if use_symm and self._halo is None:
            try:
                if isinstance(self.group, EmulatedWorld):
                    self._halo = _EmulatedHalo(self.group, self._key, rows, nb, f.device, self.rank, self.nranks)
                else:
                    self._halo = SymmetricHalo(rows, nb, f.device, self.rank, self.nranks, self.group)
            except Exception as e:
                self._symm_failed, self._symm_error = True, repr(e)
                use_symm = False
        st = _device.current_stream_ptr()
        if use_symm:
            sh = self._halo
            q = sh.step & 1
            sh.step += 1
            to_lo, to_hi = sh.targets(q)
            _lib.check(lib.ng_fd_pack_halo_fused(self.plan.handle, f.data_ptr(), pre,
                                                 to_lo.data_ptr() if to_lo is not None else None,
                                                 to_hi.data_ptr() if to_hi is not None else None, st),
                       "ng_fd_pack_halo_fused")
            sh.signal_and_wait(q)
            lo_h, hi_h = sh.mine(q)
        else:
            b = {k: self._scratch(k, rows * nb, f).view(rows, nb)
                 for k in ("lo_planes", "hi_planes", "lo_halo", "hi_halo")}
            _lib.check(lib.ng_fd_pack_halo_fused(self.plan.handle, f.data_ptr(), pre, b["lo_planes"].data_ptr(),
                                                 b["hi_planes"].data_ptr(), st), "ng_fd_pack_halo_fused")
            exchange_halos(b["lo_planes"], b["hi_planes"], b["lo_halo"], b["hi_halo"], self.rank, self.nranks,
                           self.group)
            lo_h = b["lo_halo"] if lo_n is not None else None
            hi_h = b["hi_halo"] if hi_n is not None else None
        _lib.check(lib.ng_apply_fused_slab(self.plan.handle, f.data_ptr(), out.data_ptr(),
                                           C.byref(fuse) if fuse is not None else None,
                                           lo_h.data_ptr() if lo_h is not None else None,
                                           hi_h.data_ptr() if hi_h is not None else None, st),
                   "ng_apply_fused_slab")
        return out

    def _want_symm(self, f):
        import os
        if isinstance(self.group, EmulatedWorld):
            return True
        return os.environ.get("NG_SLAB_EXCHANGE", "symm") == "symm" and not self._symm_failed

    # sharded Chebyshev / FFT axis: slab -> pencils, apply, pencils -> slab
    def _call_pencils(self, f, out, fuse):
        import torch.distributed as dist
        lib = _lib.load()
        rd = self.redist
        st = _device.current_stream_ptr()
        has_pre = fuse is not None and bool(fuse.pre.ptr)
        has_tail = fuse is not None and bool(fuse.post.ptr or fuse.divisor.ptr or fuse.minuend or fuse.addend
                                             or fuse.add_zero or fuse.finite)
        use_symm = self._want_symm(f)
        if use_symm and self._symm is None:
            try:
                if isinstance(self.group, EmulatedWorld):
                    self._symm = _EmulatedPencils(self.group, self._key, rd.pencil_elems_max, f.device,
                                                  self.rank, self.nranks)
                else:
                    self._symm = SymmetricPencils(rd.pencil_elems_max, f.device, self.rank, self.nranks, self.group)
            except Exception as e:
                self._symm_failed, self._symm_error = True, repr(e)
                use_symm = False
        if use_symm:
            sp = self._symm
            run_blocks(rd.push_blocks(f.data_ptr(), sp.in_ptrs()), self.local_shape, fuse, 1 if has_pre else 0)
            sp.all_signal_wait(8)
            _lib.check(lib.ng_apply(self.plan.handle, sp.in_ptrs()[self.rank], sp.out_ptrs()[self.rank], st),
                       "ng_apply")
            sp.all_signal_wait(9)
            run_blocks(rd.pull_blocks(sp.out_ptrs(), out.data_ptr()), self.local_shape, fuse, 2 if has_tail else 0)
            return out
        # torch.distributed all-to-all (NCCL / gloo); slab-side chunks are contiguous row ranges
        n_slab = int(np.prod(self.local_shape))
        n_pen = int(np.prod(rd.pencil_shape))
        src = f
        if has_pre:                                    # pre * f into a send buffer (one fused local copy)
            src = self._scratch("send", n_slab, f)[:n_slab]
            zl = self.local_shape[-1]
            run_blocks([rd._blk(f.data_ptr(), src.data_ptr(), n_slab // zl, zl, zl, zl, 0)],
                       self.local_shape, fuse, 1)
        chunks = self._scratch("chunks", n_pen, f)[:n_pen]
        dist.all_to_all_single(chunks, src.reshape(-1), rd.pencil_splits(), rd.slab_splits(), group=self.group)
        pin = self._scratch("pin", n_pen, f)[:n_pen]
        pout = self._scratch("pout", n_pen, f)[:n_pen]
        run_blocks(rd.unpack_blocks(chunks.data_ptr(), pin.data_ptr()))
        _lib.check(lib.ng_apply(self.plan.handle, pin.data_ptr(), pout.data_ptr(), st), "ng_apply")
        run_blocks(rd.pack_blocks(pout.data_ptr(), chunks.data_ptr()))
        if has_tail:
            back = self._scratch("send", n_slab, f)[:n_slab]
            dist.all_to_all_single(back, chunks, rd.slab_splits(), rd.pencil_splits(), group=self.group)
            zl = self.local_shape[-1]
            run_blocks([rd._blk(back.data_ptr(), out.data_ptr(), n_slab // zl, zl, zl, zl, 0)],
                       self.local_shape, fuse, 2)
        else:
            dist.all_to_all_single(out.reshape(-1), chunks, rd.slab_splits(), rd.pencil_splits(), group=self.group)
        return out


def local_coefficients(grid, z0, z1):
    """Scale factors and Jacobian of a CurvilinearGrid on the slab [.., z0:z1): the user's callables
    evaluated on the slab's own meshed coordinates (elementwise, so identical to a slice of the
    global evaluation, reference numgrids/curvilinear.py:126-140)."""
    coords = [a.coords for a in grid.axes]
    coords[-1] = coords[-1][z0:z1]
    mesh = np.meshgrid(*coords, indexing="ij")
    h = tuple(fn(mesh) for fn in grid._scale_factor_fns)
    J = np.ones(mesh[0].shape)
    for hi in h:
        J = J * hi
    return h, J


class SlabCurvilinear:
    """gradient / divergence / laplacian / curl of a (global) ``CurvilinearGrid`` on last-axis slabs.

    Same sequences and fused elementwise steps as csrc/curv.cu (which serves the single-GPU
    ``CurvilinearGrid`` methods); each per-axis apply is a :class:`SlabDiff`, so the sharded axis
    is served by a halo exchange (finite differences) or a pencil redistribution (Chebyshev / FFT)."""

    def __init__(self, grid, rank=0, nranks=1, group=None):
        self.grid, self.rank, self.nranks = grid, rank, nranks
        self.ndims = grid.ndims
        self.D = [SlabDiff(grid, 1, i, 4, rank, nranks, group) for i in range(self.ndims)]
        self.local_shape = self.D[0].local_shape
        self.z0, self.z1 = self.D[0].z0, self.D[0].z1
        self._h = None
        self._coefs = {}
        self._work = None

    def _host(self):
        if self._h is None:
            self._h, self._J = local_coefficients(self.grid, self.z0, self.z1)
        return self._h, self._J

    def _coef(self, key, make, device):
        c = self._coefs.get(key)
        if c is None:
            t = _device.torch()
            with np.errstate(divide="ignore", invalid="ignore"):
                full = make()
            values, strides = _tables.compact_coefficient(full, self.local_shape)
            c = (t.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).to(device), tuple(strides))
            self._coefs[key] = c
        return c

    def _x(self, a):
        return _device.as_device_tensor(a, self.local_shape)

    def gradient(self, f):
        h, _ = self._host()
        x = self._x(f)
        t = _device.torch()
        outs = []
        for i in range(self.ndims):
            g = self._coef(("grad", i), lambda i=i: 1.0 / h[i], x.device)
            outs.append(self.D[i](x, t.empty_like(x), make_fuse(post=g, finite=True)))
        return tuple(outs)

    def laplacian(self, f):
        h, J = self._host()
        x = self._x(f)
        t = _device.torch()
        out = t.empty_like(x)
        if self._work is None or self._work.device != x.device:
            self._work = t.empty_like(x)
        cJ = self._coef(("J",), lambda: J, x.device)
        for i in range(self.ndims):
            c = self._coef(("lap", i), lambda i=i: J / h[i] ** 2, x.device)
            self.D[i](x, self._work, make_fuse(post=c))
            last = i == self.ndims - 1
            self.D[i](self._work, out, make_fuse(addend=out if i else None, add_zero=(i == 0),
                                                 divisor=cJ if last else None, finite=last))
        return out

    def divergence(self, *v):
        if len(v) != self.ndims:
            raise ValueError(f"Expected {self.ndims} vector components, got {len(v)}.")
        h, J = self._host()
        xs = [self._x(a) for a in v]
        t = _device.torch()
        out = t.empty_like(xs[0])
        cJ = self._coef(("J",), lambda: J, out.device)
        for i in range(self.ndims):
            c = self._coef(("div", i), lambda i=i: J / h[i], out.device)
            last = i == self.ndims - 1
            self.D[i](xs[i], out, make_fuse(pre=c, addend=out if i else None, add_zero=(i == 0),
                                            divisor=cJ if last else None, finite=last))
        return out

    def curl(self, *v):
        if self.ndims not in (2, 3):
            raise ValueError("Curl is only defined for 2D and 3D grids.")
        h, _ = self._host()
        xs = [self._x(a) for a in v]
        t = _device.torch()
        pairs = [(0, 1)] if self.ndims == 2 else [(1, 2), (2, 0), (0, 1)]
        outs = []
        for c, (j, k) in enumerate(pairs):
            hk = self._coef(("h", k), lambda k=k: h[k], xs[0].device)
            hj = self._coef(("h", j), lambda j=j: h[j], xs[0].device)
            cc = self._coef(("curl", c), lambda j=j, k=k: 1.0 / (h[j] * h[k]), xs[0].device)
            o = t.empty_like(xs[0])
            self.D[j](xs[k], o, make_fuse(pre=hk))
            self.D[k](xs[j], o, make_fuse(pre=hj, minuend=o, post=cc, finite=True))
            outs.append(o)
        return outs[0] if self.ndims == 2 else tuple(outs)
