"""Streaming FD kernels (fd_march_kernel / fd_row_kernel) against the gather kernel, bit for bit, over random
shapes, orders, accuracies, fused prologues/epilogues and slab halos: python tools/check_fd_stream.py [cases]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402
from numgrids_b200 import _lib  # noqa: E402

A = ng.AxisType
lib = _lib.load()
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 200
bad = 0
kernels = {}


def coef(t, shape, vary):
    """NgCoef for a device table that varies along the axes in `vary` (others stride 0)."""
    c = _lib.NgCoef()
    c.ptr = t.data_ptr()
    st, acc = [0] * 4, 1
    for a in reversed(range(len(shape))):
        if a in vary:
            st[a] = acc
            acc *= shape[a]
    for a in range(4):
        c.stride[a] = st[a]
    return c


def table(shape, vary):
    n = int(np.prod([shape[a] for a in vary])) if vary else 1
    return torch.from_numpy(rng.uniform(0.5, 2.0, n)).cuda()


for case in range(ncase):
    ndim = int(rng.integers(1, 5))
    shape = [int(2 * rng.integers(1, 40)) if rng.random() < 0.8 else int(rng.integers(2, 70)) for _ in range(ndim)]
    axis = int(rng.integers(0, ndim))
    order = int(rng.integers(1, 4))
    acc = int(rng.choice([2, 4, 6]))
    kind = A.LOGARITHMIC if rng.random() < 0.2 else A.EQUIDISTANT
    need = 2 * ((order + 1) // 2) - 1 + acc + 2
    shape[axis] = max(shape[axis], need + int(rng.integers(0, 200)))
    if rng.random() < 0.3:
        shape[axis] += int(rng.integers(0, 600))
    axes = [ng.create_axis(kind if a == axis else A.EQUIDISTANT, n, 0.1, 2.0) for a, n in enumerate(shape)]
    g = ng.Grid(*axes)
    op = ng.Diff(g, order, axis, acc=acc).operator.operator
    f = torch.from_numpy(rng.standard_normal(shape)).cuda()
    fuse = None
    keep = []
    if rng.random() < 0.6:
        fuse = _lib.NgFuse()
        all_axes = list(range(ndim))
        if rng.random() < 0.6:
            v = [a for a in all_axes if rng.random() < 0.6]
            t = table(shape, v); keep.append(t); fuse.pre = coef(t, shape, v)
        if rng.random() < 0.6:
            v = [a for a in all_axes if rng.random() < 0.6]
            t = table(shape, v); keep.append(t); fuse.post = coef(t, shape, v)
        if rng.random() < 0.4:
            v = [a for a in all_axes if rng.random() < 0.6]
            t = table(shape, v); keep.append(t); fuse.divisor = coef(t, shape, v)
        if rng.random() < 0.4:
            t = torch.from_numpy(rng.standard_normal(shape)).cuda(); keep.append(t); fuse.addend = t.data_ptr()
        elif rng.random() < 0.3:
            fuse.add_zero = 1
        if rng.random() < 0.3:
            t = torch.from_numpy(rng.standard_normal(shape)).cuda(); keep.append(t); fuse.minuend = t.data_ptr()
        fuse.finite = int(rng.random() < 0.5)
    halo = None
    if axis == ndim - 1 and rng.random() < 0.5:
        nb = lib.ng_plan_halo_width(op.plan.handle)
        rows = int(np.prod(shape[:-1])) if ndim > 1 else 1
        mode = int(rng.integers(0, 3))
        lo = torch.from_numpy(rng.standard_normal((rows, nb))).cuda() if mode != 1 else None
        hi = torch.from_numpy(rng.standard_normal((rows, nb))).cuda() if mode != 0 else None
        halo = (lo, hi)
    outs = []
    for stream_on in ("1", "0"):
        os.environ["NG_FD_STREAM"] = stream_on
        # contiguous-axis cases: the shared-memory tile kernel (fd_tile.cu) on two cases of three, also for plain applies
        # and at any size; the shuffle kernel (fd_row.cu) on the third
        os.environ["NG_FD_TILE"], os.environ["NG_FD_TILE_MIN"] = ("2", "0") if case % 3 else ("0", "0")
        lib.ng_tuning_reload()
        o = torch.full(shape, float("nan"), dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        if halo is not None:
            lo, hi = halo
            rc = lib.ng_apply_fused_slab(op.plan.handle, f.data_ptr(), o.data_ptr(), C.byref(fuse) if fuse else None,
                                         lo.data_ptr() if lo is not None else None, hi.data_ptr() if hi is not None else None, st)
        elif fuse is not None:
            rc = lib.ng_apply_fused(op.plan.handle, f.data_ptr(), o.data_ptr(), C.byref(fuse), st)
        else:
            rc = lib.ng_apply(op.plan.handle, f.data_ptr(), o.data_ptr(), st)
        _lib.check(rc, "apply")
        if stream_on == "1":
            kernels[_lib.last_kernel()] = kernels.get(_lib.last_kernel(), 0) + 1
        outs.append(o)
    same = torch.equal(outs[0].view(torch.int64), outs[1].view(torch.int64))
    if not same:
        bad += 1
        d = (outs[0] - outs[1]).abs()
        print(f"MISMATCH case {case}: shape {shape} axis {axis} order {order} acc {acc} kind {kind} fuse {fuse is not None} "
              f"halo {None if halo is None else [h is not None for h in halo]} nbad {int((outs[0].view(torch.int64) != outs[1].view(torch.int64)).sum())} "
              f"maxdiff {float(torch.nan_to_num(d).max()):.3e}")
print("kernels:", kernels)
print(f"{ncase} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
