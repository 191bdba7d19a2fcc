"""General dense (Chebyshev) kernel against the fallback kernel, bit for bit, over random shapes / orders / axes /
fused prologues and epilogues; also forced on exact shapes against the exact-shape kernel.
python tools/check_dense_general.py [cases] [seed]"""
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
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
kernels = {}


def coef(t, shape, vary):
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


for case in range(ncase):
    ndim = int(rng.integers(1, 5))
    axis = int(rng.integers(0, ndim))
    shape = [int(2 * rng.integers(1, 24)) if rng.random() < 0.85 else int(rng.integers(1, 30)) for _ in range(ndim)]
    exact = rng.random() < 0.15
    shape[axis] = int(rng.choice([128, 256])) if exact else int(rng.integers(3, 300))
    if exact and ndim > 1:
        other = ndim - 1 if axis != ndim - 1 else 0
        shape[other] = 32 * int(rng.integers(1, 4))
    order = int(rng.integers(1, 4))
    axes = [ng.create_axis(A.CHEBYSHEV if a == axis else A.EQUIDISTANT, max(m, 2), 0.0, 1.5) for a, m in enumerate(shape)]
    shape = [len(a) for a in axes]
    g = ng.Grid(*axes)
    op = ng.Diff(g, order, axis).operator.operator
    f = torch.from_numpy(rng.standard_normal(shape)).cuda()
    fuse, keep = None, []
    if rng.random() < 0.6:
        fuse = _lib.NgFuse()
        for name in ("pre", "post", "divisor"):
            if rng.random() < 0.5:
                v = [a for a in range(ndim) if rng.random() < 0.6]
                cnt = int(np.prod([shape[a] for a in v])) if v else 1
                t = torch.from_numpy(rng.uniform(0.5, 2.0, cnt)).cuda(); keep.append(t)
                setattr(fuse, name, coef(t, shape, v))
        if rng.random() < 0.4:
            t = torch.from_numpy(rng.standard_normal(shape)).cuda(); keep.append(t); fuse.addend = t.data_ptr()
        elif rng.random() < 0.3:
            fuse.add_zero = 1
        if rng.random() < 0.3:
            t = torch.from_numpy(rng.standard_normal(shape)).cuda(); keep.append(t); fuse.minuend = t.data_ptr()
        fuse.finite = int(rng.random() < 0.5)
    outs = []
    # modes: "2" = general panel kernel, "0" = fallback kernel, "1" = exact-shape panel kernel
    for mode in ("2", "0") + (("1",) if exact else ()):
        os.environ["NG_DENSE_FAST"] = mode
        lib.ng_tuning_reload()
        o = torch.full(shape, float("nan"), dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        rc = (lib.ng_apply_fused(op.plan.handle, f.data_ptr(), o.data_ptr(), C.byref(fuse), st) if fuse is not None
              else lib.ng_apply(op.plan.handle, f.data_ptr(), o.data_ptr(), st))
        _lib.check(rc, "apply")
        kernels[_lib.last_kernel()] = kernels.get(_lib.last_kernel(), 0) + 1
        outs.append(o)
    for k, o in enumerate(outs[1:]):
        if not torch.equal(outs[0].view(torch.int64), o.view(torch.int64)):
            bad += 1
            d = torch.nan_to_num(outs[0] - o).abs().max().item()
            nb = int((outs[0].view(torch.int64) != o.view(torch.int64)).sum())
            print(f"MISMATCH case {case} vs mode {('0', '1')[k]}: shape {shape} axis {axis} order {order} fuse {fuse is not None} nbad {nb} maxdiff {d:.3e}")
print("kernels:", kernels)
print(f"{ncase} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
