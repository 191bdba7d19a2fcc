"""Strided-pencil FFT kernel (periodic axis not last) against the radix-2 shared-memory fallback and numpy.fft,
plain and with fused prologue / epilogue, including singular (non-finite) rows: python tools/check_fft_strided.py [cases]"""
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
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 60
worst = 0.0
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
    ndim = int(rng.integers(2, 5))
    n = int(rng.choice([64, 128, 256, 512, 1024]))
    axis = int(rng.integers(0, ndim - 1))
    shape = [int(rng.integers(1, 13)) for _ in range(ndim)]
    shape[axis] = n
    shape[-1] = int(rng.integers(1, 40))
    order = int(rng.integers(1, 3))
    axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == axis else A.EQUIDISTANT, m, 0.0, 2 * np.pi if a == axis else 1.0)
            if m > 1 or a == axis else ng.create_axis(A.EQUIDISTANT, 2, 0.0, 1.0) for a, m in enumerate(shape)]
    shape = [len(a) for a in axes]
    g = ng.Grid(*axes)
    op = ng.Diff(g, order, axis).operator.operator
    fh = rng.standard_normal(shape)
    f = torch.from_numpy(fh).cuda()
    fuse, keep = None, []
    inplace = False
    if rng.random() < 0.6 and shape[-1] % 2 == 0:
        fuse = _lib.NgFuse()
        for name in ("pre", "post", "divisor"):
            if rng.random() < 0.5:
                v = [a for a in range(ndim) if rng.random() < 0.6]
                cnt = int(np.prod([shape[a] for a in v])) if v else 1
                t = torch.from_numpy(rng.uniform(0.5, 2.0, cnt)).cuda(); keep.append(t)
                setattr(fuse, name, coef(t, shape, v))
        if rng.random() < 0.4:
            t = torch.from_numpy(rng.standard_normal(shape)).cuda(); keep.append(t); fuse.addend = t.data_ptr()
        if rng.random() < 0.3:
            t = torch.from_numpy(rng.standard_normal(shape)).cuda(); keep.append(t); fuse.minuend = t.data_ptr()
        fuse.finite = int(rng.random() < 0.5)
        if rng.random() < 0.35:                              # the composites' in-place accumulate: out = out + D(pre * f)
            fuse.addend, fuse.minuend, fuse.finite = 1, None, 0          # (addend is pointed at the output below)
            fuse.divisor = _lib.NgCoef()
            inplace = True
    if rng.random() < 0.3:                                   # a singular pencil: all-NaN out, neighbours clean
        idx = [int(rng.integers(0, s)) for s in shape]
        idx[axis] = int(rng.integers(0, n))
        f[tuple(idx)] = float("inf")
    outs = []
    acc0 = torch.from_numpy(rng.standard_normal(shape)).cuda() if inplace else None
    # tile kernel (fft_tile.cu: plain and fused applies where it takes them) / strided register kernel / radix-2 fallback
    for mode, tile in (("1", "1"), ("1", "0"), ("0", "0")):
        os.environ["NG_FFT_STRIDED"], os.environ["NG_FFT_TILE"] = mode, tile
        lib.ng_tuning_reload()
        o = acc0.clone() if inplace else torch.full(shape, 7.0, dtype=torch.float64, device="cuda")
        if inplace:
            fuse.addend = o.data_ptr()
        st = torch.cuda.current_stream().cuda_stream
        rc = (lib.ng_apply_fused(op.plan.handle, f.data_ptr(), o.data_ptr(), C.byref(fuse), st) if fuse is not None
              else lib.ng_apply(op.plan.handle, f.data_ptr(), o.data_ptr(), st))
        _lib.check(rc, "apply")
        if tile == "1":
            kernels[_lib.last_kernel()] = kernels.get(_lib.last_kernel(), 0) + 1
        outs.append(o)
    b = outs[2]
    for which, a in (("tile", outs[0]), ("strided", outs[1])):
        nan_same = torch.equal(torch.isnan(a), torch.isnan(b))
        fin = ~torch.isnan(b)
        den = float(b[fin].abs().max()) if fin.any() else 1.0
        err = float((a[fin] - b[fin]).abs().max() / max(den, 1e-300)) if fin.any() else 0.0
        worst = max(worst, err)
        if not nan_same or err > 5e-13:
            bad += 1
            print(f"MISMATCH case {case} ({which}): shape {shape} axis {axis} order {order} fuse {fuse is not None} inplace {inplace} nan_same {nan_same} err {err:.3e}")
print("kernels:", kernels)
print(f"{ncase} cases, {bad} mismatches, worst rel diff vs the radix-2 kernel {worst:.2e}")
sys.exit(1 if bad else 0)
