"""ctypes front-end of oracle/pencil_c.c (C/OpenMP oracle) -- TEST INFRASTRUCTURE ONLY.

Same signatures as the NumPy oracle's applies; pinned against it by tests/test_oracle_c.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle import build_c
from oracle.findiff_port import coefficients

_lib = None


class _Tables(C.Structure):
    _fields_ = [("n_center", C.c_int), ("n_side", C.c_int),
                ("wc", C.c_void_p), ("wf", C.c_void_p), ("wb", C.c_void_p),
                ("oc", C.c_void_p), ("of", C.c_void_p), ("ob", C.c_void_p),
                ("inv_h_pow", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        path = build_c.LIB if os.path.exists(build_c.LIB) else build_c.build()
        _lib = C.CDLL(path)
        _lib.po_max_threads.restype = C.c_int
    return _lib


def threads() -> int:
    return int(lib().po_max_threads())


def use_all_cores() -> int:
    """Ask OpenMP for every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    lib().po_set_threads(n)
    return threads()


def _tables(h, order, acc):
    c = coefficients(order, acc)
    keep = []
    t = _Tables()
    t.n_center = len(c["center"]["coefficients"])
    t.n_side = len(c["forward"]["coefficients"])
    for name, key in (("wc", "center"), ("wf", "forward"), ("wb", "backward")):
        w = np.array(c[key]["coefficients"], dtype=np.float64)
        w[np.abs(1 - w) < 1e-14] = 1.0        # findiff adds y instead of w*y: same as w == 1.0 exactly
        o = np.array(c[key]["offsets"], dtype=np.int32)
        keep += [w, o]
        setattr(t, name, w.ctypes.data)
        setattr(t, "o" + name[1], o.ctypes.data)
    t.inv_h_pow = float(1.0 / h ** order)
    return t, keep


def _view(f, axis):
    outer = int(np.prod(f.shape[:axis], dtype=np.int64))
    inner = int(np.prod(f.shape[axis + 1:], dtype=np.int64))
    return outer, f.shape[axis], inner


def fd_apply(f, axis, h, order, acc=4, x_div=None):
    f = np.ascontiguousarray(f, dtype=np.float64)
    out = np.empty_like(f)
    t, keep = _tables(h, order, acc)
    o, n, i = _view(f, axis)
    xd = np.ascontiguousarray(x_div, dtype=np.float64) if x_div is not None else None
    lib().po_fd_apply(C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), C.c_int64(o), C.c_int64(n),
                      C.c_int64(i), C.byref(t), C.c_void_p(xd.ctypes.data if xd is not None else None))
    return out


def dense_apply(f, axis, M, descending):
    f = np.ascontiguousarray(f, dtype=np.float64)
    M = np.ascontiguousarray(M, dtype=np.float64)
    out = np.empty_like(f)
    o, n, i = _view(f, axis)
    lib().po_dense_apply(C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), C.c_int64(o), C.c_int64(n),
                         C.c_int64(i), C.c_void_p(M.ctypes.data), C.c_int(int(bool(descending))))
    return out


def cartesian_laplacian3(f, hs, acc=4):
    f = np.ascontiguousarray(f, dtype=np.float64)
    assert f.ndim == 3
    out = np.empty_like(f)
    tabs = [_tables(h, 2, acc) for h in hs]
    lib().po_fdlap3(C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), *[C.c_int64(s) for s in f.shape],
                    C.byref(tabs[0][0]), C.byref(tabs[1][0]), C.byref(tabs[2][0]))
    return out
