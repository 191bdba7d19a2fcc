"""Integration over the whole grid -- the reference's numgrids/integration.py ``Integral``.

The reference (numgrids/integration.py:33-72) inverts, per axis, the 1-D first-derivative matrix with
its first row and column removed, and on every call multiplies the array by the Kronecker-lifted
inverse only to keep the last row; a periodic axis is ``spacing * sum``.  Both are a fixed weight
vector per axis, so here ``__init__`` derives the vectors on the host with the reference's own
recipe (same ``scipy.sparse.linalg.inv`` of the same matrix, so the same bits) and ``__call__`` is
ONE streaming pass of ``ng_quad_apply`` (csrc/quad.cu) over the array, which may live on the device.
No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.sparse import csc_matrix
from scipy.sparse.linalg import inv

from . import _device, _lib
from .axes import EquidistantAxis
from .grids import Grid


def axis_weights(axis) -> np.ndarray:
    """Quadrature weights of one axis, as the reference's ``Integral`` implies them."""
    n = len(axis)
    if isinstance(axis, EquidistantAxis) and axis.periodic:
        return np.full(n, axis.spacing)                       # integration.py:61-63
    from .api import Diff
    D = Diff(Grid(axis), 1, 0).as_matrix()
    D_inv = inv(csc_matrix(D[1:, 1:]))                        # integration.py:36-39
    w = np.zeros(n)
    w[1:] = np.asarray(D_inv.toarray())[-1]                   # `[-1, ...]` of integration.py:70
    return w


class Integral:
    """The integration operator ``\\int_V ... dV`` over the whole grid."""

    def __init__(self, grid) -> None:
        self.grid = grid
        self.weights = [axis_weights(a) for a in grid.axes]
        self._dev = None

    def _device_tables(self):
        if self._dev is None:
            _lib.require_gpu()
            t = _device.torch()
            lib = _lib.load()
            w = [_device.to_device(np.ascontiguousarray(x)) for x in self.weights]
            work = t.empty(int(lib.ng_quad_workspace_doubles()), dtype=t.float64, device="cuda")
            ptrs = (C.c_void_p * len(w))(*[x.data_ptr() for x in w])
            self._dev = (w, work, ptrs)
        return self._dev

    def __call__(self, f):
        """Integral of a meshed array; a float for host input, a 0-d CUDA tensor for device input."""
        lib = _lib.load()
        shape = self.grid.shape
        if _device.is_complex(f):            # the quadrature is real and linear
            re, im = _device.split_complex(f)
            a, b = self(re), self(im)
            return a + 1j * b
        if _device.is_device_array(f):
            x = _device.as_device_tensor(f, shape)
        else:
            x = _device.to_device(_device.as_host_array(f, shape))
        w, work, ptrs = self._device_tables()
        out = _device.torch().empty((), dtype=_device.torch().float64, device="cuda")
        shp, shp_p = _lib.as_i64(list(shape))
        _lib.check(lib.ng_quad_apply(len(shape), shp_p, ptrs, x.data_ptr(), work.data_ptr(), out.data_ptr(),
                                     _device.current_stream_ptr()), "ng_quad_apply")
        if _device.is_device_array(f):
            return out
        return np.float64(out.item())
