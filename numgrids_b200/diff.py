"""Per-axis differentiation operators -- same classes as the reference's numgrids/diff.py, with the
arithmetic running in libnumgrids_b200.so (hand-written sm_100a kernels) instead of
findiff / numpy.fft / scipy.sparse.

Protocol kept from the reference (numgrids/diff.py:26-28, :48-61): a ``GridDiff`` carries
``axis, order, grid, axis_index``, sets ``self.operator`` to a callable array -> array, and may
provide ``as_matrix()``.  The callable here is a :class:`DeviceOperator` bound to a lazily created
C plan; ``as_matrix()`` is assembled lazily on the host from the same 1-D tables (the reference
builds an ``n x grid.size`` Kronecker matrix eagerly in ``__init__``, numgrids/diff.py:128, :189).

There is no CPU fallback: calling an operator without a CUDA device raises NumgridsB200Error.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse

from . import _device, _lib, _tables
from .axes import EquidistantAxis, LogAxis


class DeviceOperator:
    """Callable ``f -> D f`` backed by an ``ng_plan`` (created on first use)."""

    STAGING_KEEP_BYTES = 1 << 30     # host-array applies of at least this size free their device staging arrays afterwards

    def __init__(self, shape, plan_factory, flexible_shape=False, real_result_only=False):
        self.shape = tuple(int(s) for s in shape)
        self._factory = plan_factory
        self._plan = None
        self._flex = flexible_shape
        self.real_result_only = real_result_only     # FFTDiff: np.real(ifft(..)) drops the imaginary part

    # -- plan ----------------------------------------------------------
    @property
    def plan(self) -> _lib.Plan:
        if self._plan is None:
            self._plan = self._factory()
        return self._plan

    # -- device-pointer level (used by composites and the slab layer) ---
    def apply_ptr(self, d_in: int, d_out: int, stream: int | None = None, fuse=None) -> None:
        lib = _lib.load()
        st = _device.current_stream_ptr() if stream is None else stream
        if fuse is None:
            _lib.check(lib.ng_apply(self.plan.handle, d_in, d_out, st), "ng_apply")
        else:
            _lib.check(lib.ng_apply_fused(self.plan.handle, d_in, d_out, C.byref(fuse), st),
                       "ng_apply_fused")

    def apply_batch(self, fs):
        """`fs`: CUDA array of shape (K, *grid.shape) -- K fields stacked; returns the K derivatives in one
        launch (``ng_apply_batch``).  For grids so small that one apply is launch-bound (BASELINE cfg1)."""
        x = _device.as_device_tensor(fs)
        if tuple(x.shape[1:]) != self.shape or x.dim() != len(self.shape) + 1:
            raise ValueError(f"expected a stack of shape (K, {', '.join(map(str, self.shape))}), got {tuple(x.shape)}")
        with _device.device_of(x):
            out = _device.empty_like_device(x)
            _lib.check(_lib.load().ng_apply_batch(self.plan.handle, x.data_ptr(), out.data_ptr(), int(x.shape[0]),
                                                  _device.current_stream_ptr()), "ng_apply_batch")
        return out

    # -- user level ----------------------------------------------------
    def __call__(self, f):
        if _device.is_complex(f):
            # The operator is real and linear.  FD / log / Chebyshev in the reference differentiate complex
            # fields componentwise (findiff's `w * y`, scipy's sparse matvec); FFTDiff keeps `np.real(...)` of
            # the result (numgrids/diff.py:141), i.e. the derivative of the real part as a REAL array.
            re, im = _device.split_complex(f)
            if self.real_result_only:
                return self(re)
            return _device.join_complex(self(re), self(im))
        if _device.is_device_array(f):
            x = _device.as_device_tensor(f, self.shape)
            with _device.device_of(x):
                out = _device.empty_like_device(x)
                self.apply_ptr(x.data_ptr(), out.data_ptr())
            return out
        a = _device.as_host_array(f, self.shape)
        _lib.require_gpu()
        out = np.empty_like(a)
        _lib.check(_lib.load().ng_apply_host(self.plan.handle, a.ctypes.data, out.ctypes.data),
                   "ng_apply_host")
        if a.nbytes >= self.STAGING_KEEP_BYTES:         # big arrays: hand the two device staging arrays back at once
            _lib.check(_lib.load().ng_plan_release_staging(self.plan.handle), "ng_plan_release_staging")
        if _device.is_torch_tensor(f):
            return _device.torch().from_numpy(out)
        return out


def _shape_arg(grid):
    return _lib.as_i64(list(grid.shape))


def _kron_embed(mat_1d, shape, axis_index):
    """I (x) ... (x) D (x) ... (x) I in axis order (the layout of numgrids/diff.py:162-171, :226-239)."""
    if len(shape) == 1:
        return mat_1d
    result = None
    for i, n in enumerate(shape):
        m = mat_1d if i == axis_index else scipy.sparse.identity(n)
        result = m if result is None else scipy.sparse.kron(result, m)
    return result


class GridDiff:
    """Base class: partial derivative of a given order along one grid axis."""

    def __init__(self, grid, order, axis_index):
        self.axis = grid.get_axis(axis_index)
        self.order = order
        self.grid = grid
        self.axis_index = axis_index

    def __call__(self, f):
        return self.operator(f)

    def as_matrix(self):
        raise NotImplementedError

    def as_linear_operator(self):
        """Matrix-free ``scipy.sparse.linalg.LinearOperator`` of shape (grid.size, grid.size): ``matvec`` runs the
        device operator on the flattened vector (no Kronecker matrix is ever built -- what makes the reference's
        ``as_matrix()`` consumers unusable at scale, numgrids/diff.py:90-91, :130-131, :201-202); ``rmatvec`` /
        dense products fall back to the lazily assembled sparse matrix."""
        from scipy.sparse.linalg import LinearOperator
        n = int(np.prod(self.grid.shape))
        shape = tuple(self.grid.shape)

        def matvec(v):
            v = np.asarray(v)
            return np.asarray(self.operator(v.reshape(shape))).reshape(-1)

        def rmatvec(v):
            return self.as_matrix().T @ np.asarray(v).reshape(-1)

        return LinearOperator((n, n), matvec=matvec, rmatvec=rmatvec,
                              dtype=np.float64)


# ---------------------------------------------------------------------------------------------
class _StencilMixin:
    """Shared by FiniteDifferenceDiff and LogDiff: findiff-style three-scheme stencil plan."""

    def _stencil_setup(self, h, x_div):
        self._h = h
        self._schemes = _tables.fd_schemes(self.order, self.acc)
        n = len(self.axis)
        need = _tables.fd_min_points(self.order, self.acc)
        if n < need:
            raise ValueError(f"axis with {n} points is too short for order={self.order}, "
                             f"acc={self.acc} finite differences (needs {need})")
        self._x_div = x_div
        self.operator = DeviceOperator(self.grid.shape, self._make_plan)

    def _make_plan(self, shape=None, x_div=None) -> _lib.Plan:
        """Plan for the grid's shape, or (slab layer) for an explicit local `shape` whose extent
        along this axis may be a slab of the axis -- the tables stay those of the global axis."""
        lib = _lib.load()
        oc, wc = self._schemes["center"]
        of, wf = self._schemes["forward"]
        ob, wb = self._schemes["backward"]
        keep = [_lib.as_f64(_tables.snap_unit_weights(w)) for w in (wc, wf, wb)]
        offs = [_lib.as_i32(o) for o in (oc, of, ob)]
        shp, shp_p = _shape_arg(self.grid) if shape is None else _lib.as_i64(list(shape))
        inv_h_pow = 1.0 / self._h ** self.order
        x_div = self._x_div if x_div is None else x_div
        xd = _lib.as_f64(x_div) if x_div is not None else (None, None)
        out = C.c_void_p()
        _lib.check(lib.ng_fd_plan_create(
            len(shp), shp_p, self.axis_index, len(oc), keep[0][1], offs[0][1], len(of),
            keep[1][1], offs[1][1], keep[2][1], offs[2][1], float(inv_h_pow), xd[1],
            C.byref(out)), "ng_fd_plan_create")
        return _lib.Plan(out.value)

    def _stencil_matrix_1d(self):
        n = len(self.axis)
        nb = len(self._schemes["center"][0]) // 2
        rows, cols, vals = [], [], []
        for i in range(n):
            key = "forward" if i < nb else ("backward" if i >= n - nb else "center")
            offs, w = self._schemes[key]
            for o, wk in zip(offs, w):
                rows.append(i)
                cols.append(i + o)
                vals.append(wk / self._h ** self.order)
        return scipy.sparse.csr_matrix((vals, (rows, cols)), shape=(n, n))


class FiniteDifferenceDiff(_StencilMixin, GridDiff):
    """Finite differences on an equidistant, non-periodic axis (reference numgrids/diff.py:82-91)."""

    def __init__(self, grid, order, axis_index, acc=4):
        super().__init__(grid, order, axis_index)
        if not isinstance(self.axis, EquidistantAxis):
            raise TypeError(f"Axis must be of type EquidistantAxis. Got: {type(self.axis)}")
        self.acc = acc
        self._stencil_setup(self.axis.spacing, None)

    def as_matrix(self):
        return scipy.sparse.csr_matrix(
            _kron_embed(self._stencil_matrix_1d(), self.grid.shape, self.axis_index))


class LogDiff(_StencilMixin, GridDiff):
    """FD on ln x, then ``/ x`` once for any order (reference numgrids/diff.py:252-275)."""

    def __init__(self, grid, order, axis_index, acc=6):
        super().__init__(grid, order, axis_index)
        if not isinstance(self.axis, LogAxis):
            raise TypeError(f"Axis must be of type LogAxis. Got: {type(self.axis)}")
        self.acc = acc
        xi = self.axis.coords_internal
        self._stencil_setup(xi[1] - xi[0], self.axis.coords)

    def as_matrix(self):
        D_log = _kron_embed(self._stencil_matrix_1d(), self.grid.shape, self.axis_index)
        shp = [1] * self.grid.ndims
        shp[self.axis_index] = len(self.axis)
        x = np.broadcast_to(self.axis.coords.reshape(shp), self.grid.shape)
        return scipy.sparse.diags(1.0 / x.ravel()) @ scipy.sparse.csr_matrix(D_log)


# ---------------------------------------------------------------------------------------------
class FFTDiff(GridDiff):
    """Spectral derivative on a periodic equidistant axis (reference numgrids/diff.py:109-171)."""

    PADDED_MIN_POINTS = 65          # non-power-of-two axes from this length on: zero-padded transform
    # Derivative orders the zero-padded transform serves.  Order 2 amplifies ANY transform's rounding by (n/2)^2:
    # measured against numpy.fft (tools/check_fft_order2_nonpow2.py, profiles/r02b_fft_order2_nonpow2.txt) the padded
    # transform and the dense ordered contraction it replaces are equally close (n = 1000: 1.2e-11 / 9.7e-12,
    # n = 1536: 3.1e-11 / 3.6e-11), both ten times closer to numpy.fft than numpy.fft is to a long-double DFT
    # (9e-11 / 3e-10) -- so the O(n log n) path costs no accuracy.  Higher orders keep the dense contraction.
    PADDED_ORDERS = (1, 2)
    PADDED_FAST_MAX = 512           # padded transforms up to 1024 points run on the register kernels (fft2.cu, fft_tile.cu)
    MIXED_MIN_POINTS = 3            # (the native transform is at least as fast as the dense contraction from the shortest axes on)
    NONPOW2 = os.environ.get("NUMGRIDS_B200_FFT_NONPOW2", "auto")       # auto | native | padded

    def __init__(self, grid, order, axis_index, acc=6):
        super().__init__(grid, order, axis_index)
        if not isinstance(self.axis, EquidistantAxis):
            raise TypeError("Spectral FFT differentiation requires equidistant axis.")
        if not self.axis.periodic:
            raise TypeError("Spectral FFT differentiation requires periodic boundary conditions.")
        self.acc = acc
        n = len(self.axis)
        self._W_1d = _tables.fft_multiplier(n, self.axis.spacing, order)
        self._dense = None
        self._D = None
        self.operator = DeviceOperator(grid.shape, self._make_plan, real_result_only=True)

    def _dense_matrix(self):
        if self._dense is None:
            self._dense = _tables.fft_dense_matrix(self._W_1d)
        return self._dense

    @classmethod
    def route(cls, n: int, order: int, prefer_native: bool = False) -> str:
        """Which transform serves a periodic axis of n points: "pow2", "mixed", "padded" or "dense".

        Powers of two (<= 8192): the register / shared-memory FFT kernels.  Any other length (NONPOW2 = "auto"):
          * first and second derivatives with 65 <= n <= 512 (n even, and not just above 128: _padded_is_faster):
            zero-padded power-of-two transform on the two-pass register kernels (O(n log n) like the reference's
            pocketfft, 100-160 Gpts/s; as close to numpy.fft as the native transform: see PADDED_ORDERS);
          * otherwise, n = 2^a 3^b 5^c 7^d 11^e 13^f: the library's native n-point mixed-radix transform (any order;
            60-155 Gpts/s, faster than the padded transform beyond n = 512 and than the dense contraction everywhere);
          * what is left (other prime factors): padded for orders 1-2, else the dense O(n^2) contraction.
        "native" -- and prefer_native, which the curvilinear composites pass (make_composite_plan) -- takes the
        mixed-radix transform wherever it exists, "padded" never does."""
        if _tables.is_pow2(n) and n <= 8192:
            return "pow2"
        mixed = cls.NONPOW2 != "padded" and _tables.fft_mixed_supported(n) and n >= cls.MIXED_MIN_POINTS
        padded = order in cls.PADDED_ORDERS and n >= cls.PADDED_MIN_POINTS
        if (mixed and padded and cls.NONPOW2 == "auto" and not prefer_native
                and n <= cls.PADDED_FAST_MAX and cls._padded_is_faster(n)):
            mixed = False
        return "mixed" if mixed else "padded" if padded else "dense"

    @staticmethod
    def _padded_is_faster(n: int) -> bool:
        """Measured on a B200 (profiles/r02b_fft_mixed.txt): the zero-padded transform runs at about 1.55 n / 0.55 n /
        0.33 n Gpts/s on its 256 / 512 / 1024-point register kernels (n = 65..128 / ..256 / ..512) when n is even -- an odd n
        misses the persistent ring kernel and halves that -- against 85-105 Gpts/s for the native mixed-radix transform."""
        if n % 2:
            return False
        return not 128 < n < 182

    def _make_plan(self, shape=None) -> _lib.Plan:
        return self._plan_for(self._W_1d, self.order, shape, self._dense_matrix)

    def make_composite_plan(self, shape=None) -> _lib.Plan:
        """Plan of this operator as a building block of ``CurvilinearGrid.laplacian / gradient / divergence / curl``.
        On a non-power-of-two length those prefer the native mixed-radix transform to the (2-3 x faster) zero-padded one:
        the composites amplify rounding -- 1/h^2 factors, terms that cancel -- and against the reference the native
        transform is 3-5 x closer (spherical Laplacian of a harmonic field, n_phi = 500: 2.4e-11 vs 1.1e-10,
        gradient 5e-14 vs 1.3e-13; profiles/r02b_fft_mixed.txt)."""
        return self._plan_for(self._W_1d, self.order, shape, self._dense_matrix, prefer_native=True)

    def make_square_plan(self, shape=None) -> _lib.Plan:
        """Plan of this operator applied TWICE in one transform: multiplier W**2 (for a first derivative: -(k s)**2 with
        the Nyquist bin zeroed, exactly what two applications do).  ``CurvilinearGrid.laplacian`` uses it for
        D(c D f) = c D D f on a periodic axis whose coefficient c does not vary along that axis: one transform instead
        of two, and none of the first transform's rounding noise amplified by the second (on zero-padded lengths the
        chained form reaches 6e-11 at n = 360, the single transform 1.4e-12)."""
        W2 = self._W_1d * self._W_1d
        return self._plan_for(W2, 2 * self.order, shape, lambda: _tables.fft_dense_matrix(W2), prefer_native=True)

    def _plan_for(self, W1d, order, shape, dense, prefer_native=False) -> _lib.Plan:
        lib = _lib.load()
        shp, shp_p = _shape_arg(self.grid) if shape is None else _lib.as_i64(list(shape))
        W = np.ascontiguousarray(W1d, dtype=np.complex128).view(np.float64)
        Wk, W_p = _lib.as_f64(W)
        n = len(self.axis)
        out = C.c_void_p()
        route = self.route(n, order, prefer_native)
        pow2, mixed_ok, pad_ok = route == "pow2", route == "mixed", route == "padded"
        if mixed_ok:
            _lib.check(lib.ng_fft_plan_create(len(shp), shp_p, self.axis_index, W_p, None, C.byref(out)),
                       "ng_fft_plan_create")
            return _lib.Plan(out.value)
        padded = _tables.fft_padded_multiplier(W1d) if pad_ok else None
        if padded is not None:
            m, Wm = padded
            Wmk, Wm_p = _lib.as_f64(np.ascontiguousarray(Wm, dtype=np.complex128).view(np.float64))
            _lib.check(lib.ng_fft_plan_create_padded(len(shp), shp_p, self.axis_index, m, Wm_p, C.byref(out)),
                       "ng_fft_plan_create_padded")
            return _lib.Plan(out.value)
        Dk, D_p = (None, None) if pow2 else _lib.as_f64(dense())
        _lib.check(lib.ng_fft_plan_create(len(shp), shp_p, self.axis_index, W_p, D_p, C.byref(out)),
                   "ng_fft_plan_create")
        return _lib.Plan(out.value)

    def as_matrix(self):
        if self._D is None:
            self._D = _kron_embed(scipy.sparse.csc_matrix(self._dense_matrix()),
                                  self.grid.shape, self.axis_index)
        return self._D


# ---------------------------------------------------------------------------------------------
class ChebyshevDiff(GridDiff):
    """Chebyshev spectral derivative (reference numgrids/diff.py:186-239).

    The kernel sums ``y_i = sum_j M_ij f_j`` one j at a time in the direction scipy's matvec uses on
    the reference's Kronecker matrix, without FMA, so results match the reference bit for bit.
    """

    def __init__(self, grid, order, axis_index):
        super().__init__(grid, order, axis_index)
        self._scale = self.axis[-1] - self.axis[0]
        last_nd = grid.ndims > 1 and axis_index == grid.ndims - 1
        self._M = _tables.chebyshev_matrix(self.axis.coords_internal, self.axis.coords, order, last_nd)
        self._descending = _tables.chebyshev_descending(order, axis_index, grid.ndims)
        self._sparse = None
        self.operator = DeviceOperator(grid.shape, self._make_plan)

    def _make_plan(self, shape=None) -> _lib.Plan:
        lib = _lib.load()
        shp, shp_p = _shape_arg(self.grid) if shape is None else _lib.as_i64(list(shape))
        Mk, M_p = _lib.as_f64(self._M)
        out = C.c_void_p()
        _lib.check(lib.ng_cheb_plan_create(len(shp), shp_p, self.axis_index, M_p,
                                           int(self._descending), C.byref(out)), "ng_cheb_plan_create")
        return _lib.Plan(out.value)

    def as_matrix(self):
        if self._sparse is None:
            self._sparse = _kron_embed(scipy.sparse.csc_matrix(self._M), self.grid.shape,
                                       self.axis_index)
        return self._sparse
