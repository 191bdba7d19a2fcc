"""Orthogonal curvilinear grids -- drop-in for the reference's numgrids/curvilinear.py.

``gradient / divergence / laplacian / curl`` keep the reference's signatures, lazy caches
(``_diff_ops``, ``_h_arrays``, ``_jacobian``: None until first use, identity-stable afterwards,
reference numgrids/curvilinear.py:120-149) and error behaviour, and dispatch to the fused
``ng_curv_*`` entry points of libnumgrids_b200.so.

Host side (once per grid): the user's scale-factor callables are evaluated on the meshed
coordinates exactly like the reference (curvilinear.py:126-140), the per-operator coefficient
arrays are formed with the reference's NumPy expressions (``J / h[i] ** 2``, ``J / h[i]``,
``1.0 / h[i]``, ``1.0 / (h[j] * h[k])``) and then reduced to broadcast-compact tables (a
spherical ``J/h_r**2 = r^2 sin(theta)`` becomes an ``[nr][ntheta]`` table) before upload.
Device side: per-axis applies with the elementwise work fused in (csrc/curv.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _device, _lib, _tables
from .grids import Grid


class CurvilinearGrid(Grid):
    """Grid plus one scale-factor callable ``h_i(meshed_coords) -> array`` per axis."""

    def __init__(self, *axes, scale_factors):
        super().__init__(*axes)
        if len(scale_factors) != self.ndims:
            raise ValueError(f"Expected {self.ndims} scale factors, got {len(scale_factors)}.")
        self._scale_factor_fns = tuple(scale_factors)
        self._diff_ops = None
        self._h_arrays = None
        self._jacobian = None
        self._curv_plan = None
        self._coef_done = set()

    # ------------------------------------------------------------ lazy host state
    def _evaluate_scale_factors(self):
        if self._h_arrays is None:
            mesh = self.meshed_coords
            self._h_arrays = tuple(fn(mesh) for fn in self._scale_factor_fns)
        return self._h_arrays

    def _evaluate_jacobian(self):
        if self._jacobian is None:
            J = np.ones(self.shape)
            for hi in self._evaluate_scale_factors():
                J = J * hi
            self._jacobian = J
        return self._jacobian

    def _ensure_diff_ops(self):
        if self._diff_ops is not None:
            return
        from .api import Diff
        self._diff_ops = {i: Diff(self, 1, i) for i in range(self.ndims)}

    # ------------------------------------------------------------ device plan
    def _plan(self) -> _lib.Plan:
        if self._curv_plan is None:
            self._ensure_diff_ops()
            lib = _lib.load()
            _lib.require_gpu()
            from .diff import FFTDiff
            ops = [self._diff_ops[i].operator for i in range(self.ndims)]
            # (periodic axes: the composites' own plan -- native-length transform where the plain Diff would zero-pad)
            children = [op.make_composite_plan() if isinstance(op, FFTDiff) else op.operator.plan for op in ops]
            arr = _lib.ptr_array([p.handle.value for p in children])
            shp, shp_p = _lib.as_i64(list(self.shape))
            out = C.c_void_p()
            _lib.check(lib.ng_curv_plan_create(self.ndims, shp_p, arr, C.byref(out)),
                       "ng_curv_plan_create")
            keep = list(children)
            for i in range(self.ndims):             # periodic axes: D applied twice as ONE transform (see make_square_plan)
                op = self._diff_ops[i].operator
                if isinstance(op, FFTDiff):
                    sq = op.make_square_plan()
                    _lib.check(lib.ng_curv_set_axis_square(out, i, sq.handle), "ng_curv_set_axis_square")
                    keep.append(sq)
            self._curv_plan = _lib.Plan(out.value, keepalive=keep)
        return self._curv_plan

    def _register(self, kind, index, make):
        """Evaluate a coefficient array on the host (reference expression) and upload it once."""
        key = (kind, index)
        if key in self._coef_done:
            return
        with np.errstate(divide="ignore", invalid="ignore"):
            full = make()
        values, strides = _tables.compact_coefficient(full, self.shape)
        v, v_p = _lib.as_f64(values)
        s, s_p = _lib.as_i64(strides)
        _lib.check(_lib.load().ng_curv_set_coef(self._plan().handle, kind, index, v_p, len(v), s_p),
                   "ng_curv_set_coef")
        self._coef_done.add(key)

    # ------------------------------------------------------------ array plumbing
    def _stage(self, arrays):
        """Return (device tensors, host_output?)."""
        on_device = [_device.is_device_array(a) for a in arrays]
        if all(on_device):
            xs = [_device.as_device_tensor(a, self.shape) for a in arrays]
            if len({x.device for x in xs}) > 1:
                raise ValueError("all arrays of one call must live on the same CUDA device")
            return xs, False
        _lib.require_gpu()
        out = []
        for a, dev in zip(arrays, on_device):
            out.append(_device.as_device_tensor(a, self.shape) if dev
                       else _device.to_device(_device.as_host_array(a, self.shape)))
        return out, True

    @staticmethod
    def _finish(t, to_host):
        return t.cpu().numpy() if to_host else t

    def _complex_split(self, arrays, run):
        """Complex fields: the operators are real and linear, so they are applied to the real and the
        imaginary parts separately (the reference's NumPy arithmetic does the same componentwise).  On grids
        with a periodic axis the reference's FFTDiff keeps only the real part of that axis' terms
        (numgrids/diff.py:141) -- a mixed result that is refused here rather than reproduced."""
        if not any(_device.is_complex(a) for a in arrays):
            return None
        from .axes import EquidistantAxis
        if any(isinstance(a, EquidistantAxis) and a.periodic for a in self.axes):
            raise TypeError("complex fields are not supported on curvilinear grids with a periodic axis "
                            "(the reference's FFTDiff discards the imaginary part there)")
        parts = [_device.split_complex(a) if _device.is_complex(a) else (a, None) for a in arrays]
        re = run(*[p[0] for p in parts])
        zero = lambda a: a * 0.0
        im = run(*[p[1] if p[1] is not None else zero(p[0]) for p in parts])
        if isinstance(re, tuple):
            return tuple(_device.join_complex(a, b) for a, b in zip(re, im))
        return _device.join_complex(re, im)

    # ------------------------------------------------------------ operators
    def gradient(self, f):
        """(grad f)_i = (1/h_i) d f / d q_i, non-finite values replaced by 0 (curvilinear.py:153-177)."""
        c = self._complex_split([f], self.gradient)
        if c is not None:
            return c
        self._ensure_diff_ops()
        h = self._evaluate_scale_factors()
        (x,), to_host = self._stage([f])
        with _device.device_of(x):
            plan = self._plan()
            for i in range(self.ndims):
                self._register(_lib.NG_COEF_GRAD, i, lambda i=i: 1.0 / h[i])
            outs = [_device.empty_like_device(x) for _ in range(self.ndims)]
            ptrs = _lib.ptr_array([o.data_ptr() for o in outs])
            _lib.check(_lib.load().ng_curv_gradient(plan.handle, x.data_ptr(), ptrs,
                                                    _device.current_stream_ptr()), "ng_curv_gradient")
        return tuple(self._finish(o, to_host) for o in outs)

    def divergence(self, *v):
        """(1/J) sum_i d/dq_i (J/h_i v_i) (curvilinear.py:179-215)."""
        if len(v) == self.ndims:
            c = self._complex_split(list(v), self.divergence)
            if c is not None:
                return c
        if len(v) != self.ndims:
            raise ValueError(f"Expected {self.ndims} vector components, got {len(v)}.")
        self._ensure_diff_ops()
        h = self._evaluate_scale_factors()
        J = self._evaluate_jacobian()
        xs, to_host = self._stage(list(v))
        with _device.device_of(xs[0]):
            plan = self._plan()
            self._register(_lib.NG_COEF_J, 0, lambda: J)
            for i in range(self.ndims):
                self._register(_lib.NG_COEF_DIV, i, lambda i=i: J / h[i])
            out = _device.empty_like_device(xs[0])
            ptrs = _lib.ptr_array([x.data_ptr() for x in xs])
            _lib.check(_lib.load().ng_curv_divergence(plan.handle, ptrs, out.data_ptr(),
                                                      _device.current_stream_ptr()), "ng_curv_divergence")
        return self._finish(out, to_host)

    def laplacian(self, f):
        """(1/J) sum_i d/dq_i (J/h_i^2 d f/d q_i): a DOUBLE first-derivative apply per axis
        (curvilinear.py:217-244), not a sum of second-derivative operators."""
        c = self._complex_split([f], self.laplacian)
        if c is not None:
            return c
        self._ensure_diff_ops()
        h = self._evaluate_scale_factors()
        J = self._evaluate_jacobian()
        (x,), to_host = self._stage([f])
        with _device.device_of(x):
            plan = self._plan()
            self._register(_lib.NG_COEF_J, 0, lambda: J)
            for i in range(self.ndims):
                self._register(_lib.NG_COEF_LAP, i, lambda i=i: J / h[i] ** 2)
            out = _device.empty_like_device(x)
            _lib.check(_lib.load().ng_curv_laplacian(plan.handle, x.data_ptr(), out.data_ptr(),
                                                     _device.current_stream_ptr()), "ng_curv_laplacian")
        return self._finish(out, to_host)

    def curl(self, *v):
        """3-D: tuple of three components; 2-D: the scalar out-of-plane component
        (curvilinear.py:246-317)."""
        if self.ndims == 2:
            return self._curl_2d(*v)
        if self.ndims == 3:
            return self._curl_3d(*v)
        raise ValueError("Curl is only defined for 2D and 3D grids.")

    def _curl_common(self, v, pairs):
        c = self._complex_split(list(v), lambda *w: tuple(self._curl_common(w, pairs)))
        if c is not None:
            return list(c)
        self._ensure_diff_ops()
        h = self._evaluate_scale_factors()
        xs, to_host = self._stage(list(v))
        with _device.device_of(xs[0]):
            plan = self._plan()
            for i in range(self.ndims):
                self._register(_lib.NG_COEF_H, i, lambda i=i: h[i])
            for c, (j, k) in enumerate(pairs):
                self._register(_lib.NG_COEF_CURL, c, lambda j=j, k=k: 1.0 / (h[j] * h[k]))
            outs = [_device.empty_like_device(xs[0]) for _ in pairs]
            vin = _lib.ptr_array([x.data_ptr() for x in xs])
            vout = _lib.ptr_array([o.data_ptr() for o in outs])
            _lib.check(_lib.load().ng_curv_curl(plan.handle, vin, vout, _device.current_stream_ptr()),
                       "ng_curv_curl")
        return [self._finish(o, to_host) for o in outs]

    def _curl_2d(self, v0, v1):
        return self._curl_common((v0, v1), [(0, 1)])[0]

    def _curl_3d(self, v0, v1, v2):
        return tuple(self._curl_common((v0, v1, v2), [(1, 2), (2, 0), (0, 1)]))
