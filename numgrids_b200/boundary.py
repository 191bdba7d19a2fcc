"""Boundary conditions -- the API surface of the reference's numgrids/boundary.py.

``BoundaryFace``, ``DirichletBC``, ``NeumannBC``, ``RobinBC`` and ``apply_bcs`` keep the reference's
names, arguments, return values and exceptions.  What changes is where the work runs:

* ``apply(u)`` (numgrids/boundary.py:157-166, :216-245) -- the step that follows a derivative apply in
  a time-stepping loop -- runs ``ng_bc_apply`` (csrc/bc.cu) in place on the array.  A device array
  (CUDA ``torch.Tensor`` / ``__cuda_array_interface__``) never leaves the device; for a host
  ``ndarray`` only the face and its adjacent layer travel.  No CPU fallback.
* ``apply_to_system(L, rhs)`` (:168-187, :247-261, :289-304) is sparse-matrix assembly for a host
  solver, not data-parallel work: it stays on the host, but replaces all face rows in one sparse
  operation instead of the reference's per-row LIL loop (same entries, same ``csc`` result).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse

from . import _device, _lib

_KIND_DIRICHLET, _KIND_NEUMANN = 0, 1


class BoundaryFace:
    """One face of a grid boundary: ``side`` "low" (index 0) or "high" (last index) of a non-periodic axis."""

    def __init__(self, grid, axis_index: int, side: str) -> None:
        if side not in ("low", "high"):
            raise ValueError(f"side must be 'low' or 'high', got {side!r}")
        if axis_index < 0 or axis_index >= grid.ndims:
            raise ValueError(
                f"axis_index {axis_index} out of range for {grid.ndims}-dimensional grid"
            )
        if grid.get_axis(axis_index).periodic:
            raise ValueError(
                f"Axis {axis_index} is periodic and has no boundary faces"
            )
        self.grid = grid
        self.axis_index = axis_index
        self.side = side

    def _index(self, adjacent: bool = False):
        idx = [slice(None)] * self.grid.ndims
        if self.side == "low":
            idx[self.axis_index] = 1 if adjacent else 0
        else:
            idx[self.axis_index] = -2 if adjacent else -1
        return tuple(idx)

    @property
    def shape(self):
        """Shape of the face: the grid shape with this face's axis removed."""
        s = self.grid.shape
        return s[:self.axis_index] + s[self.axis_index + 1:]

    @property
    def mask(self):
        """Boolean mask of shape ``grid.shape``, True on this face."""
        m = np.zeros(self.grid.shape, dtype=bool)
        m[self._index()] = True
        return m

    @property
    def flat_indices(self):
        """Flat indices of this face's points in the ravelled grid (ascending, C order)."""
        shape = self.grid.shape
        ax = self.axis_index
        inner = int(np.prod(shape[ax + 1:], dtype=np.int64))
        outer = int(np.prod(shape[:ax], dtype=np.int64))
        i_face = 0 if self.side == "low" else shape[ax] - 1
        o = np.arange(outer, dtype=np.int64)[:, None]
        c = np.arange(inner, dtype=np.int64)[None, :]
        return ((o * shape[ax] + i_face) * inner + c).ravel()

    @property
    def normal_sign(self) -> int:
        """-1 on the "low" face, +1 on the "high" face (outward normal along the axis)."""
        return -1 if self.side == "low" else 1

    def __repr__(self) -> str:
        return f"BoundaryFace(axis={self.axis_index}, side={self.side!r})"


class BoundaryCondition:
    """Base class; ``value`` is a scalar, an array of face values, or a callable of the face coordinates."""

    def __init__(self, face: BoundaryFace, value=0.0) -> None:
        self.face = face
        self._value_spec = value

    def _resolve_value(self):
        """1-D array of values, one per face point (numgrids/boundary.py:117-132)."""
        face = self.face
        grid = face.grid
        if callable(self._value_spec):
            # coordinates restricted to the face, without building the full meshgrid
            coords = [np.asarray(a.coords) for a in grid.axes]
            coords[face.axis_index] = coords[face.axis_index][[0 if face.side == "low" else -1]]
            mesh = np.meshgrid(*coords, indexing="ij")
            sel = [slice(None)] * grid.ndims
            sel[face.axis_index] = 0
            face_coords = tuple(m[tuple(sel)] for m in mesh)
            return np.asarray(self._value_spec(face_coords)).ravel()
        if isinstance(self._value_spec, np.ndarray):
            return self._value_spec.ravel()
        if _device.is_device_array(self._value_spec):
            return _device.as_host_array(self._value_spec).ravel()
        return np.full(int(np.prod(face.shape, dtype=np.int64)), float(self._value_spec))

    # -- device side -----------------------------------------------------
    def _device_value(self):
        """(device pointer or None, scalar, keep-alive) for the face values."""
        spec = self._value_spec
        n_face = int(np.prod(self.face.shape, dtype=np.int64))
        if not callable(spec) and not isinstance(spec, np.ndarray) and not _device.is_device_array(spec):
            return None, float(spec), None
        if _device.is_device_array(spec):
            g = _device.as_device_tensor(spec).reshape(-1)
        else:
            g = _device.to_device(np.ascontiguousarray(self._resolve_value(), dtype=np.float64))
        if g.numel() == 1:                       # `u[mask] = g` broadcasts a one-element value (numgrids/boundary.py:165)
            return None, float(g.reshape(-1)[0].item()), None
        if g.numel() != n_face:
            raise ValueError(f"boundary value has {g.numel()} entries, the face has {n_face} points")
        return g.data_ptr(), 0.0, g

    def _apply_device(self, u, kind: int, hs: float):
        face = self.face
        shape = face.grid.shape
        lib = _lib.load()
        _lib.require_gpu()
        side = 0 if face.side == "low" else 1
        g_ptr, g_scalar, keep = self._device_value()
        if _device.is_device_array(u):
            if _device.is_torch_tensor(u):
                if not u.is_contiguous() or u.dtype != _device.torch().float64 or tuple(u.shape) != tuple(shape):
                    raise ValueError("in-place boundary apply needs a contiguous float64 array of the grid's shape")
                t = u
            else:
                t = _device.as_device_tensor(u, shape)   # zero-copy view of a __cuda_array_interface__ object
            shp, shp_p = _lib.as_i64(list(shape))
            _lib.check(lib.ng_bc_apply(len(shape), shp_p, face.axis_index, side, kind, hs, g_ptr, g_scalar,
                                       t.data_ptr(), _device.current_stream_ptr()), "ng_bc_apply")
            return u
        # host array: only the face and its adjacent layer travel (a 2-layer array along the axis).  Like the
        # reference's `u[mask] = g` / slice assignment (numgrids/boundary.py:165, :232-244), any ndarray of the
        # grid's shape is accepted -- other float dtypes and non-contiguous views go through a float64 copy of the
        # two layers and NumPy casts the face back on assignment; complex arrays layer by real / imaginary part
        if not isinstance(u, np.ndarray) or u.shape != tuple(shape):
            raise ValueError("in-place boundary apply needs an ndarray of the grid's shape")
        if np.iscomplexobj(u):
            if kind == _KIND_DIRICHLET:
                re = np.ascontiguousarray(u.real, dtype=np.float64)
                self._apply_device(re, kind, hs)
                u[face._index()] = re[face._index()]            # Dirichlet: u = g (real g; imaginary part 0)
                return u
            raise TypeError("Neumann / Robin conditions on complex host arrays are not supported")
        lo, hi = (face._index(), face._index(True)) if side == 0 else (face._index(True), face._index())
        two = np.stack((u[lo], u[hi]), axis=face.axis_index).astype(np.float64, copy=False)
        d_two = _device.to_device(np.ascontiguousarray(two))
        shp, shp_p = _lib.as_i64(list(two.shape))
        _lib.check(lib.ng_bc_apply(two.ndim, shp_p, face.axis_index, side, kind, hs, g_ptr, g_scalar,
                                   d_two.data_ptr(), _device.current_stream_ptr()), "ng_bc_apply")
        sel = [slice(None)] * two.ndim
        sel[face.axis_index] = side
        u[face._index()] = d_two[tuple(sel)].cpu().numpy()
        return u

    # Subclasses override these:
    def apply(self, u):
        raise NotImplementedError

    def apply_to_system(self, L, rhs):
        raise NotImplementedError

    def _replace_rows(self, L, rhs, new_rows):
        """Rows ``face.flat_indices`` of L := new_rows (sparse, one row per face point), rhs := g."""
        rows = self.face.flat_indices
        g = self._resolve_value()
        L = scipy.sparse.csr_matrix(L, copy=True)
        n = L.shape[0]
        row_of = np.repeat(np.arange(n), np.diff(L.indptr))
        hit = np.zeros(n, dtype=bool)
        hit[rows] = True
        L.data[hit[row_of]] = 0.0
        L.eliminate_zeros()
        scatter = scipy.sparse.csr_matrix((np.ones(len(rows)), (rows, np.arange(len(rows)))),
                                          shape=(n, len(rows)))
        L = (L + scatter @ scipy.sparse.csr_matrix(new_rows)).tocsc()
        rhs = rhs.copy()
        rhs[rows] = g
        return L, rhs


class DirichletBC(BoundaryCondition):
    """Dirichlet condition ``u = g`` on a boundary face."""

    def apply(self, u):
        """Set ``u`` to ``g`` on the face, in place; returns ``u`` (numgrids/boundary.py:157-166)."""
        return self._apply_device(u, _KIND_DIRICHLET, 0.0)

    def apply_to_system(self, L, rhs):
        """Face rows of ``L`` become identity rows, ``rhs = g`` there (numgrids/boundary.py:168-187)."""
        rows = self.face.flat_indices
        n = L.shape[0]
        ident = scipy.sparse.csr_matrix((np.ones(len(rows)), (np.arange(len(rows)), rows)), shape=(len(rows), n))
        return self._replace_rows(L, rhs, ident)


class NeumannBC(BoundaryCondition):
    """Neumann condition ``du/dn = g``; the outward normal derivative is ``normal_sign * du/d(axis)``."""

    def __init__(self, face: BoundaryFace, value=0.0) -> None:
        super().__init__(face, value)
        from .api import Diff
        self._diff = Diff(face.grid, order=1, axis_index=face.axis_index)

    def apply(self, u):
        """First-order, in place: ``u[face] = u[adjacent] + h * normal_sign * g`` (numgrids/boundary.py:216-245)."""
        coords = self.face.grid.get_axis(self.face.axis_index).coords
        h = coords[1] - coords[0] if self.face.side == "low" else coords[-1] - coords[-2]
        return self._apply_device(u, _KIND_NEUMANN, float(h * self.face.normal_sign))

    def apply_to_system(self, L, rhs):
        """Face rows become ``normal_sign * D`` rows of the first-derivative matrix (numgrids/boundary.py:247-261)."""
        D = scipy.sparse.csr_matrix(self._diff.as_matrix())
        return self._replace_rows(L, rhs, self.face.normal_sign * D[self.face.flat_indices, :])


class RobinBC(BoundaryCondition):
    """Robin condition ``a u + b du/dn = g``."""

    def __init__(self, face: BoundaryFace, a: float, b: float, value=0.0) -> None:
        super().__init__(face, value)
        self.a = a
        self.b = b
        from .api import Diff
        self._diff = Diff(face.grid, order=1, axis_index=face.axis_index)

    def apply(self, u):
        """Not supported for Robin conditions; use :meth:`apply_to_system` (numgrids/boundary.py:283-287)."""
        raise NotImplementedError(
            "Function-level apply is not supported for Robin BCs. "
            "Use apply_to_system() for sparse linear system modification."
        )

    def apply_to_system(self, L, rhs):
        """Face rows become ``a * I + b * normal_sign * D`` (numgrids/boundary.py:289-304)."""
        rows = self.face.flat_indices
        n = L.shape[0]
        D = scipy.sparse.csr_matrix(self._diff.as_matrix())
        ident = scipy.sparse.csr_matrix((np.ones(len(rows)), (np.arange(len(rows)), rows)), shape=(len(rows), n))
        return self._replace_rows(L, rhs, self.a * ident + self.b * self.face.normal_sign * D[rows, :])


def apply_bcs(bcs, L, rhs):
    """Apply boundary conditions to a linear system in list order; on shared points the last one wins."""
    for bc in bcs:
        L, rhs = bc.apply_to_system(L, rhs)
    return L, rhs
