"""ctypes binding of libnumgrids_b200.so (C ABI declared in include/numgrids_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is usable, every
compute call raises ``NumgridsB200Error``.  Table construction and validation (host logic) work
without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

NG_MAX_DIM = 4
NG_MAX_TAPS = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NUMGRIDS_B200_LIB", os.path.join(_HERE, "lib", "libnumgrids_b200.so"))


class NumgridsB200Error(RuntimeError):
    """Raised when the CUDA library is missing or a library call fails."""


class NgCoef(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride", C.c_int64 * NG_MAX_DIM)]


class NgFuse(C.Structure):
    _fields_ = [("pre", NgCoef), ("minuend", C.c_void_p), ("post", NgCoef), ("addend", C.c_void_p),
                ("add_zero", C.c_int32), ("divisor", NgCoef), ("finite", C.c_int32)]


class NgSlabComm(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("my_halo", (C.c_void_p * 2) * 2),
                ("peer_lo_halo", C.c_void_p * 2), ("peer_hi_halo", C.c_void_p * 2), ("my_flags", C.c_void_p),
                ("peer_lo_flag", C.c_void_p), ("peer_hi_flag", C.c_void_p)]


class NgCopyBlock(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int64), ("row_len", C.c_int64),
                ("src_stride", C.c_int64), ("dst_stride", C.c_int64), ("local_offset", C.c_int64)]


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_I32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/numgrids_b200.h declares
SIGNATURES = {
    "ng_abi_version": (C.c_int, []),
    "ng_last_error": (C.c_char_p, []),
    "ng_device_count": (C.c_int, []),
    "ng_launch_count": (C.c_int64, []),
    "ng_last_kernel": (C.c_char_p, []),
    "ng_tuning_reload": (None, []),
    "ng_fd_plan_create": (C.c_int, [C.c_int, _I64, C.c_int, C.c_int, _D, _I32, C.c_int, _D, _I32, _D,
                                    _I32, C.c_double, _D, _PP]),
    "ng_cheb_plan_create": (C.c_int, [C.c_int, _I64, C.c_int, _D, C.c_int, _PP]),
    "ng_fft_plan_create": (C.c_int, [C.c_int, _I64, C.c_int, _D, _D, _PP]),
    "ng_fft_plan_create_padded": (C.c_int, [C.c_int, _I64, C.c_int, C.c_int64, _D, _PP]),
    "ng_fft_mixed_radices": (C.c_int, [C.c_int64, _I32, _I32]),
    "ng_apply": (C.c_int, [_P, _P, _P, _P]),
    "ng_apply_fused": (C.c_int, [_P, _P, _P, C.POINTER(NgFuse), _P]),
    "ng_apply_batch": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "ng_apply_host": (C.c_int, [_P, _P, _P]),
    "ng_plan_release_staging": (C.c_int, [_P]),
    "ng_fd_apply_slab": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "ng_fd_pack_halo": (C.c_int, [_P, _P, _P, _P, _P]),
    "ng_plan_halo_width": (C.c_int, [_P]),
    "ng_apply_fused_slab": (C.c_int, [_P, _P, _P, C.POINTER(NgFuse), _P, _P, _P]),
    "ng_fd_pack_halo_fused": (C.c_int, [_P, _P, C.POINTER(NgCoef), _P, _P, _P]),
    "ng_slab_block_copy": (C.c_int, [C.c_int, C.POINTER(NgCopyBlock), C.c_int, _I64, C.POINTER(NgFuse), C.c_int, _P]),
    "ng_fdlap_plan_create": (C.c_int, [C.c_int, _I64, _PP, _PP]),
    "ng_fdlap_apply": (C.c_int, [_P, _P, _P, _P]),
    "ng_fdlap_apply_slab": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "ng_fdlap_apply_slab_range": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P]),
    "ng_fdlap_apply_slab_part": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P]),
    "ng_fdlap_apply_host": (C.c_int, [_P, _P, _P]),
    "ng_fdlap_pack_halo_range": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, _P]),
    "ng_slab_plan_create": (C.c_int, [C.POINTER(NgSlabComm), C.c_int, _I64, _PP, _PP]),
    "ng_slab_lap_apply": (C.c_int, [_P, _P, _P, _P]),
    "ng_slab_timed_out": (C.c_int, [_P]),
    "ng_curv_plan_create": (C.c_int, [C.c_int, _I64, _PP, _PP]),
    "ng_curv_set_coef": (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int64, _I64]),
    "ng_curv_set_axis_square": (C.c_int, [_P, C.c_int, _P]),
    "ng_curv_laplacian": (C.c_int, [_P, _P, _P, _P]),
    "ng_curv_gradient": (C.c_int, [_P, _P, _PP, _P]),
    "ng_curv_divergence": (C.c_int, [_P, _PP, _P, _P]),
    "ng_curv_curl": (C.c_int, [_P, _PP, _PP, _P]),
    "ng_bc_apply": (C.c_int, [C.c_int, _I64, C.c_int, C.c_int, C.c_int, C.c_double, _P, C.c_double, _P, _P]),
    "ng_quad_workspace_doubles": (C.c_int64, []),
    "ng_quad_apply": (C.c_int, [C.c_int, _I64, _PP, _P, _P, _P, _P]),
    "ng_diff_norm": (C.c_int, [C.c_int64, _P, _P, C.c_int, _P, _P, _P]),
    "ng_interp_linear": (C.c_int, [C.c_int, _I64, _I64, C.c_int, _PP, _PP, _PP, _P, _P, _P]),
    "ng_interp_bspline": (C.c_int, [C.c_int, _I64, _I64, C.c_int, _I32, _PP, _PP, _P, _P, _P]),
    "ng_plan_destroy": (C.c_int, [_P]),
    "ng_plan_device_bytes": (C.c_int64, [_P]),
}

NG_COEF_J, NG_COEF_LAP, NG_COEF_DIV, NG_COEF_GRAD, NG_COEF_H, NG_COEF_CURL = range(6)

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and bind every entry point; fail loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NumgridsB200Error(
            f"{LIB_PATH} not found: build it with `python -m numgrids_b200.build` "
            "(numgrids_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here means header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.ng_abi_version() != 1:
        raise NumgridsB200Error(f"ABI version mismatch: library {lib.ng_abi_version()}, binding 1")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ng_last_error().decode(errors="replace")
        raise NumgridsB200Error(f"{what or 'libnumgrids_b200'} failed (code {rc}): {msg}")


def device_count() -> int:
    return int(load().ng_device_count())


def require_gpu() -> None:
    if device_count() < 1:
        raise NumgridsB200Error("no usable CUDA device: numgrids_b200 has no CPU fallback")


def launch_count() -> int:
    return int(load().ng_launch_count())


def last_kernel() -> str:
    """Name of the kernel instance this thread launched last (tests pin kernel selection with it)."""
    return load().ng_last_kernel().decode()


def tuning_reload() -> None:
    """Make the library re-read its NG_* tuning switches from the environment."""
    load().ng_tuning_reload()


# ---------------------------------------------------------------- small ctypes helpers
def as_i64(seq):
    a = np.ascontiguousarray(seq, dtype=np.int64)
    return a, a.ctypes.data_as(_I64)


def as_i32(seq):
    a = np.ascontiguousarray(seq, dtype=np.int32)
    return a, a.ctypes.data_as(_I32)


def as_f64(seq):
    a = np.ascontiguousarray(seq, dtype=np.float64)
    return a, a.ctypes.data_as(_D)


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
    return arr


class Plan:
    """Owns one ``ng_plan*``; destroyed with the Python object."""

    def __init__(self, handle: int, keepalive=()) -> None:
        self.handle = C.c_void_p(handle)
        self._keepalive = tuple(keepalive)   # child plans referenced by the C object

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.ng_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def device_bytes(self) -> int:
        return int(load().ng_plan_device_bytes(self.handle))

    @property
    def halo_width(self) -> int:
        return int(load().ng_plan_halo_width(self.handle))
