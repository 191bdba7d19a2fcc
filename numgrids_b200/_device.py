"""Array adapters between user containers and raw device pointers.

Accepted inputs (SURVEY.md 8b "Array conventions"): ``numpy.ndarray`` / array-likes (host; staged
H2D/D2H by the C library's ``*_host`` entry points or through pinned torch buffers), CUDA
``torch.Tensor`` and any object exposing ``__cuda_array_interface__`` (wrapped zero-copy by
torch).  All arrays are float64 and C-contiguous; outputs are freshly allocated by this layer --
the C library never allocates user-visible memory.  PyTorch is used only as the allocator and
stream provider.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def is_torch_tensor(x) -> bool:
    return _torch is not None and isinstance(x, _torch.Tensor) or (
        type(x).__module__.startswith("torch") and hasattr(x, "data_ptr"))


def is_device_array(x) -> bool:
    if is_torch_tensor(x):
        return bool(x.is_cuda)
    return hasattr(x, "__cuda_array_interface__")


def is_complex(x) -> bool:
    """Complex-valued input (numpy / torch / array-like)?  The kernels are real; complex fields are split
    into real and imaginary parts by the callers (`split_complex`), never silently truncated."""
    if is_torch_tensor(x):
        return bool(x.is_complex())
    dt = getattr(x, "dtype", None)
    if dt is not None:
        try:
            return np.issubdtype(np.dtype(dt), np.complexfloating)
        except TypeError:
            return False
    if isinstance(x, (list, tuple)):
        return np.iscomplexobj(np.asarray(x))
    return isinstance(x, complex)


def split_complex(x):
    """(real part, imaginary part) of a complex array in its own container (host ndarray or CUDA tensor)."""
    if is_torch_tensor(x):
        return x.real.contiguous(), x.imag.contiguous()
    a = np.asarray(x)
    return np.ascontiguousarray(a.real, dtype=np.float64), np.ascontiguousarray(a.imag, dtype=np.float64)


def join_complex(re, im):
    """re + 1j * im in the parts' container."""
    if is_torch_tensor(re):
        return torch().complex(re, im)
    return np.asarray(re) + 1j * np.asarray(im)


def _reject_complex(x):
    if is_complex(x):
        raise TypeError("complex-valued arrays are not accepted here: apply the (real, linear) operator to the "
                        "real and imaginary parts separately")


class device_of:
    """Context manager: make the CUDA device that holds `x` current for the calls inside (plans, streams and
    kernel attributes belong to one device); a no-op for host arrays."""

    def __init__(self, x):
        self.ctx = None
        if is_torch_tensor(x) and x.is_cuda:
            self.ctx = torch().cuda.device(x.device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def as_device_tensor(x, shape=None):
    """Zero-copy float64 contiguous CUDA tensor view of a device array (copies only if needed)."""
    t = torch()
    _reject_complex(x)
    if not is_torch_tensor(x):
        x = t.as_tensor(x, device="cuda")
    if x.dtype != t.float64:
        x = x.to(t.float64)
    if not x.is_contiguous():
        x = x.contiguous()
    if shape is not None and tuple(x.shape) != tuple(shape):
        if x.numel() != int(np.prod(shape)):
            raise ValueError(f"array of shape {tuple(x.shape)} does not match grid shape {tuple(shape)}")
        x = x.reshape(tuple(shape))
    return x


def as_host_array(x, shape=None):
    _reject_complex(x)
    if is_torch_tensor(x):
        x = x.detach().cpu().numpy()
    a = np.ascontiguousarray(x, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        if a.size != int(np.prod(shape)):
            raise ValueError(f"array of shape {a.shape} does not match grid shape {tuple(shape)}")
        a = a.reshape(tuple(shape))
    return a


def current_stream_ptr() -> int:
    return int(torch().cuda.current_stream().cuda_stream)


def to_device(a: np.ndarray):
    """Host ndarray -> CUDA tensor (H2D on the current stream)."""
    _lib.require_gpu()
    t = torch()
    return t.from_numpy(a).to("cuda", non_blocking=False)


def empty_like_device(x):
    return torch().empty_like(x)
