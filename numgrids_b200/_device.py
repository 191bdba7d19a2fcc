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


def as_device_tensor(x, shape=None):
    """Zero-copy float64 contiguous CUDA tensor view of a device array (copies only if needed)."""
    t = torch()
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
