"""Argument plumbing for ``evaluate``: turn the caller's arrays into raw
addresses for the C ABI without copying.

Accepted array kinds (all must be float64, C-contiguous, 1-D or reshapeable):
  * numpy.ndarray                      -> HOST path (fcx_*_evaluate_host)
  * torch.Tensor on a CUDA device      -> DEVICE path, current torch stream
  * any object with ``__dlpack__`` or ``__cuda_array_interface__`` living on a
    CUDA device (cupy, numba, jax, a dolfinx-on-GPU vector ...) -> DEVICE path,
    zero-copy through DLPack.
PyTorch is used only as the owner of device memory and streams.
"""
from __future__ import annotations

import numpy as np

HOST = "host"
DEVICE = "device"


def _torch():
    import torch

    return torch


class Buf:
    __slots__ = ("kind", "ptr", "size", "owner", "device_index")

    def __init__(self, kind, ptr, size, owner, device_index=None):
        self.kind = kind
        self.ptr = ptr
        self.size = size
        self.owner = owner  # keeps the memory alive for the duration of the call
        self.device_index = device_index


class NullBuf(Buf):
    """An absent optional array (``tangent=None``): a NULL address on the side of ``like``."""

    __slots__ = ()

    def __init__(self, like: Buf):
        super().__init__(like.kind, None, 0, None, like.device_index)


def as_buf(a, name: str, writable: bool = False, dtype=np.float64) -> Buf:
    """Classify one array argument and return its address and element count."""
    if isinstance(a, np.ndarray):
        if a.dtype != dtype:
            raise TypeError(f"{name}: expected {np.dtype(dtype)}, got {a.dtype}")
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError(f"{name}: array must be C-contiguous (it is updated in place)")
        if writable and not a.flags["WRITEABLE"]:
            raise ValueError(f"{name}: array is read-only but must be written")
        return Buf(HOST, a.ctypes.data, a.size, a)
    torch = _torch()
    t = None
    if isinstance(a, torch.Tensor):
        t = a
    elif hasattr(a, "__cuda_array_interface__") or hasattr(a, "__dlpack__"):
        t = torch.from_dlpack(a) if hasattr(a, "__dlpack__") else torch.as_tensor(a, device="cuda")
    if t is None:
        raise TypeError(f"{name}: unsupported array type {type(a).__name__}")
    want = torch.float64 if dtype == np.float64 else getattr(torch, np.dtype(dtype).name)
    if t.dtype != want:
        raise TypeError(f"{name}: expected {want}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous (it is updated in place)")
    if t.device.type == "cpu":
        return Buf(HOST, t.data_ptr(), t.numel(), t)
    if t.device.type != "cuda":
        raise TypeError(f"{name}: unsupported device {t.device}")
    return Buf(DEVICE, t.data_ptr(), t.numel(), t, t.device.index)


def common_kind(bufs) -> str:
    kinds = {b.kind for b in bufs}
    if len(kinds) != 1:
        raise ValueError("evaluate: all arrays must live on the same side (all host or all CUDA)")
    kind = kinds.pop()
    if kind == DEVICE:
        devs = {b.device_index for b in bufs}
        if len(devs) != 1:
            raise ValueError("evaluate: all CUDA arrays must be on the same device")
    return kind


def current_stream_ptr(device_index: int) -> int:
    torch = _torch()
    return int(torch.cuda.current_stream(device_index).cuda_stream)


def host_ptr(a: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.ctypes.data, a
