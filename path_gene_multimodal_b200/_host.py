"""Host <-> device staging helpers (pinned buffers, dtype checks). Plumbing only."""
from __future__ import annotations

import numpy as np
import torch

INT32_MAX = 2**31 - 1


def to_device(arr, dtype: np.dtype, device: torch.device, name: str = "array") -> torch.Tensor:
    """numpy (any layout) -> contiguous device tensor of `dtype`, staged through pinned memory."""
    if isinstance(arr, torch.Tensor):
        t = arr.to(device=device, dtype=_torch_dtype(dtype), non_blocking=True)
        return t.contiguous()
    a = np.ascontiguousarray(arr, dtype=dtype)
    if a.size == 0:
        return torch.empty(a.shape, dtype=_torch_dtype(dtype), device=device)
    if a.flags.writeable:  # (a read-only view cannot be page-locked memory of ours; torch would also warn about it)
        t = torch.from_numpy(a)
        if t.is_pinned():  # caller already staged it (e.g. pinned_empty): DMA straight from it
            return t.to(device, non_blocking=True)
    pinned = torch.empty(a.shape, dtype=_torch_dtype(dtype), pin_memory=True)  # torch caches pinned blocks
    pinned.numpy()[...] = a
    return pinned.to(device, non_blocking=True)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array backed by page-locked memory; to_device() copies from it without re-staging."""
    return torch.empty(shape, dtype=_torch_dtype(dtype), pin_memory=True).numpy()


def to_host(t: torch.Tensor) -> np.ndarray:
    """device tensor -> numpy array that owns a pinned buffer (synchronises the current stream)."""
    if t.numel() == 0:
        return np.empty(tuple(t.shape), dtype=_numpy_dtype(t.dtype))
    pinned = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    pinned.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return pinned.numpy()


def to_host_many(tensors: dict, sync: bool = True) -> dict:
    """Several device tensors -> numpy with one stream synchronisation (``sync=False``: the caller synchronises the
    stream itself before it reads the arrays)."""
    out, dev = {}, None
    for k, t in tensors.items():
        if t is None:
            continue
        if t.numel() == 0:
            out[k] = np.empty(tuple(t.shape), dtype=_numpy_dtype(t.dtype))
            continue
        pinned = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        pinned.copy_(t, non_blocking=True)
        out[k] = pinned
        dev = t.device
    if dev is not None and sync:
        torch.cuda.current_stream(dev).synchronize()
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def as_int32(arr, name: str) -> np.ndarray:
    a = np.asarray(arr)
    if a.dtype == np.int32:
        return a
    if a.dtype.kind not in "iu":
        if a.dtype.kind == "f" and np.all(a == np.floor(a)):
            a = a.astype(np.int64)
        else:
            raise TypeError(f"{name}: integer values required, got dtype {a.dtype}")
    if a.size and (a.max() > INT32_MAX or a.min() < -INT32_MAX - 1):
        raise OverflowError(f"{name}: values do not fit int32")
    return a.astype(np.int32)


def _torch_dtype(dt) -> torch.dtype:
    return {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
            np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
            np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16}[np.dtype(dt)]


def _numpy_dtype(dt: torch.dtype):
    return {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32,
            torch.int64: np.int64, torch.uint8: np.uint8, torch.int16: np.int16}[dt]
