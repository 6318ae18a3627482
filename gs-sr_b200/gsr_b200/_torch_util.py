"""Small torch-side helpers shared by the drop-in packages (plumbing only: dtype/device checks,
device guard, current-stream handle)."""
from __future__ import annotations

import torch


def f32c(t, name, device):
    """float32, contiguous, on `device`; empty / None tensors pass through (-> NULL pointer)."""
    if t is None:
        return None
    if t.numel() == 0:
        return t
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.device != device:
        raise RuntimeError(f"{name} is on {t.device}, expected {device}")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def on_device(device):
    """Device guard only when the tensors live on a non-current device."""
    return _NULL if torch.cuda.current_device() == device.index else torch.cuda.device(device)


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream
