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


def check_per_gaussian(P, **named):
    """Host-side shape checks of the per-Gaussian inputs: every non-empty tensor must have P rows and the trailing shape
    the kernels index with.  `named` maps an argument name to (tensor, allowed trailing shapes); a trailing shape may
    contain None as a wildcard (e.g. the SH coefficient count).  A mismatch (e.g. rotations left behind by a densify /
    prune desync) would otherwise be an out-of-bounds device read."""
    for name, (t, trailing) in named.items():
        if t is None or t.numel() == 0:
            continue
        if t.shape[0] != P:
            raise RuntimeError(f"{name} has {t.shape[0]} rows, expected {P} (one per Gaussian)")
        tail = tuple(t.shape[1:])
        ok = any(len(tail) == len(want) and all(w is None or w == d for w, d in zip(want, tail)) for want in trailing)
        if not ok:
            raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected ({P}, " + " or ".join(
                ", ".join("*" if w is None else str(w) for w in want) for want in trailing) + ")")
