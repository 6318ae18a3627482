"""``simple_knn._C`` -- drop-in for the reference pybind module
(/root/reference/submodules/simple-knn/ext.cpp:14-16, spatial.cu:14-25), backed by libgsr_b200.so.

``distCUDA2(points)`` -> (P,) float32: mean of the three smallest squared distances from each point to the
other points (callers: vanilla_gaussian.py:23, vastgaussian_utils.py:12).  Runs asynchronously on torch's
current stream; the scratch comes from torch's caching allocator (the reference cudaMallocs and synchronises)."""
from __future__ import annotations

import torch

from gsr_b200 import check, lib, ptr
from gsr_b200._torch_util import f32c, on_device, stream_ptr


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor (gsr_b200 has no CPU path)")
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    dev = points.device
    P = points.shape[0]
    pts = f32c(points, "points", dev)
    means = torch.zeros((P,), dtype=torch.float32, device=dev)
    if P:
        L = lib()
        ws = torch.empty((int(L.gsr_dist2_knn3_workspace(P)),), dtype=torch.uint8, device=dev)
        with on_device(dev):
            check(L.gsr_dist2_knn3(P, ptr(pts), ptr(means), ptr(ws), stream_ptr(dev)), "gsr_dist2_knn3")
    return means
