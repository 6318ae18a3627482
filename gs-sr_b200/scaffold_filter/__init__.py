"""Drop-in replacement for GS-SR's ``scaffold_filter`` extension (anchor pre-filter of
Scaffold-GS / Octree-GS), backed by libgsr_b200.so.

Public surface mirrored from /root/reference/submodules/scaffold-filter/scaffold_filter/__init__.py:
  GaussianRasterizationSettings (:160-172), GaussianRasterizer.visible_filter(means3D, scales, rotations,
  cov3D_precomp) -> radii (P,) int32 (:225-251).  Callers use ``radii > 0`` (scaffold_scene.py:149-155,
  octree_scene.py:164-172).

Unlike the reference (which resizes a ~100 B/anchor geometry chunk and a 12 B/pixel image chunk per call,
F/cuda_rasterizer/rasterizer_impl.cu:361-376) the call allocates nothing but the result and stays
asynchronous on torch's current stream.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from gsr_b200 import check, lib, ptr
from gsr_b200._torch_util import f32c, on_device, stream_ptr
from gsr_b200._torch_util import check_per_gaussian


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise RuntimeError("means3D must be a CUDA tensor (gsr_b200 has no CPU path)")
        with torch.no_grad():
            dev = means3D.device
            P = means3D.shape[0]
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            if P:
                sc = f32c(scales, "scales", dev)
                if sc is not None and sc.numel() and sc.shape[-1] != 3:
                    raise RuntimeError("scales must have shape (P, 3)")
                check_per_gaussian(P, scales=(sc, [(3,)]), rotations=(rotations, [(4,)]), cov3D_precomp=(cov3D_precomp, [(6,)]))
                with on_device(dev):
                    check(lib().gsr_visible_filter(
                        P, int(rs.image_width), int(rs.image_height), ptr(f32c(means3D, "means3D", dev)), ptr(sc),
                        float(rs.scale_modifier), ptr(f32c(rotations, "rotations", dev)),
                        ptr(f32c(cov3D_precomp, "cov3D_precomp", dev)), ptr(f32c(rs.viewmatrix, "viewmatrix", dev)),
                        ptr(f32c(rs.projmatrix, "projmatrix", dev)), float(rs.tanfovx), float(rs.tanfovy),
                        int(bool(rs.prefiltered)), ptr(radii), int(bool(rs.debug)), stream_ptr(dev)),
                        "gsr_visible_filter")
        return radii
