"""Drop-in replacement for GS-SR's ``diff_gaussian_rasterization`` extension (3DGS),
backed by libgsr_b200.so (hand-written sm_100a CUDA behind the C ABI in include/gsr_b200.h).

Public surface mirrored from the reference package
(/root/reference/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py):
  GaussianRasterizationSettings  (:157-169)  same fields, same order
  GaussianRasterizer             (:171-221)  .forward(...) -> (color, radii), .markVisible
  rasterize_gaussians            (:21-42)
Gradients are returned for (means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
cov3Ds_precomp) as the reference's autograd.Function does (:143-155); means2D.grad is the
NDC-scaled screen-space gradient that drives densification.

Differences: runs on torch's CURRENT stream; strided inputs are made contiguous for the backward too;
``prefiltered=True`` violations raise RuntimeError instead of trapping the context.  There is no CPU
fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from gsr_b200 import TorchBuffers, check, lib, ptr
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


def last_num_rendered():
    """True num_rendered (R) of this thread's most recent forward (the forward itself returns the binning layout size >= R)."""
    return int(lib().gsr_last_num_rendered())


def _check_inputs(means3D, scales):
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (gsr_b200 has no CPU path)")
    if scales is not None and scales.numel() and scales.shape[-1] != 3:
        raise RuntimeError("scales must have shape (P, 3)")


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        rs = raster_settings
        _check_inputs(means3D, scales)
        dev = means3D.device
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        M = sh.shape[1] if sh.numel() != 0 else 0
        means3D_c = f32c(means3D, "means3D", dev)
        sh_c = f32c(sh, "sh", dev)
        colors_c = f32c(colors_precomp, "colors_precomp", dev)
        opac_c = f32c(opacities, "opacities", dev)
        scales_c = f32c(scales, "scales", dev)
        rot_c = f32c(rotations, "rotations", dev)
        cov_c = f32c(cov3Ds_precomp, "cov3Ds_precomp", dev)
        check_per_gaussian(P, opacities=(opac_c, [(1,), ()]), scales=(scales_c, [(3,)]), rotations=(rot_c, [(4,)]),
                           colors_precomp=(colors_c, [(3,)]), sh=(sh_c, [(None, 3)]), cov3Ds_precomp=(cov_c, [(6,)]),
                           means2D=(means2D, [(3,)]))
        bg = f32c(rs.bg, "bg", dev)
        view = f32c(rs.viewmatrix, "viewmatrix", dev)
        proj = f32c(rs.projmatrix, "projmatrix", dev)
        campos = f32c(rs.campos, "campos", dev)

        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        bufs = TorchBuffers(dev)
        with on_device(dev), bufs:
            if P == 0:
                color.zero_()   # reference: zero-filled outputs, nothing launched (G/rasterize_points.cu:76-77)
                num_rendered = 0
            else:
                num_rendered = check(lib().gsr_gaussian_forward(
                    bufs.geom_fn, bufs.binning_fn, bufs.image_fn, bufs.user, P, int(rs.sh_degree), M, ptr(bg), W, H,
                    ptr(means3D_c), ptr(sh_c), ptr(colors_c), ptr(opac_c), ptr(scales_c), float(rs.scale_modifier),
                    ptr(rot_c), ptr(cov_c), ptr(view), ptr(proj), ptr(campos), float(rs.tanfovx), float(rs.tanfovy),
                    int(bool(rs.prefiltered)), ptr(color), ptr(radii), int(bool(rs.debug)), stream_ptr(dev)),
                    "gsr_gaussian_forward")
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.dims = (P, M, H, W)
        ctx.small = (bg, view, proj, campos)
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(colors_c if colors_c is not None else empty, means3D_c, scales_c, rot_c, cov_c, radii,
                              sh_c, bufs.get("geom"), bufs.get("binning"), bufs.get("image"))
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        rs = ctx.raster_settings
        P, M, H, W = ctx.dims
        bg, view, proj, campos = ctx.small
        (colors_c, means3D_c, scales_c, rot_c, cov_c, radii, sh_c, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        dev = means3D_c.device
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        g_color = f32c(grad_out_color, "grad_out_color", dev)

        def out(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)

        def zeros(*shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev)

        has_sr = scales_c.numel() != 0
        grad_means2D, grad_colors, grad_opac, grad_means3D, grad_cov = out(P, 3), out(P, 3), out(P, 1), out(P, 3), out(P, 6)
        grad_sh = zeros(P, M, 3) if sh_c.numel() == 0 else out(P, M, 3)
        grad_scales = out(P, 3) if has_sr else zeros(P, 3)
        grad_rot = out(P, 4) if has_sr else zeros(P, 4)
        if P != 0:
            with on_device(dev):
                check(lib().gsr_gaussian_backward(
                    P, int(rs.sh_degree), M, int(ctx.num_rendered), ptr(bg), W, H, ptr(means3D_c), ptr(sh_c),
                    ptr(colors_c), ptr(scales_c), float(rs.scale_modifier), ptr(rot_c), ptr(cov_c), ptr(view), ptr(proj),
                    ptr(campos), float(rs.tanfovx), float(rs.tanfovy), ptr(radii), ptr(geomBuffer), ptr(binningBuffer),
                    ptr(imgBuffer), ptr(g_color), ptr(grad_means2D), None, ptr(grad_opac), ptr(grad_colors),
                    ptr(grad_means3D), ptr(grad_cov), ptr(grad_sh), ptr(grad_scales), ptr(grad_rot),
                    int(bool(rs.debug)), stream_ptr(dev)), "gsr_gaussian_backward")
        return (grad_means3D, grad_means2D, grad_sh, grad_colors, grad_opac, grad_scales, grad_rot, grad_cov, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


def _mark_visible(rs, positions):
    with torch.no_grad():
        if not positions.is_cuda:
            raise RuntimeError("positions must be a CUDA tensor")
        dev = positions.device
        P = positions.shape[0]
        pos = f32c(positions, "positions", dev)
        present = torch.zeros((P,), dtype=torch.bool, device=dev)
        if P:
            with on_device(dev):
                check(lib().gsr_mark_visible(P, ptr(pos), ptr(f32c(rs.viewmatrix, "viewmatrix", dev)),
                                             ptr(f32c(rs.projmatrix, "projmatrix", dev)), present.data_ptr(),
                                             stream_ptr(dev)), "gsr_mark_visible")
    return present


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points in front of the near plane (view z > 0.2)."""
        return _mark_visible(self.raster_settings, positions)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.empty(0, dtype=torch.float32, device=means3D.device)
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, rs)
