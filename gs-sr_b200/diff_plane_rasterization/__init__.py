"""Drop-in replacement for GS-SR's ``diff_plane_rasterization`` extension (PGSR planar Gaussians),
backed by libgsr_b200.so (hand-written sm_100a CUDA behind the C ABI in include/gsr_b200.h).

Public surface mirrored from the reference package
(/root/reference/submodules/diff-plane-rasterization/diff_plane_rasterization/__init__.py):
  GaussianRasterizationSettings  (:173-186)  same fields, same order (adds render_geo)
  GaussianRasterizer             (:188-245)  .forward(means3D, means2D, means2D_abs, opacities, shs, colors_precomp,
                                             scales, rotations, cov3D_precomp, all_map)
                                             -> (color, radii, out_observe, out_all_map, plane_depth)
  rasterize_gaussians            (:21-46)
Gradients are returned for (means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations,
cov3Ds_precomp, all_map) as the reference's autograd.Function does (:155-169).

Differences: current stream, contiguous copies of strided inputs for the backward too, RuntimeError instead of
a device trap for prefiltered violations; ``debug=True`` works (the reference's debug branch unpacks the wrong
number of results).  No CPU fallback.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from gsr_b200 import TorchBuffers, check, lib, ptr
from gsr_b200._torch_util import f32c, on_device, stream_ptr
from diff_gaussian_rasterization import _check_inputs, _mark_visible
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
    render_geo: bool
    debug: bool


def last_num_rendered():
    """True num_rendered (R) of this thread's most recent forward (the forward itself returns the binning layout size >= R)."""
    return int(lib().gsr_last_num_rendered())


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                all_maps, raster_settings):
        rs = raster_settings
        _check_inputs(means3D, scales)
        dev = means3D.device
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        M = sh.shape[1] if sh.numel() != 0 else 0
        geo = bool(rs.render_geo)
        means3D_c = f32c(means3D, "means3D", dev)
        sh_c = f32c(sh, "sh", dev)
        colors_c = f32c(colors_precomp, "colors_precomp", dev)
        opac_c = f32c(opacities, "opacities", dev)
        scales_c = f32c(scales, "scales", dev)
        rot_c = f32c(rotations, "rotations", dev)
        cov_c = f32c(cov3Ds_precomp, "cov3Ds_precomp", dev)
        am_c = f32c(all_maps, "all_map", dev)
        check_per_gaussian(P, opacities=(opac_c, [(1,), ()]), scales=(scales_c, [(3,)]), rotations=(rot_c, [(4,)]),
                           colors_precomp=(colors_c, [(3,)]), sh=(sh_c, [(None, 3)]), cov3Ds_precomp=(cov_c, [(6,)]),
                           all_map=(am_c, [(5,)]), means2D=(means2D, [(3,)]), means2D_abs=(means2D_abs, [(3,)]))
        if geo and P and (am_c is None or am_c.numel() == 0 or am_c.shape[-1] != 5):
            raise RuntimeError("all_map must have shape (P, 5) when render_geo is set")
        bg = f32c(rs.bg, "bg", dev)
        view = f32c(rs.viewmatrix, "viewmatrix", dev)
        proj = f32c(rs.projmatrix, "projmatrix", dev)
        campos = f32c(rs.campos, "campos", dev)

        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        out_observe = torch.empty((P,), dtype=torch.int32, device=dev)
        out_all_map = torch.empty((5, H, W), dtype=torch.float32, device=dev)
        out_plane_depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        bufs = TorchBuffers(dev)
        with on_device(dev), bufs:
            if P == 0:
                color.zero_(); out_all_map.zero_(); out_plane_depth.zero_()
                num_rendered = 0
            else:
                num_rendered = check(lib().gsr_plane_forward(
                    bufs.geom_fn, bufs.binning_fn, bufs.image_fn, bufs.user, P, int(rs.sh_degree), M, ptr(bg), W, H,
                    ptr(means3D_c), ptr(sh_c), ptr(colors_c), ptr(opac_c), ptr(scales_c), float(rs.scale_modifier),
                    ptr(rot_c), ptr(cov_c), ptr(am_c), ptr(view), ptr(proj), ptr(campos), float(rs.tanfovx),
                    float(rs.tanfovy), int(bool(rs.prefiltered)), ptr(color), ptr(radii), ptr(out_observe),
                    ptr(out_all_map), ptr(out_plane_depth), int(geo), int(bool(rs.debug)), stream_ptr(dev)),
                    "gsr_plane_forward")
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.dims = (P, M, H, W)
        ctx.small = (bg, view, proj, campos)
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(out_all_map, colors_c if colors_c is not None else empty,
                              am_c if am_c is not None else empty, means3D_c, scales_c, rot_c, cov_c, radii, sh_c,
                              bufs.get("geom"), bufs.get("binning"), bufs.get("image"))
        ctx.mark_non_differentiable(radii, out_observe)
        return color, radii, out_observe, out_all_map, out_plane_depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_out_observe, grad_out_all_map, grad_out_plane_depth):
        rs = ctx.raster_settings
        P, M, H, W = ctx.dims
        geo = bool(rs.render_geo)
        bg, view, proj, campos = ctx.small
        (all_map_pixels, colors_c, am_c, means3D_c, scales_c, rot_c, cov_c, radii, sh_c, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        dev = means3D_c.device

        def out(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)

        def zeros(*shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev)

        g_color = f32c(grad_out_color if grad_out_color is not None else zeros(3, H, W), "grad_out_color", dev)
        g_am = f32c(grad_out_all_map if grad_out_all_map is not None else zeros(5, H, W), "grad_out_all_map", dev)
        g_pd = f32c(grad_out_plane_depth if grad_out_plane_depth is not None else zeros(1, H, W),
                    "grad_out_plane_depth", dev)
        has_sr = scales_c.numel() != 0
        grad_means2D, grad_means2D_abs, grad_colors, grad_opac = out(P, 3), out(P, 3), out(P, 3), out(P, 1)
        grad_means3D, grad_cov, grad_all_map = out(P, 3), out(P, 6), out(P, 5)
        grad_sh = zeros(P, M, 3) if sh_c.numel() == 0 else out(P, M, 3)
        grad_scales = out(P, 3) if has_sr else zeros(P, 3)
        grad_rot = out(P, 4) if has_sr else zeros(P, 4)
        if P != 0:
            with on_device(dev):
                check(lib().gsr_plane_backward(
                    P, int(rs.sh_degree), M, int(ctx.num_rendered), ptr(bg), ptr(all_map_pixels), W, H, ptr(means3D_c),
                    ptr(sh_c), ptr(colors_c), ptr(am_c), ptr(scales_c), float(rs.scale_modifier), ptr(rot_c),
                    ptr(cov_c), ptr(view), ptr(proj), ptr(campos), float(rs.tanfovx), float(rs.tanfovy), ptr(radii),
                    ptr(geomBuffer), ptr(binningBuffer), ptr(imgBuffer), ptr(g_color), ptr(g_am), ptr(g_pd),
                    ptr(grad_means2D), ptr(grad_means2D_abs), None, ptr(grad_opac), ptr(grad_colors),
                    ptr(grad_means3D), ptr(grad_cov), ptr(grad_sh), ptr(grad_scales), ptr(grad_rot),
                    ptr(grad_all_map), int(geo), int(bool(rs.debug)), stream_ptr(dev)), "gsr_plane_backward")
        if am_c.numel() == 0:
            grad_all_map = None
        return (grad_means3D, grad_means2D, grad_means2D_abs, grad_sh, grad_colors, grad_opac, grad_scales, grad_rot,
                grad_cov, grad_all_map, None)


def rasterize_gaussians(means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, all_map, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, all_map, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        return _mark_visible(self.raster_settings, positions)

    def forward(self, means3D, means2D, means2D_abs, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, all_map=None):
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
        all_map = empty if all_map is None else all_map
        return rasterize_gaussians(means3D, means2D, means2D_abs, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, all_map, rs)
