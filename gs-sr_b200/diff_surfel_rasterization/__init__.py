"""Drop-in replacement for GS-SR's ``diff_surfel_rasterization`` extension (2DGS surfels),
backed by libgsr_b200.so (hand-written sm_100a CUDA behind the C ABI in include/gsr_b200.h).

Public surface mirrored from the reference package
(/root/reference/submodules/diff-surfel-rasterization/diff_surfel_rasterization/__init__.py):
  GaussianRasterizationSettings  (:158-170)  same fields, same order
  GaussianRasterizer             (:172-222)  .forward(...) -> (color, radii, allmap), .markVisible
  rasterize_gaussians            (:21-42)
Gradients are returned for (means3D, means2D, sh, colors_precomp, opacities, scales,
rotations, cov3Ds_precomp) exactly as the reference's autograd.Function does (:144-156).

Differences (documented in DESIGN.md):
  * runs on torch's CURRENT stream (the reference uses the legacy default stream);
  * strided inputs (e.g. ``scaling[:, :2]``) are made contiguous for the backward too --
    the reference backward reads the raw strided storage;
  * ``prefiltered=True`` violations raise RuntimeError instead of trapping the context.
There is no CPU fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from gsr_b200 import TorchBuffers, check, lib, ptr
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


def _f32c(t, name, device):
    """float32, contiguous, on `device`; empty tensors pass through (-> NULL)."""
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


def last_num_rendered():
    """True num_rendered (R) of this thread's most recent forward call (bench statistics).  The value the forward
    returns -- and the autograd context carries to the backward -- is the binning layout size (>= R), see
    include/gsr_b200.h "num_rendered WITHOUT a stream sync"."""
    return int(lib().gsr_last_num_rendered())


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def _on_device(device):
    """Device guard only when the tensors live on a non-current device."""
    return _NULL if torch.cuda.current_device() == device.index else torch.cuda.device(device)


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise RuntimeError("means3D must be a CUDA tensor (gsr_b200 has no CPU path)")
        dev = means3D.device
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        M = sh.shape[1] if sh.numel() != 0 else 0

        means3D_c = _f32c(means3D, "means3D", dev)
        sh_c = _f32c(sh, "sh", dev)
        colors_c = _f32c(colors_precomp, "colors_precomp", dev)
        opac_c = _f32c(opacities, "opacities", dev)
        scales_c = _f32c(scales, "scales", dev)
        rot_c = _f32c(rotations, "rotations", dev)
        tm_c = _f32c(cov3Ds_precomp, "cov3Ds_precomp", dev)
        bg = _f32c(rs.bg, "bg", dev)
        view = _f32c(rs.viewmatrix, "viewmatrix", dev)
        proj = _f32c(rs.projmatrix, "projmatrix", dev)
        campos = _f32c(rs.campos, "campos", dev)
        if scales_c is not None and scales_c.numel() and scales_c.shape[-1] != 2:
            raise RuntimeError("surfel scales must have shape (P, 2)")
        check_per_gaussian(P, opacities=(opac_c, [(1,), ()]), scales=(scales_c, [(2,)]), rotations=(rot_c, [(4,)]),
                           colors_precomp=(colors_c, [(3,)]), sh=(sh_c, [(None, 3)]), cov3Ds_precomp=(tm_c, [(9,)]),
                           means2D=(means2D, [(3,)]))

        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        others = torch.empty((11, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        bufs = TorchBuffers(dev)
        with _on_device(dev), bufs:
            if P == 0:
                # reference returns zero-filled outputs without launching (S/rasterize_points.cu:99-100)
                color.zero_(); others.zero_()
                num_rendered = 0
            else:
                num_rendered = check(lib().gsr_surfel_forward(
                    bufs.geom_fn, bufs.binning_fn, bufs.image_fn, bufs.user, P, int(rs.sh_degree), M, ptr(bg), W, H,
                    ptr(means3D_c), ptr(sh_c), ptr(colors_c), ptr(opac_c), ptr(scales_c),
                    float(rs.scale_modifier), ptr(rot_c), ptr(tm_c), ptr(view), ptr(proj), ptr(campos),
                    float(rs.tanfovx), float(rs.tanfovy), int(bool(rs.prefiltered)), ptr(color), ptr(others),
                    ptr(radii), int(bool(rs.debug)), _stream_ptr(dev)), "gsr_surfel_forward")

        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.dims = (P, M, H, W)
        ctx.small = (bg, view, proj, campos)
        ctx.save_for_backward(colors_c if colors_c is not None else torch.empty(0, device=dev),
                              means3D_c, scales_c, rot_c, tm_c, radii, sh_c,
                              bufs.get("geom"), bufs.get("binning"), bufs.get("image"))
        ctx.mark_non_differentiable(radii)
        return color, radii, others

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        rs = ctx.raster_settings
        P, M, H, W = ctx.dims
        bg, view, proj, campos = ctx.small
        (colors_c, means3D_c, scales_c, rot_c, tm_c, radii, sh_c, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        dev = means3D_c.device
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        if grad_depth is None:
            grad_depth = torch.zeros((11, H, W), dtype=torch.float32, device=dev)
        g_color = _f32c(grad_out_color, "grad_out_color", dev)
        g_others = _f32c(grad_depth, "grad_depth", dev)

        def out(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)

        grad_means2D, grad_colors, grad_opac = out(P, 3), out(P, 3), out(P, 1)
        grad_means3D, grad_tm, grad_normal = out(P, 3), out(P, 9), out(P, 3)
        grad_sh = torch.zeros((P, M, 3), dtype=torch.float32, device=dev) if sh_c.numel() == 0 else out(P, M, 3)
        # surfel_preprocess_bwd writes every row of both (zeros for culled Gaussians and with a precomputed transMat)
        grad_scales, grad_rot = out(P, 2), out(P, 4)
        if P != 0:
            with _on_device(dev):
                check(lib().gsr_surfel_backward(
                    P, int(rs.sh_degree), M, int(ctx.num_rendered), ptr(bg), W, H, ptr(means3D_c), ptr(sh_c),
                    ptr(colors_c), ptr(scales_c), float(rs.scale_modifier), ptr(rot_c), ptr(tm_c), ptr(view),
                    ptr(proj), ptr(campos), float(rs.tanfovx), float(rs.tanfovy), ptr(radii), ptr(geomBuffer),
                    ptr(binningBuffer), ptr(imgBuffer), ptr(g_color), ptr(g_others), ptr(grad_means2D),
                    ptr(grad_normal), ptr(grad_opac), ptr(grad_colors), ptr(grad_means3D), ptr(grad_tm),
                    ptr(grad_sh), grad_scales.data_ptr(), grad_rot.data_ptr(), int(bool(rs.debug)),
                    _stream_ptr(dev)), "gsr_surfel_backward")
        return (grad_means3D, grad_means2D, grad_sh, grad_colors, grad_opac, grad_scales, grad_rot, grad_tm,
                None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points in front of the near plane (view z > 0.2)."""
        rs = self.raster_settings
        with torch.no_grad():
            if not positions.is_cuda:
                raise RuntimeError("positions must be a CUDA tensor")
            dev = positions.device
            P = positions.shape[0]
            pos = _f32c(positions, "positions", dev)
            present = torch.zeros((P,), dtype=torch.bool, device=dev)
            if P:
                with _on_device(dev):
                    check(lib().gsr_mark_visible(P, ptr(pos), ptr(_f32c(rs.viewmatrix, "viewmatrix", dev)),
                                                 ptr(_f32c(rs.projmatrix, "projmatrix", dev)),
                                                 present.data_ptr(), _stream_ptr(dev)), "gsr_mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
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
