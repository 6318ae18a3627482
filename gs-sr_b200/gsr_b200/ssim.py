"""Fused SSIM -- drop-in for ``VanillaScene.ssim(img1, img2, window_size=11, size_average=True)``
(/root/reference/gssr/scene/vanilla_scene.py:32-61) backed by ``gsr_ssim_forward/backward`` (gs-sr_b200/csrc/ssim.cu).

    from gsr_b200.ssim import ssim
    loss = lambda_dssim * (1.0 - ssim(image, gt_image))

Same value and the same gradient w.r.t. ``img1`` as the reference's conv2d formulation (zero-padded 11x11 Gaussian
window, sigma 1.5, C1 = 0.01^2, C2 = 0.03^2, mean over all pixels and channels).  ``img2`` is treated as a constant
(GS-SR passes the ground-truth image); asking for its gradient raises.  No CPU / PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import check, lib
from ._torch_util import f32c, on_device, stream_ptr


def _gaussian11():
    # vanilla_scene.py:50-52 evaluated the same way: float32 taps divided by their float32 sum
    g = torch.Tensor([math.exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = g / g.sum()
    return (ctypes.c_float * 11)(*g.tolist())


_WINDOW = _gaussian11()


class _FusedSSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2):
        dev = img1.device
        x, y = f32c(img1, "img1", dev), f32c(img2, "img2", dev)
        H, W = x.shape[-2], x.shape[-1]
        Cn = x.numel() // (H * W)
        L = lib()
        maps = torch.empty((3,) + tuple(x.shape), dtype=torch.float32, device=dev)
        sums = torch.empty((int(L.gsr_ssim_tile_count(Cn, H, W)),), dtype=torch.float32, device=dev)
        with on_device(dev):
            check(L.gsr_ssim_forward(Cn, H, W, x.data_ptr(), y.data_ptr(), _WINDOW, maps[0].data_ptr(), maps[1].data_ptr(),
                                     maps[2].data_ptr(), sums.data_ptr(), stream_ptr(dev)), "gsr_ssim_forward")
        ctx.save_for_backward(x, y, maps)
        ctx.dims = (Cn, H, W)
        return sums.sum() / float(Cn * H * W)

    @staticmethod
    def backward(ctx, grad_out):
        x, y, maps = ctx.saved_tensors
        Cn, H, W = ctx.dims
        if ctx.needs_input_grad[1]:
            raise RuntimeError("gsr_b200.ssim: gradient w.r.t. img2 is not implemented (GS-SR passes the ground truth there)")
        dev = x.device
        g = grad_out.to(dev, torch.float32).reshape(1).contiguous()
        dx = torch.empty_like(x)
        with on_device(dev):
            check(lib().gsr_ssim_backward(Cn, H, W, x.data_ptr(), y.data_ptr(), _WINDOW, maps[0].data_ptr(), maps[1].data_ptr(),
                                          maps[2].data_ptr(), g.data_ptr(), dx.data_ptr(), stream_ptr(dev)), "gsr_ssim_backward")
        return dx, None


def ssim(img1, img2, window_size=11, size_average=True):
    if window_size != 11 or not size_average:
        raise NotImplementedError("gsr_b200.ssim implements the configuration GS-SR uses: window_size=11, size_average=True")
    if img1.shape != img2.shape or img1.dim() not in (3, 4):
        raise RuntimeError("ssim expects two (C,H,W) or (B,C,H,W) images of the same shape")
    if not img1.is_cuda or not img2.is_cuda:
        raise RuntimeError("ssim needs CUDA tensors (gsr_b200 has no CPU path)")
    return _FusedSSIM.apply(img1, img2)
