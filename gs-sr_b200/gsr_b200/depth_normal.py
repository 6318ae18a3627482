"""Fused ``normal_from_depth_image`` for PGSR -- drop-in for the torch chain
``depth2point_world -> depth_pcd2normal`` (/root/reference/gssr/utils/graphics_utils.py:88-146) that
``PGSRScene.render_normal`` runs on the rendered plane depth every iteration (gssr/scene/pgsr_scene.py:227-238,320):

    from gsr_b200.depth_normal import normal_from_depth_image
    normal_ref = normal_from_depth_image(depth, intrinsic_matrix, extrinsic_matrix)          # (H, W, 3), like the reference
    depth_normal = render_normal_weighted(depth, intrinsic_matrix, rendered_alpha.detach())  # (3, H, W), x alpha fused

The reference spends ~25 full-frame torch kernels (meshgrid, stack, matmul, four slices, cross, normalize, pad) and their
autograd twins on this; here it is one forward kernel and two backward kernels (csrc/depth_normal.cu).  Only
``offset=None`` is implemented -- GS-SR never passes an offset (pgsr_scene.py:320) -- anything else raises.
``extrinsic_matrix`` is accepted and ignored, exactly as the reference ignores it (graphics_utils.py:101-108).
No CPU / PyTorch fallback.
"""
from __future__ import annotations

import torch

from . import check, lib
from ._torch_util import f32c, on_device, stream_ptr


class _DepthNormal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, kinv, weight):
        dev = depth.device
        d = f32c(depth, "depth", dev)
        H, W = d.shape
        w = None if weight is None else f32c(weight, "weight", dev).reshape(H, W)
        out = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        with on_device(dev):
            check(lib().gsr_depth_normal_forward(H, W, d.data_ptr(), kinv.data_ptr(), None if w is None else w.data_ptr(),
                                                 out.data_ptr(), stream_ptr(dev)), "gsr_depth_normal_forward")
        ctx.save_for_backward(d, kinv, w if w is not None else torch.empty(0, device=dev))
        return out

    @staticmethod
    def backward(ctx, g):
        d, kinv, w = ctx.saved_tensors
        dev = d.device
        H, W = d.shape
        g = f32c(g, "grad", dev)
        scratch = torch.empty((6, H, W), dtype=torch.float32, device=dev)
        gd = torch.empty((H, W), dtype=torch.float32, device=dev)
        with on_device(dev):
            check(lib().gsr_depth_normal_backward(H, W, d.data_ptr(), kinv.data_ptr(), w.data_ptr() if w.numel() else None,
                                                  g.data_ptr(), scratch.data_ptr(), gd.data_ptr(), stream_ptr(dev)),
                  "gsr_depth_normal_backward")
        return gd, None, None


def _kinv(intrinsic_matrix, device):
    # ndc_2_cam (:86): cam_xyz @ torch.inverse(intrinsic.t()) -- the same float32 op on the same device, no host read-back
    # linalg.inv_ex = torch.inverse's factorisation without its status read-back (no stream sync, graph-capturable)
    return torch.linalg.inv_ex(intrinsic_matrix.to(device=device, dtype=torch.float32).t())[0].contiguous()


def render_normal_weighted(depth, intrinsic_matrix, weight=None):
    """(3, H, W) normals of a (H, W) depth map, optionally multiplied per pixel by ``weight`` (no gradient to it)."""
    if not depth.is_cuda:
        raise RuntimeError("depth must be a CUDA tensor (gsr_b200 has no CPU path)")
    if depth.dim() != 2:
        raise RuntimeError("depth must have shape (H, W)")
    return _DepthNormal.apply(depth, _kinv(intrinsic_matrix, depth.device), None if weight is None else weight.detach())


def normal_from_depth_image(depth, intrinsic_matrix, extrinsic_matrix=None, offset=None, gt_image=None):
    """Same signature and (H, W, 3) result as the reference function."""
    if offset is not None:
        raise NotImplementedError("normal_from_depth_image: only offset=None is implemented (GS-SR never passes an offset)")
    return render_normal_weighted(depth, intrinsic_matrix).permute(1, 2, 0)
