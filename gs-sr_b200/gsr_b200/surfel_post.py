"""Fused post-processing of the 2DGS ``allmap`` -- drop-in for the torch block that follows the rasterizer call in
``TwoDGSScene.render`` (/root/reference/gssr/scene/twodgs_scene.py:88-117, with ``depth_to_normal`` /
``depths_to_points`` of gssr/utils/point_utils.py:9-37):

    from gsr_b200.surfel_post import surfel_postprocess
    post = surfel_postprocess(allmap, viewpoint_camera.world_view_transform, viewpoint_camera.full_proj_transform,
                              depth_ratio=self.config.depth_ratio)
    rets.update({'rend_alpha': post['rend_alpha'], 'rend_dist': post['rend_dist'], 'surf_normal': post['surf_normal'],
                 'depth': post['depth'], 'normal': post['normal']})

Values and gradients w.r.t. ``allmap`` equal autograd's through the reference ops (tests/test_surfel_post_gpu.py), except
that pixels with zero alpha get a zero gradient where the reference's 0/0 division adjoint produces NaN.  The tiny per-view
camera algebra (two small inverses, as in depths_to_points) stays in torch on the device (``linalg.inv_ex``: no host
synchronisation, capturable into a CUDA graph).
No CPU / PyTorch fallback for the image-sized work.
"""
from __future__ import annotations

import torch

from . import check, lib
from ._torch_util import f32c, on_device, stream_ptr


def _inv(m):
    """torch.inverse without its error check: the same LU factorisation (identical values), but no device->host read
    of the status word -- torch.inverse synchronises the stream, which also makes it illegal during CUDA-graph capture."""
    return torch.linalg.inv_ex(m)[0]


_NDC2PIX = {}


def _ndc2pix(W, H, device):
    """The constant matrix of depths_to_points (:14-17), uploaded once per (W, H, device): building it from a Python
    list on every call is a synchronous host->device copy (and cannot be recorded into a CUDA graph)."""
    key = (W, H, str(device))
    m = _NDC2PIX.get(key)
    if m is None:
        m = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]], dtype=torch.float32, device=device).T
        _NDC2PIX[key] = m
    return m


def _camera_constants(world_view_transform, full_proj_transform, W, H):
    """K, rays_o, R packed as 21 floats on the device -- the same float32 torch ops as depths_to_points (:10-22)."""
    wvt = world_view_transform.float()
    c2w = _inv(wvt.T)
    ndc2pix = _ndc2pix(W, H, wvt.device)
    projection_matrix = c2w.T @ full_proj_transform.float()
    intrins = (projection_matrix @ ndc2pix)[:3, :3].T
    K = _inv(intrins).T @ c2w[:3, :3].T
    return torch.cat([K.reshape(9), c2w[:3, 3].reshape(3), wvt[:3, :3].reshape(9)]).contiguous()


class _SurfelPost(torch.autograd.Function):
    @staticmethod
    def forward(ctx, allmap, cam21, depth_ratio):
        dev = allmap.device
        am = f32c(allmap, "allmap", dev)
        _, H, W = am.shape
        rn = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        sd = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        sn = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        with on_device(dev):
            check(lib().gsr_surfel_post_forward(H, W, am.data_ptr(), cam21.data_ptr(), float(depth_ratio), rn.data_ptr(),
                                                sd.data_ptr(), sn.data_ptr(), stream_ptr(dev)), "gsr_surfel_post_forward")
        ctx.save_for_backward(am, cam21, sd)
        ctx.depth_ratio = float(depth_ratio)
        return rn, sd, sn

    @staticmethod
    def backward(ctx, g_rn, g_sd, g_sn):
        am, cam21, sd = ctx.saved_tensors
        dev = am.device
        _, H, W = am.shape
        ptr = lambda t: None if t is None else f32c(t, "grad", dev).data_ptr()  # noqa: E731
        keep = [None if t is None else f32c(t, "grad", dev) for t in (g_rn, g_sd, g_sn)]
        scratch = torch.empty((6, H, W), dtype=torch.float32, device=dev)
        out = torch.empty_like(am)
        with on_device(dev):
            check(lib().gsr_surfel_post_backward(H, W, am.data_ptr(), cam21.data_ptr(), ctx.depth_ratio, sd.data_ptr(),
                                                 *[None if t is None else t.data_ptr() for t in keep], scratch.data_ptr(),
                                                 out.data_ptr(), stream_ptr(dev)), "gsr_surfel_post_backward")
        return out, None, None


def surfel_postprocess(allmap, world_view_transform, full_proj_transform, depth_ratio=0.0):
    if allmap.dim() != 3 or allmap.shape[0] != 11:
        raise RuntimeError("allmap must have dimensions (11, H, W)")
    if not allmap.is_cuda:
        raise RuntimeError("allmap must be a CUDA tensor (gsr_b200 has no CPU path)")
    H, W = allmap.shape[1], allmap.shape[2]
    with torch.no_grad():
        cam21 = _camera_constants(world_view_transform.to(allmap.device), full_proj_transform.to(allmap.device), W, H)
    normal, depth, surf_normal = _SurfelPost.apply(allmap, cam21, depth_ratio)
    return {"rend_alpha": allmap[1:2], "rend_dist": allmap[6:7], "surf_normal": surf_normal, "depth": depth, "normal": normal}
