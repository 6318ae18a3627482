"""TEST INFRASTRUCTURE -- CPU restatement (torch, float32 or float64) of the image-space block that follows the
rasterizer call in GS-SR's 2DGS scene.  Follows /root/reference/gssr/scene/twodgs_scene.py:88-117 (alpha / normal
rotation / median + expected depth with nan_to_num / depth_ratio blend / surf_normal * alpha.detach()) and
/root/reference/gssr/utils/point_utils.py:9-22 (depths_to_points), :24-37 (depth_to_normal).
Parity pinned: tests/golden/post_*.npz come from the reference's own depth_to_normal / depths_to_points (cut out of
point_utils.py at generation time and executed verbatim) composed with the restated render() lines, values and
autograd gradients (tests/golden/make_golden_post.py).  Only tests/ may import this."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def depths_to_points(wvt, full_proj, W, H, depthmap):
    c2w = (wvt.T).inverse()
    ndc2pix = torch.tensor([[W / 2, 0, 0, (W) / 2], [0, H / 2, 0, (H) / 2], [0, 0, 0, 1]]).to(wvt.dtype).T
    projection_matrix = c2w.T @ full_proj
    intrins = (projection_matrix @ ndc2pix)[:3, :3].T
    grid_x, grid_y = torch.meshgrid(torch.arange(W).to(wvt.dtype), torch.arange(H).to(wvt.dtype), indexing="xy")
    points = torch.stack([grid_x, grid_y, torch.ones_like(grid_x)], dim=-1).reshape(-1, 3)
    rays_d = points @ intrins.inverse().T @ c2w[:3, :3].T
    rays_o = c2w[:3, 3]
    return depthmap.reshape(-1, 1) * rays_d + rays_o


def depth_to_normal(wvt, full_proj, W, H, depth):
    points = depths_to_points(wvt, full_proj, W, H, depth).reshape(*depth.shape[1:], 3)
    output = torch.zeros_like(points)
    dx = points[2:, 1:-1] - points[:-2, 1:-1]
    dy = points[1:-1, 2:] - points[1:-1, :-2]
    output[1:-1, 1:-1, :] = F.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    return output


def postprocess(allmap, wvt, full_proj, depth_ratio):
    H, W = allmap.shape[1:]
    render_alpha = allmap[1:2]
    render_normal = (allmap[2:5].permute(1, 2, 0) @ (wvt[:3, :3].T)).permute(2, 0, 1)
    render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
    render_depth_expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
    surf_depth = render_depth_expected * (1 - depth_ratio) + depth_ratio * render_depth_median
    surf_normal = depth_to_normal(wvt, full_proj, W, H, surf_depth).permute(2, 0, 1) * render_alpha.detach()
    return render_normal, surf_depth, surf_normal
