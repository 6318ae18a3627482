"""CPU oracle for PGSR's normal_from_depth_image -- TEST INFRASTRUCTURE (tests/, smoke(); never imported by the product).

Restates /root/reference/gssr/utils/graphics_utils.py:79-146 (ndc_2_cam :79-86, depth2point_cam :88-99,
depth2point_world :101-108, depth_pcd2normal :110-137 with offset=None, normal_from_depth_image :139-146) in plain torch on
the CPU, float32 or float64; gradients through autograd.  Pinned by tests/golden/depth_normal_*.npz, which are outputs of
the reference's own functions (tests/golden/make_golden_depth_normal.py).
"""
import torch


def normal_from_depth_image(depth, intrinsic, dtype=torch.float32):
    depth = depth.to(dtype)
    H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    # (x/(W-1)) * (W-1) * z of the reference collapses to x * z up to one rounding; keep the reference's two steps
    vx, vy = xs / (W - 1), ys / (H - 1)
    cam = torch.stack([vx * (W - 1) * depth, vy * (H - 1) * depth, depth], dim=-1) @ torch.inverse(intrinsic.to(dtype).t())
    l2r = cam[1:H - 1, 2:W] - cam[1:H - 1, 0:W - 2]
    b2t = cam[0:H - 2, 1:W - 1] - cam[2:H, 1:W - 1]
    n = torch.nn.functional.normalize(torch.cross(l2r, b2t, dim=-1), p=2, dim=-1)
    return torch.nn.functional.pad(n.permute(2, 0, 1), (1, 1, 1, 1), mode="constant").permute(1, 2, 0)


def value_and_grad(depth_np, intrinsic_np, g_np, weight_np=None, dtype=torch.float32):
    """-> (normal (3,H,W), dL/ddepth (H,W)) for upstream g (3,H,W); weight multiplies the normals (no gradient to it)."""
    d = torch.from_numpy(depth_np).to(dtype).requires_grad_(True)
    n = normal_from_depth_image(d, torch.from_numpy(intrinsic_np), dtype).permute(2, 0, 1)
    if weight_np is not None:
        n = n * torch.from_numpy(weight_np).to(dtype)
    n.backward(torch.from_numpy(g_np).to(dtype))
    return n.detach().numpy(), d.grad.numpy()
