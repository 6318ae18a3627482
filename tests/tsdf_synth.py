"""Seeded synthetic inputs for the TSDF-fusion tests: V cameras (GS-SR conventions: world_view_transform
and full_proj_transform transposed / row-vector, gssr/cameras/__init__.py:80-88, graphics_utils.py:38-71)
looking at a sphere of radius 0.5 above a ground plane, with analytically ray-cast depth (view-space z,
0 = no hit) and procedural RGB maps, plus sample points in contracted or world coordinates."""
from __future__ import annotations

import math

import numpy as np

TSDF_CASES = ("contracted", "world_rgb", "ragged")


def _world2view(R, t):
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = R.T
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    return Rt


def _projection(znear, zfar, fovx, fovy):
    th, tw = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = th * znear, tw * znear
    P = np.zeros((4, 4), dtype=np.float64)
    P[0, 0] = 2 * znear / (2 * right)
    P[1, 1] = 2 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(eye, target, W, H, fovx):
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, -1.0, 0.0]) if abs(fwd[1]) < 0.95 else np.array([0.0, 0.0, 1.0])
    right = np.cross(up, fwd); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w_R = np.stack([right, down, fwd], axis=1)           # columns: camera x, y, z axes in world
    w2c_R = c2w_R.T
    t = -w2c_R @ eye
    fovy = 2 * math.atan(math.tan(fovx / 2) * H / W)
    wvt = _world2view(w2c_R.T, t).T                        # world_view_transform (transposed, row-vector)
    proj = _projection(0.01, 100.0, fovx, fovy).T
    full = wvt @ proj
    return dict(world_view=wvt.astype(np.float32), full_proj=full.astype(np.float32), W=W, H=H, fovx=fovx, fovy=fovy,
                w2c_R=w2c_R, t=t, eye=eye)


def raycast_maps(cam, sphere_r=0.5, ground_y=0.55):
    """Depth = view-space z of the first hit of {sphere at origin, plane y = ground_y} (y points down in this
    world), 0 where nothing is hit; RGB = smooth function of the hit point."""
    W, H = cam["W"], cam["H"]
    tx, ty = math.tan(cam["fovx"] / 2), math.tan(cam["fovy"] / 2)
    xs = ((np.arange(W) + 0.5) / W * 2 - 1) * tx
    ys = ((np.arange(H) + 0.5) / H * 2 - 1) * ty
    gx, gy = np.meshgrid(xs, ys)
    d_cam = np.stack([gx, gy, np.ones_like(gx)], -1)       # z = 1 in camera space: t_ray == view z
    d_w = d_cam @ cam["w2c_R"]                              # = R_c2w @ d_cam
    o = cam["eye"]
    a = (d_w * d_w).sum(-1); b = 2 * (d_w @ o); c = o @ o - sphere_r ** 2
    disc = b * b - 4 * a * c
    ts = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
    ts = np.where(ts > 0.05, ts, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tp = (ground_y - o[1]) / d_w[..., 1]
    tp = np.where((tp > 0.05) & (tp < 6.0), tp, np.inf)
    t = np.minimum(ts, tp)
    hit = np.isfinite(t)
    depth = np.where(hit, t, 0.0).astype(np.float32)
    pw = o + d_w * np.where(hit, t, 0.0)[..., None]
    rgb = np.stack([0.5 + 0.5 * np.sin(3 * pw[..., 0]), 0.5 + 0.5 * np.cos(2 * pw[..., 1] + 1), 0.5 + 0.5 * np.sin(4 * pw[..., 2] + 2)], 0)
    rgb = np.where(hit[None], rgb, 0.0).astype(np.float32)
    return depth[None], rgb                                 # (1,H,W), (3,H,W) like GaussianExtractor.depthmaps / rgbmaps


def build_tsdf_case(name, n=None):
    rng = np.random.default_rng({"contracted": 301, "world_rgb": 302, "ragged": 303, "bench": 304}[name])
    if name == "ragged":
        sizes = [(96, 64), (80, 50), (33, 47), (96, 64)]
        eyes = [(0.0, -0.3, -2.0), (1.8, -0.5, 0.6), (-1.2, -1.4, 1.0), (0.2, -0.2, 2.2)]
        targets = [(0, 0, 0), (0, 0, 0), (0, 0, 0), (0.5, -0.2, 5.0)]      # last camera looks away from the scene
    elif name == "bench":
        V = 32
        sizes = [(1600, 1060)] * V
        ang = np.linspace(0, 2 * np.pi, V, endpoint=False)
        eyes = [(2.0 * math.cos(a), -0.4 - 0.3 * math.sin(3 * a), 2.0 * math.sin(a)) for a in ang]
        targets = [(0, 0, 0)] * V
    else:
        V = 5
        sizes = [(96, 64)] * V
        ang = np.linspace(0, 2 * np.pi, V, endpoint=False)
        eyes = [(2.0 * math.cos(a), -0.4 - 0.3 * math.sin(3 * a), 2.0 * math.sin(a)) for a in ang]
        targets = [(0, 0, 0)] * V
    cams = [look_at_camera(e, t, W, H, math.radians(60)) for e, t, (W, H) in zip(eyes, targets, sizes)]
    maps = [raycast_maps(c) for c in cams]
    center = np.array([0.05, -0.1, 0.02], dtype=np.float32)
    radius = 1.5
    if name == "world_rgb":
        # points near the surfaces in WORLD coordinates (the mesh-vertex colouring call, inv_contraction=None)
        n = n or 20000
        dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        samples = (dirs * (0.5 + rng.normal(scale=0.02, size=(n, 1)))).astype(np.float32)
        contracted = False
        voxel = 0.01
    else:
        n = n or 30000
        samples = rng.uniform(-1.7, 1.7, size=(n, 3)).astype(np.float32)
        samples[: n // 3] *= 0.3                              # a good share inside the unit ball / near the sphere
        contracted = True
        voxel = 2 * radius / 256
    return dict(samples=samples, contracted=contracted, center=center, radius=radius, voxel_size=voxel,
                projs=[c["full_proj"] for c in cams], depthmaps=[m[0] for m in maps], rgbmaps=[m[1] for m in maps],
                eyes=np.array(eyes, dtype=np.float64))


def torch_rule_unbounded(samples, projs, depths, rgbs, center, radius, voxel_size):
    """The per-view torch rule of GaussianExtractor.extract_mesh_unbounded (mesh_utils.py:195-246) with the same torch ops
    on whatever device the tensors live on -- the reference arm of the TSDF timing (tests/bench_extras.py,
    tests/gpu_tsdf_bench.py).  rgbs may be None (depth only)."""
    import torch
    norm = torch.linalg.norm(samples, dim=-1)
    mask = norm > 1
    sdf_trunc = 5 * voxel_size * torch.ones_like(samples[:, 0])
    sdf_trunc[mask] *= 1 / (2 - norm[mask].clamp(max=1.9))
    mag = norm[..., None]
    samples = torch.where(mag < 1, samples, (1 / (2 - mag) * (samples / mag))) * radius + center
    tsdfs = torch.ones_like(samples[:, 0])
    rgbo = torch.zeros((samples.shape[0], 3), device=samples.device)
    weights = torch.ones_like(samples[:, 0])
    for i in range(len(projs)):
        new_points = torch.cat([samples, torch.ones_like(samples[..., :1])], dim=-1) @ projs[i]
        z = new_points[..., -1:]
        pix = new_points[..., :2] / new_points[..., -1:]
        mask_proj = ((pix > -1.) & (pix < 1.) & (z > 0)).all(dim=-1)
        sd = torch.nn.functional.grid_sample(depths[i][None], pix[None, None], mode='bilinear', padding_mode='border',
                                             align_corners=True).reshape(-1, 1)
        sdf = (sd - z).flatten()
        mask_proj = mask_proj & (sdf > -sdf_trunc)
        sdf = torch.clamp(sdf / sdf_trunc, min=-1.0, max=1.0)[mask_proj]
        w = weights[mask_proj]
        wp = w + 1
        tsdfs[mask_proj] = (tsdfs[mask_proj] * w + sdf) / wp
        if rgbs is not None:
            sr = torch.nn.functional.grid_sample(rgbs[i][None], pix[None, None], mode='bilinear', padding_mode='border',
                                                 align_corners=True).reshape(3, -1).T
            rgbo[mask_proj] = (rgbo[mask_proj] * w[:, None] + sr[mask_proj]) / wp[:, None]
        weights[mask_proj] = wp
    return tsdfs
