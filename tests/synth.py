"""Seeded synthetic scenes for parity tests and bench.py.

Camera conventions follow the reference's Camera class
(/root/reference/gssr/cameras/__init__.py:85-88, gssr/utils/graphics_utils.py:51-71):
row-vector matrices, ``world_view_transform`` = W2C transposed,
``full_proj_transform`` = view @ proj, znear 0.01, zfar 100.

Everything is numpy float32 so the same arrays feed the CPU oracle, the
reference CUDA build and the B200 library.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

C0 = 0.28209479177387814


def projection_matrix(znear, zfar, fovx, fovy):
    """getProjectionMatrix (graphics_utils.py:51-71), float32 like the torch original."""
    thy, thx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = thy * znear, thx * znear
    bottom, left = -top, -right
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def quat_to_rot(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


@dataclass
class Camera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # (4,4) float32 row-vector convention
    projmatrix: np.ndarray   # (4,4) float32 = view @ proj
    campos: np.ndarray       # (3,)
    bg: np.ndarray           # (3,)


def make_camera(W, H, focal_ratio=1.2, R=None, t=None, bg=(0.0, 0.0, 0.0)):
    """Pinhole camera with f = focal_ratio * W.  R (3x3 world->cam rotation) and
    t (cam translation, x_cam = R x_world + t) default to identity / zero."""
    f = focal_ratio * W
    fovx = 2 * math.atan(W / (2 * f))
    fovy = 2 * math.atan(H / (2 * f))
    W2C = np.eye(4, dtype=np.float64)
    if R is not None:
        W2C[:3, :3] = R
    if t is not None:
        W2C[:3, 3] = t
    view = np.float32(W2C).T.copy()                      # world_view_transform
    proj = projection_matrix(0.01, 100.0, fovx, fovy).T  # projection_matrix (transposed)
    full = (view @ proj).astype(np.float32)
    campos = np.linalg.inv(view.astype(np.float64))[3, :3].astype(np.float32)
    return Camera(W, H, math.tan(fovx * 0.5), math.tan(fovy * 0.5), view, full, campos,
                  np.asarray(bg, dtype=np.float32))


@dataclass
class Scene:
    cam: Camera
    means3D: np.ndarray      # (P,3)
    scales: np.ndarray       # (P,2) surfel / (P,3) gaussian
    rotations: np.ndarray    # (P,4) (w,x,y,z), unit norm
    opacities: np.ndarray    # (P,1)
    colors: np.ndarray | None = None   # (P,3) precomputed colours
    shs: np.ndarray | None = None      # (P,16,3)
    sh_degree: int = 0
    extras: dict = field(default_factory=dict)

    @property
    def P(self):
        return self.means3D.shape[0]


def make_scene(P, W, H, seed=0, sh=False, scale_dims=2, sigma_px=None, opacity_sigma=1.5,
               z_range=(2.0, 20.0), bg=(0.0, 0.0, 0.0), rotate_camera=False,
               behind_fraction=0.0):
    """SURVEY 8(d) distribution: z ~ U(2,20), x,y ~ U(-1.05,1.05)*z*tanfov,
    log-scale ~ N(log(0.0014*z*(1600/W-normalised)), 0.5^2), unit quats,
    opacity = sigmoid(N(0,1.5^2)), colours U(0,1) or SH N(0,0.3^2) with DC = RGB2SH(U(0,1)).

    ``sigma_px`` overrides the projected std-dev in pixels (default: the survey's
    0.0014*z world scale, i.e. ~2.7 px at W = 1600)."""
    rng = np.random.default_rng(seed)
    Rm = tv = None
    if rotate_camera:
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        q = np.array([1.0, 0.05 * q[1], 0.05 * q[2], 0.05 * q[3]]); q /= np.linalg.norm(q)
        Rm = quat_to_rot(q)
        tv = np.array([0.1, -0.05, 0.2])
    cam = make_camera(W, H, R=Rm, t=tv, bg=bg)
    z = rng.uniform(z_range[0], z_range[1], size=P)
    x = rng.uniform(-1.05, 1.05, size=P) * z * cam.tanfovx
    y = rng.uniform(-1.05, 1.05, size=P) * z * cam.tanfovy
    if behind_fraction > 0:
        nb = int(P * behind_fraction)
        z[:nb] = rng.uniform(-1.0, 0.25, size=nb)
    pc = np.stack([x, y, z], axis=1)  # camera-space
    if Rm is not None:
        pw = (pc - tv[None, :]) @ Rm  # x_world = R^T (x_cam - t)
    else:
        pw = pc
    f = 1.2 * W
    if sigma_px is None:
        base = 0.0014 * np.abs(z) + 1e-4
    else:
        base = sigma_px * (np.abs(z) + 1e-2) / f
    logs = np.log(base)[:, None] + rng.normal(0.0, 0.5, size=(P, scale_dims))
    scales = np.exp(logs)
    q = rng.normal(size=(P, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    op = 1.0 / (1.0 + np.exp(-rng.normal(0.0, opacity_sigma, size=(P, 1))))
    sc = Scene(cam, pw.astype(np.float32), scales.astype(np.float32), q.astype(np.float32),
               op.astype(np.float32))
    if sh:
        shs = rng.normal(0.0, 0.3, size=(P, 16, 3))
        shs[:, 0, :] = (rng.uniform(0, 1, size=(P, 3)) - 0.5) / C0
        sc.shs = shs.astype(np.float32)
        sc.sh_degree = 3
    else:
        sc.colors = rng.uniform(0, 1, size=(P, 3)).astype(np.float32)
    return sc


def make_upstream_grads(W, H, seed=1, n_others=11, zero_from=7, n_color=3):
    """dL/dcolor ~ N(0,1)/N_pix; dL/dothers same for channels < zero_from, zero above
    (GS-SR never consumes allmap[7:11], twodgs_scene.py:88-105)."""
    rng = np.random.default_rng(seed)
    n = W * H
    g_color = (rng.normal(size=(n_color, H, W)) / n).astype(np.float32)
    g_others = (rng.normal(size=(n_others, H, W)) / n).astype(np.float32)
    g_others[zero_from:] = 0
    return g_color, g_others


def make_all_map(scene):
    """Per-Gaussian PGSR input [n_view(3), 1.0, |n_view . p_view|] as built by the caller
    (/root/reference/gssr/scene/pgsr_scene.py:241-302): normal = smallest-scale rotation axis,
    flipped to face the camera, expressed in the view frame (row-vector convention)."""
    view = scene.cam.viewmatrix.astype(np.float64)
    q = scene.rotations.astype(np.float64)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    Rm = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                   np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                   np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], 1)
    k = scene.scales.argmin(axis=1)
    n = Rm[np.arange(scene.P), :, k]
    to_cam = scene.cam.campos.astype(np.float64)[None] - scene.means3D
    n[(n * to_cam).sum(-1) < 0] *= -1
    ln = n @ view[:3, :3]
    pc = scene.means3D.astype(np.float64) @ view[:3, :3] + view[3, :3]
    am = np.zeros((scene.P, 5), np.float32)
    am[:, :3] = ln
    am[:, 3] = 1.0
    am[:, 4] = np.abs((ln * pc).sum(-1))
    return am


def make_points(P, seed=0, clustered=False):
    """Point clouds for distCUDA2: uniform box or a mixture of tight clusters + duplicates."""
    rng = np.random.default_rng(seed)
    if not clustered:
        return rng.uniform(-3, 3, size=(P, 3)).astype(np.float32)
    k = max(4, P // 200)
    cen = rng.uniform(-5, 5, size=(k, 3))
    idx = rng.integers(0, k, size=P)
    pts = cen[idx] + rng.normal(0, 0.05, size=(P, 3)) * rng.uniform(0.1, 3.0, size=(k, 1))[idx]
    pts = pts.astype(np.float32)
    nd = max(1, P // 50)
    pts[rng.integers(0, P, nd)] = pts[rng.integers(0, P, nd)]   # exact duplicates (distance 0)
    return pts
