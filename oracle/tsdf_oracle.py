"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float32) of GS-SR's multi-view TSDF / colour fusion
on sample points.  Only tests/, __graft_entry__.smoke() and bench-style helpers may import this; the
product path (gs-sr_b200/csrc/tsdf.cu) never does.

Follows /root/reference/gssr/utils/mesh_utils.py:
  contract / uncontract            :187-193
  compute_sdf_perframe             :195-207
  compute_unbounded_tsdf           :209-246
  normalize / unnormalize / inv_contraction  :248-250
and torch.nn.functional.grid_sample(mode='bilinear', padding_mode='border', align_corners=True)
(ATen GridSampler: unnormalize ((x+1)/2)*(size-1), clip to [0, size-1], 4 taps nw/ne/sw/se, taps outside
the image dropped).

Parity pinned: tests/golden/tsdf_*.npz hold the outputs of the reference's own nested functions, executed
verbatim (extracted from the reference source by tests/golden/make_golden_tsdf.py) with CPU torch.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def grid_sample_border(img, gx, gy):
    """img (H,W) float32, gx/gy (n,) normalised coords in [-1,1] -> (n,) float32."""
    H, W = img.shape
    ix = ((gx + F(1)) / F(2)) * F(W - 1)
    iy = ((gy + F(1)) / F(2)) * F(H - 1)
    ix = np.minimum(F(W - 1), np.maximum(ix, F(0))).astype(F)
    iy = np.minimum(F(H - 1), np.maximum(iy, F(0))).astype(F)
    x0f, y0f = np.floor(ix), np.floor(iy)
    x1f, y1f = x0f + F(1), y0f + F(1)
    nw = (x1f - ix) * (y1f - iy)
    ne = (ix - x0f) * (y1f - iy)
    sw = (x1f - ix) * (iy - y0f)
    se = (ix - x0f) * (iy - y0f)
    x0, y0 = x0f.astype(np.int64), y0f.astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    x1_in, y1_in = x1 < W, y1 < H
    x1c, y1c = np.minimum(x1, W - 1), np.minimum(y1, H - 1)
    out = img[y0, x0] * nw
    out = out + np.where(x1_in, img[y0, x1c] * ne, F(0))
    out = out + np.where(y1_in, img[y1c, x0] * sw, F(0))
    out = out + np.where(x1_in & y1_in, img[y1c, x1c] * se, F(0))
    return out.astype(F)


def uncontract(y):
    mag = np.linalg.norm(y, axis=-1, keepdims=True).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        far = (F(1) / (F(2) - mag)) * (y / mag)
    return np.where(mag < 1, y, far).astype(F)


def compute_unbounded_tsdf(samples, contracted, center, radius, voxel_size, projs, depthmaps, rgbmaps=None,
                           return_rgb=False):
    """samples (n,3); projs: list of (4,4) full_proj_transform (row-vector convention); depthmaps: list of
    (H,W); rgbmaps: list of (3,H,W).  Returns tsdf (n,) [, rgb (n,3)] exactly as mesh_utils.py:209-246."""
    samples = np.asarray(samples, dtype=F)
    n = samples.shape[0]
    if contracted:
        norm = np.linalg.norm(samples, axis=-1).astype(F)
        mask = norm > 1
        sdf_trunc = (F(5) * F(voxel_size)) * np.ones(n, dtype=F)
        sdf_trunc[mask] *= (F(1) / (F(2) - np.minimum(norm[mask], F(1.9)))).astype(F)
        samples = (uncontract(samples) * F(radius) + np.asarray(center, dtype=F)).astype(F)
    else:
        sdf_trunc = (F(5) * F(voxel_size)) * np.ones(n, dtype=F)
    tsdfs = np.ones(n, dtype=F)
    rgbs = np.zeros((n, 3), dtype=F)
    weights = np.ones(n, dtype=F)
    hom = np.concatenate([samples, np.ones((n, 1), dtype=F)], axis=-1)
    for v, M in enumerate(projs):
        M = np.asarray(M, dtype=F)
        new_points = (hom @ M).astype(F)
        z = new_points[:, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            pix = (new_points[:, :2] / z[:, None]).astype(F)
        mask_proj = (pix > -1).all(-1) & (pix < 1).all(-1) & (z > 0)
        gx = np.where(mask_proj, pix[:, 0], F(0)).astype(F)
        gy = np.where(mask_proj, pix[:, 1], F(0)).astype(F)
        sdf = grid_sample_border(np.asarray(depthmaps[v], dtype=F).reshape(depthmaps[v].shape[-2:]), gx, gy) - z
        mask = mask_proj & (sdf > -sdf_trunc)
        s = np.clip(sdf / sdf_trunc, F(-1), F(1)).astype(F)
        w = weights[mask]
        wp = w + F(1)
        tsdfs[mask] = (tsdfs[mask] * w + s[mask]) / wp
        if return_rgb:
            rgb = np.stack([grid_sample_border(np.asarray(rgbmaps[v][c], dtype=F), gx, gy) for c in range(3)], -1)
            rgbs[mask] = (rgbs[mask] * w[:, None] + rgb[mask]) / wp[:, None]
        weights[mask] = wp
    if return_rgb:
        return tsdfs, rgbs
    return tsdfs


def integrate_grid(origin, voxel_size, dims, sdf_trunc, depth_trunc, projs, depthmaps, rgbmaps=None, state=None):
    """Bounded volume (extract_mesh_bounded, mesh_utils.py:138-179, with the torch rule above standing in for Open3D's
    ScalableTSDFVolume -- parity unpinned against Open3D): lattice origin + (ix, iy, iz) * voxel_size, x fastest; fixed
    truncation; a view is skipped where its sampled depth is <= 0 (:160-162) or > depth_trunc (:165-170).
    Returns (tsdf, weight, rgb) shaped (nz, ny, nx[, 3]); `state` continues an earlier call."""
    nx, ny, nz = dims
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    o = np.asarray(origin, dtype=F)
    pts = np.stack([ix.ravel().astype(F) * F(voxel_size) + o[0], iy.ravel().astype(F) * F(voxel_size) + o[1],
                    iz.ravel().astype(F) * F(voxel_size) + o[2]], axis=-1).astype(F)
    n = pts.shape[0]
    if state is None:
        tsdfs, weights, rgbs = np.ones(n, dtype=F), np.ones(n, dtype=F), np.zeros((n, 3), dtype=F)
    else:
        tsdfs, weights = state[0].reshape(-1).copy(), state[1].reshape(-1).copy()
        rgbs = state[2].reshape(-1, 3).copy() if state[2] is not None else np.zeros((n, 3), dtype=F)
    hom = np.concatenate([pts, np.ones((n, 1), dtype=F)], axis=-1)
    trunc = F(sdf_trunc)
    for v, M in enumerate(projs):
        new_points = (hom @ np.asarray(M, dtype=F)).astype(F)
        z = new_points[:, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            pix = (new_points[:, :2] / z[:, None]).astype(F)
        mask_proj = (pix > -1).all(-1) & (pix < 1).all(-1) & (z > 0)
        gx = np.where(mask_proj, pix[:, 0], F(0)).astype(F)
        gy = np.where(mask_proj, pix[:, 1], F(0)).astype(F)
        d = grid_sample_border(np.asarray(depthmaps[v], dtype=F).reshape(depthmaps[v].shape[-2:]), gx, gy)
        sdf = d - z
        mask = mask_proj & (d > 0) & (d <= F(depth_trunc)) & (sdf > -trunc)
        s = np.clip(sdf / trunc, F(-1), F(1)).astype(F)
        w = weights[mask]
        wp = w + F(1)
        tsdfs[mask] = (tsdfs[mask] * w + s[mask]) / wp
        if rgbmaps is not None:
            rgb = np.stack([grid_sample_border(np.asarray(rgbmaps[v][c], dtype=F), gx, gy) for c in range(3)], -1)
            rgbs[mask] = (rgbs[mask] * w[:, None] + rgb[mask]) / wp[:, None]
        weights[mask] = wp
    shape = (nz, ny, nx)
    return tsdfs.reshape(shape), weights.reshape(shape), (rgbs.reshape(shape + (3,)) if rgbmaps is not None else None)
