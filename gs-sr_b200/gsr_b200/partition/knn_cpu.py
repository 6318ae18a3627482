"""CPU distCUDA2: mean squared distance to the 3 nearest neighbours, so that BASELINE config 1 (VastGaussian partition of
a COLMAP model) runs without a GPU -- SURVEY.md section 8(f4).

The reference partitioner calls the CUDA extension for this (/root/reference/gssr/utils/vastgaussian_utils.py:12,225-226 ->
submodules/simple-knn/simple_knn.cu:132-222).  This is a separate, explicitly CPU function (a k-d tree for the candidate
search, scipy.spatial.cKDTree); ``simple_knn._C.distCUDA2`` itself stays GPU-only and never falls back to it.

Result contract (K/simple_knn.cu:186-222): for every point the three smallest squared distances to OTHER points (an exact
duplicate counts, at distance 0), each formed in float32 as fma(dz, dz, fma(dx, dx, dy*dy)), summed smallest first and
divided by 3 -- the same arithmetic as csrc/knn.cu, so the two agree to float32 rounding (tests/test_partition_cpu.py
compares with the golden vectors of the reference kernel).
"""
from __future__ import annotations

import numpy as np


def _d2_f32(a, b):
    """float32 squared distance with the GPU kernels' operation order (fused multiply-adds emulated in float64:
    a product of two float32 is exact in float64)."""
    d = (a - b).astype(np.float32).astype(np.float64)
    t = np.float32(d[..., 1] * d[..., 1]).astype(np.float64)
    t = np.float32(d[..., 0] * d[..., 0] + t).astype(np.float64)
    return np.float32(d[..., 2] * d[..., 2] + t)


def dist2_knn3_cpu(points, candidates=8):
    """points (P,3) float32 -> (P,) float32 mean of the 3 smallest squared neighbour distances."""
    from scipy.spatial import cKDTree
    pts = np.ascontiguousarray(points, dtype=np.float32)
    P = pts.shape[0]
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError("points must have shape (P, 3)")
    if P == 0:
        return np.zeros((0,), np.float32)
    if P < 4:
        raise ValueError("dist2_knn3_cpu needs at least 4 points (3 neighbours per point)")
    k = min(max(candidates, 4), P)
    tree = cKDTree(pts.astype(np.float64))
    _, idx = tree.query(pts.astype(np.float64), k=k, workers=-1)      # all host cores; per-point results do not depend on it
    d2 = _d2_f32(pts[:, None, :], pts[idx])                         # (P,k) float32, GPU arithmetic
    # drop the query point itself: its own index when the tree returned it, else (more than k coincident points) any zero
    is_self = idx == np.arange(P)[:, None]
    none = ~is_self.any(axis=1)
    is_self[none, 0] = True
    d2 = np.where(is_self, np.float32(np.inf), d2)
    d2.sort(axis=1)
    s = (d2[:, 0] + d2[:, 1]).astype(np.float32)
    s = (s + d2[:, 2]).astype(np.float32)
    return (s / np.float32(3.0)).astype(np.float32)
