"""Poisoned Gaussians for the robustness tests: a few entries of one input array replaced by NaN / Inf / zero / huge / negative
values -- what a diverging optimisation hands the rasterizer."""
import numpy as np

KINDS = ("nan_mean", "inf_mean", "huge_mean", "zero_scale", "huge_scale", "nan_scale", "neg_scale", "zero_quat", "nan_quat",
         "opacity_edges", "nan_color", "at_camera")


def poison(sc, kind, rng, dims):
    P = sc.means3D.shape[0]
    idx = rng.choice(P, 24, replace=False)
    if kind == "nan_mean":
        sc.means3D[idx[:8], rng.integers(0, 3, 8)] = np.nan
    elif kind == "inf_mean":
        sc.means3D[idx[:8], rng.integers(0, 3, 8)] = np.inf * rng.choice([-1, 1], 8)
    elif kind == "huge_mean":
        sc.means3D[idx[:8]] *= 1e30
    elif kind == "zero_scale":
        sc.scales[idx[:8]] = 0.0
        sc.scales[idx[8:16], 0] = 0.0
    elif kind == "huge_scale":
        sc.scales[idx[:8]] = 1e18
        sc.scales[idx[8:12]] = 1e38
    elif kind == "nan_scale":
        sc.scales[idx[:8], 0] = np.nan
    elif kind == "neg_scale":
        sc.scales[idx[:8]] *= -1.0
    elif kind == "zero_quat":
        sc.rotations[idx[:8]] = 0.0
    elif kind == "nan_quat":
        sc.rotations[idx[:8], 1] = np.nan
    elif kind == "opacity_edges":
        sc.opacities[idx[:8]] = 0.0
        sc.opacities[idx[8:16]] = 1.0
        sc.opacities[idx[16:20]] = np.nan
        sc.opacities[idx[20:]] = -0.5
    elif kind == "nan_color":
        sc.colors[idx[:8], 0] = np.nan
    elif kind == "at_camera":
        sc.means3D[idx[:8]] = 0.0
        sc.means3D[idx[8:16], 2] = 0.2
    return idx
