"""Seeded depth maps for the PGSR depth->normal tests (plane depth of a tilted ground + bumps, a ragged size, a case with
zero / constant depth where the cross product vanishes)."""
import numpy as np

DN_CASES = ("tilted", "ragged", "flat_and_zero")


def build_dn_case(name):
    rng = np.random.default_rng({"tilted": 11, "ragged": 12, "flat_and_zero": 13}[name])
    W, H = {"tilted": (96, 64), "ragged": (53, 37), "flat_and_zero": (40, 24)}[name]
    fx = 1.2 * W
    K = np.array([[fx, 0, W / 2.0], [0, fx * 1.02, H / 2.0], [0, 0, 1]], np.float32)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    if name == "flat_and_zero":
        depth = np.full((H, W), 3.0)
        depth[:, : W // 3] = 0.0                                   # zero depth: degenerate cross product -> zero normal
        depth[H // 2:, W // 2:] += rng.normal(0, 0.05, (H - H // 2, W - W // 2))
    else:
        depth = 4.0 + 0.01 * xs - 0.02 * ys + 0.3 * np.sin(xs / 7.0) * np.cos(ys / 5.0) + rng.normal(0, 0.01, (H, W))
    g = (rng.normal(size=(3, H, W)) / (W * H)).astype(np.float32)
    weight = rng.uniform(0, 1, (H, W)).astype(np.float32)
    return dict(W=W, H=H, K=K, depth=depth.astype(np.float32), g=g, weight=weight)
