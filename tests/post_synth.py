"""Seeded allmaps / cameras / upstream gradients for the surfel post-processing tests."""
import numpy as np

import synth

POST_CASES = {"basic": (70, 45, 0.0, 501), "median": (64, 48, 1.0, 502), "mixed": (97, 33, 0.3, 503)}


def build_post_case(name, W=None, H=None):
    w0, h0, ratio, seed = POST_CASES[name]
    W, H = W or w0, H or h0
    rng = np.random.default_rng(seed)
    cam = synth.make_camera(W, H, R=synth.quat_to_rot(np.array([0.98, 0.1, -0.12, 0.08]) / np.linalg.norm([0.98, 0.1, -0.12, 0.08])),
                            t=np.array([0.2, -0.1, 0.4]))
    yy, xx = np.meshgrid(np.arange(H) / H, np.arange(W) / W, indexing="ij")
    alpha = np.clip(0.6 + 0.5 * np.sin(5 * xx + 3 * yy) + 0.05 * rng.normal(size=(H, W)), 0.0, 1.0)
    alpha[: H // 5, : W // 4] = 0.0                                  # empty region: 0/0 in the expected depth
    depth = 3.0 + np.sin(4 * xx) + 0.5 * np.cos(6 * yy) + 0.02 * rng.normal(size=(H, W))
    nrm = rng.normal(size=(3, H, W)); nrm /= np.linalg.norm(nrm, axis=0, keepdims=True)
    allmap = np.zeros((11, H, W), dtype=np.float32)
    allmap[0] = depth * alpha
    allmap[1] = alpha
    allmap[2:5] = nrm * alpha
    allmap[5] = np.where(alpha > 0, depth + 0.05 * rng.normal(size=(H, W)), 0.0)
    allmap[6] = 0.01 * rng.uniform(size=(H, W))
    g = {k: rng.normal(size=s).astype(np.float32) for k, s in (("normal", (3, H, W)), ("depth", (1, H, W)), ("surf_normal", (3, H, W)))}
    return dict(allmap=allmap, wvt=cam.viewmatrix, full_proj=cam.projmatrix, depth_ratio=ratio, W=W, H=H, g=g)
