"""Seeded image pairs for the SSIM tests (shared by the golden generator and the tests)."""
import numpy as np

SSIM_CASES = {"small": (3, 45, 70, 401), "ragged": (3, 33, 97, 402), "one_channel": (1, 64, 64, 403), "tiny": (3, 7, 9, 404)}


def build_ssim_case(name):
    C, H, W, seed = SSIM_CASES[name]
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(H) / H, np.arange(W) / W, indexing="ij")
    base = np.stack([0.5 + 0.4 * np.sin(6 * xx + 2 * yy + c) for c in range(C)]).astype(np.float32)
    img2 = np.clip(base + 0.05 * rng.normal(size=base.shape), 0, 1).astype(np.float32)
    img1 = np.clip(base + 0.15 * rng.normal(size=base.shape), 0, 1).astype(np.float32)
    img1[:, : H // 3] = 0.0                     # a flat region (sigma = 0: exercises the C1/C2 terms)
    return img1, img2
