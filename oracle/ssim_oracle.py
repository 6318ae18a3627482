"""TEST INFRASTRUCTURE -- CPU restatement (torch conv2d, float32 or float64) of GS-SR's SSIM.
Follows /root/reference/gssr/scene/vanilla_scene.py: _gaussian :50-52, ssim :53-61 (window = outer product of the
normalised 1-D Gaussian, expanded per channel), _ssim :32-48 (zero-padded depthwise conv2d, C1/C2, mean).
Parity pinned: tests/golden/ssim_*.npz hold value and gradient produced by the reference's own methods, cut out of the
reference source and executed verbatim by tests/golden/make_golden_ssim.py.  Only tests/ may import this."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def window(channel, dtype=torch.float32):
    g = torch.Tensor([math.exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, 11, 11).contiguous().to(dtype)


def ssim(img1, img2):
    channel = img1.size(-3)
    w = window(channel, img1.dtype).to(img1.device)
    pad = 5
    mu1 = F.conv2d(img1, w, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, w, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, w, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, w, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, w, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def ssim_value_and_grad(img1_np, img2_np, dtype=torch.float32):
    x = torch.from_numpy(img1_np).to(dtype).requires_grad_(True)
    y = torch.from_numpy(img2_np).to(dtype)
    v = ssim(x, y)
    v.backward()
    return float(v.detach()), x.grad.numpy()
