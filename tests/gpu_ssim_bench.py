"""Developer helper (not a pytest): fused SSIM fwd+bwd vs the reference's conv2d formulation on the GPU, 3x1060x1600."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa
import torch
from gsr_b200.ssim import ssim
from oracle import ssim_oracle
x = torch.rand((3, 1060, 1600), device="cuda", requires_grad=True); y = torch.rand((3, 1060, 1600), device="cuda")


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ours():
    x.grad = None; ssim(x, y).backward()


def ref():
    x.grad = None; ssim_oracle.ssim(x, y).backward()


a, b = t(ours), t(ref)
px = 3 * 1060 * 1600
print(f"SSIM fwd+bwd 3x1060x1600: fused {a*1e3:.0f} us ({px*84/a/1e6:.0f} GB/s of the 84 B/pixel-channel algorithmic bytes), conv2d formulation {b*1e3:.0f} us, x{b/a:.1f}")
