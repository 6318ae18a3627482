"""Developer helper (not a pytest): train iters/s of the GS-SR-style 2DGS iteration (tests/train_harness.py),
drop-in rasterizer vs reference kernels, cfg-A (P=100k, 800x800, SH3) and a cfg-B-sized scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa: F401  (sys.path for the drop-in packages)
from train_harness import measure_iters_per_s
from oracle import refcuda
for P, W, H, it in ((100_000, 800, 800, 40), (2_000_000, 1600, 1060, 12)):
    a, _ = measure_iters_per_s("ours", P, W, H, iters=it)
    af, _ = measure_iters_per_s("ours", P, W, H, iters=it, fused_ssim=True)
    ap, _ = measure_iters_per_s("ours", P, W, H, iters=it, fused_ssim=True, fused_post=True)
    msg = f"P={P} {W}x{H} SH3: ours {a:.1f} it/s (+ fused SSIM {af:.1f}, + fused post-processing {ap:.1f} it/s)"
    if refcuda.available("surfel"):
        b, _ = measure_iters_per_s("reference", P, W, H, iters=it)
        msg += f", reference kernels {b:.1f} it/s, x{a/b:.2f}"
    print(msg)
for N in (400_000,):
    a, _ = measure_iters_per_s("ours", N, 1600, 1060, iters=12, scaffold=True)
    af, _ = measure_iters_per_s("ours", N, 1600, 1060, iters=12, scaffold=True, fused_ssim=True, fused_post=True)
    msg = f"Scaffold-2DGS flow, {N} anchors x5 offsets, 1600x1060: ours {a:.1f} it/s (+ fused SSIM and post-processing {af:.1f} it/s)"
    if refcuda.available("surfel"):
        b, _ = measure_iters_per_s("reference", N, 1600, 1060, iters=12, scaffold=True)
        msg += f", reference kernels {b:.1f} it/s, x{a/b:.2f}"
    print(msg)
a, _ = measure_iters_per_s("ours", 1_000_000, 1600, 900, iters=12, pgsr=True)
af, _ = measure_iters_per_s("ours", 1_000_000, 1600, 900, iters=12, pgsr=True, fused_ssim=True)
msg = f"PGSR flow (2 views/iter), P=1M, 1600x900: ours {a:.1f} it/s (+ fused SSIM {af:.1f} it/s)"
if refcuda.available("plane"):
    b, _ = measure_iters_per_s("reference", 1_000_000, 1600, 900, iters=12, pgsr=True)
    msg += f", reference kernels {b:.1f} it/s, x{a/b:.2f}"
print(msg)
