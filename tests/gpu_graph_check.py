"""Developer helper (not a pytest): eager vs CUDA-graph iterations/s of the 2DGS harness iteration (config 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from train_harness import measure_iters_per_s, measure_graph_iters_per_s
P, W, H = 100_000, 800, 800
v, _ = measure_iters_per_s("ours", P, W, H, iters=40, warmup=5, fused_ssim=True, fused_post=True)
print(f"eager, fused ops:      {v:8.1f} it/s")
g, loss, ov = measure_graph_iters_per_s(P, W, H, iters=200, warmup=5)
print(f"one CUDA graph/iter:   {g:8.1f} it/s  (loss {loss:.5f}, capture_overflow {ov})")
try:
    r, _ = measure_iters_per_s("reference", P, W, H, iters=40, warmup=5)
    print(f"reference kernels:     {r:8.1f} it/s")
except Exception as e:
    print("reference arm unavailable:", e)
