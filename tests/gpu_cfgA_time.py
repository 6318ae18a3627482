"""Developer helper: async wall-clock and event time per fwd+bwd step at cfg-A (P=100k, 800x800, SH 3), where the
host-side bubble around num_rendered is a visible share of the step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_profile as gp
import torch
P, W, H = 100_000, 800, 800
sc, tt, gct, got, rast, leaves, m2d = gp.setup(P, W, H, sh=True)
for _ in range(10): gp.product_step(rast, leaves, m2d, gct, got)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(200): gp.product_step(rast, leaves, m2d, gct, got)
    e1.record(); torch.cuda.synchronize()
    print(f"cfg-A: {e0.elapsed_time(e1)/200:.4f} ms/step (events), {(time.perf_counter()-t0)*5:.4f} ms/step (wall)")
