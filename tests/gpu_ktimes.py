"""Developer helper: per-kernel device times (library cudaEvents) for cfg-B, averaged."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_profile as gp
import torch, gsr_b200
P, W, H = 2_000_000, 1600, 1060
if len(sys.argv) > 3: P, W, H = map(int, sys.argv[1:4])
sc, tt, gct, got, rast, leaves, m2d = gp.setup(P, W, H)
L = gsr_b200.lib()
names = ["preprocess_fwd", "scan", "duplicate", "sort", "build_records", "render_fwd", "render_bwd", "preprocess_bwd"]
for _ in range(3): gp.product_step(rast, leaves, m2d, gct, got)
acc = np.zeros(16); n = 5
for _ in range(n):
    L.gsr_profile_enable(1)
    gp.product_step(rast, leaves, m2d, gct, got)
    buf = (ctypes.c_float * 16)(); L.gsr_profile_read(buf); acc += np.array(list(buf))
L.gsr_profile_enable(0)
print(" ".join(f"{k}={acc[i]/n*1e3:.0f}us" for i, k in enumerate(names)), "| sum=%.3f ms" % (acc[:8].sum() / n))
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20): gp.product_step(rast, leaves, m2d, gct, got)
torch.cuda.synchronize()
print("async wall: %.3f ms/step" % ((time.perf_counter() - t0) * 50))
