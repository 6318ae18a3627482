"""Developer helper (not a pytest): kernel times with ONE heavy tile (P_heavy splats stacked on it) among ordinary ones --
the cost of the per-tile sort's large-bucket paths (tile_sort.cuh) and of the render kernels' longest list."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz, synth
from test_ring_stress_gpu import heavy_scene
import torch, gsr_b200
L = gsr_b200.lib()
W, H = 640, 480
gc, go = synth.make_upstream_grads(W, H, seed=3)
names = ["pre_fwd", "scan", "dup", "sort", "build", "render_fwd", "render_bwd", "pre_bwd"]
for P_heavy in (0, 1500, 3000, 6000, 12000, 30000, 100000):
    sc = heavy_scene(max(P_heavy, 1), 150_000, W, H, seed=7)
    tt = hz.to_torch(sc)
    for _ in range(3):
        hz.run_product_surfel(sc, gc, go, tt=tt)
    acc = np.zeros(16)
    for _ in range(5):
        L.gsr_profile_enable(1)
        hz.run_product_surfel(sc, gc, go, tt=tt)
        buf = (ctypes.c_float * 16)(); L.gsr_profile_read(buf); acc += np.array(list(buf))
    L.gsr_profile_enable(0)
    print(f"P_heavy={P_heavy:6d}: " + " ".join(f"{k}={acc[i]/5*1e3:.0f}" for i, k in enumerate(names) if acc[i] >= 0), flush=True)
