"""Developer helper (not a pytest): kernel times on a SURFACE-LIKE depth distribution -- 95 % of the splats in a thin
depth shell (a wall seen frontally), 5 % spread over the frustum -- against the uniform-depth benchmark scene.  The
per-tile sort buckets keys linearly between the tile's min and max depth; a shell puts most keys of a tile into a few
buckets."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz, synth
import torch, gsr_b200
L = gsr_b200.lib()
P, W, H = 2_000_000, 1600, 1060
gc, go = synth.make_upstream_grads(W, H, seed=3)
names = ["pre_fwd", "scan", "dup", "sort", "build", "render_fwd", "render_bwd", "pre_bwd"]
for label, thick in (("uniform z in [2,20]", None), ("shell z = 6 +- 0.3", 0.3), ("shell z = 6 +- 0.02", 0.02), ("shell z = 6 +- 0.002", 0.002)):
    sc = synth.make_scene(P, W, H, seed=0)
    if thick is not None:
        rng = np.random.default_rng(5)
        z = sc.means3D[:, 2].astype(np.float64)
        znew = np.where(rng.random(P) < 0.95, 6.0 + rng.normal(0.0, thick, P), z)
        k = (znew / z).astype(np.float32)
        sc.means3D[:, 0] *= k; sc.means3D[:, 1] *= k; sc.means3D[:, 2] = znew.astype(np.float32)
        sc.scales *= k[:, None]                     # same screen-space footprint
    tt = hz.to_torch(sc)
    for _ in range(2):
        hz.run_product_surfel(sc, gc, go, tt=tt)
    acc = np.zeros(16)
    for _ in range(3):
        L.gsr_profile_enable(1)
        hz.run_product_surfel(sc, gc, go, tt=tt)
        buf = (ctypes.c_float * 16)(); L.gsr_profile_read(buf); acc += np.array(list(buf))
    L.gsr_profile_enable(0)
    from diff_surfel_rasterization import last_num_rendered
    print(f"{label:24s} R={last_num_rendered()/1e6:.2f}M: " + " ".join(f"{k}={acc[i]/3*1e3:.0f}" for i, k in enumerate(names) if acc[i] >= 0), flush=True)
