"""Developer helper (not a pytest): kernel times with a few GIANT splats (screen-filling background / sky Gaussians) among
ordinary ones.  Rectangles of more than 32 tiles have no tile mask: the scatter repeats the tile test for them."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz, synth
import torch, gsr_b200
L = gsr_b200.lib()
W, H, P = 1600, 1060, 500_000
gc, go = synth.make_upstream_grads(W, H, seed=3)
names = ["pre_fwd", "scan", "dup", "sort", "build", "render_fwd", "render_bwd", "pre_bwd"]
for n_giant, sig, clustered in ((0, 0, False), (50, 300.0, False), (500, 300.0, False), (5000, 60.0, False), (50000, 20.0, False),
                               (50, 300.0, True), (500, 300.0, True)):
    sc = synth.make_scene(P, W, H, seed=11)
    if n_giant:
        f = 1.2 * W
        # giants at random indices (the usual case) or in one contiguous run of indices (all in the same one or two CTAs)
        sel = np.arange(n_giant) if clustered else np.random.default_rng(1).choice(P, n_giant, replace=False)
        z = sc.means3D[sel, 2]
        sc.scales[sel] = (sig * z / f)[:, None].astype(np.float32)
        sc.opacities[sel] = 0.05
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
    print(f"giant={n_giant:6d}{' (contiguous)' if clustered else '':13s} sigma_px={sig:5.0f} R={last_num_rendered()/1e6:6.2f}M: " + " ".join(f"{k}={acc[i]/3*1e3:.0f}" for i, k in enumerate(names) if acc[i] >= 0), flush=True)
