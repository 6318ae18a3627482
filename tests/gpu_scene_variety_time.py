"""Developer helper (not a pytest): fwd+bwd time of the surfel rasterizer against the reference build on scenes that are
NOT the uniform benchmark scene: screen-space clustering, opaque splats (early termination), a wide size distribution,
half of the splats behind the camera.  Looks for performance cliffs outside the benchmark's comfort zone."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz, synth
import torch
from oracle import refcuda
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer, last_num_rendered
P, W, H = 1_000_000, 1600, 1060
gc, go = synth.make_upstream_grads(W, H, seed=3)
gct, got = torch.from_numpy(gc).cuda(), torch.from_numpy(go).cuda()


def ev(fn, n=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def variants():
    rng = np.random.default_rng(9)
    sc = synth.make_scene(P, W, H, seed=21); yield "uniform (benchmark-like)", sc
    sc = synth.make_scene(P, W, H, seed=21)
    z = sc.means3D[:, 2]
    sc.means3D[:, 0] = (rng.normal(0, 0.12, P) * z).astype(np.float32); sc.means3D[:, 1] = (rng.normal(0, 0.08, P) * z).astype(np.float32)
    yield "clustered in the image centre", sc
    sc = synth.make_scene(P, W, H, seed=21); sc.opacities[:] = rng.uniform(0.85, 0.99, (P, 1)).astype(np.float32)
    yield "opaque (early termination)", sc
    sc = synth.make_scene(P, W, H, seed=21); sc.scales *= np.exp(rng.normal(0, 1.0, (P, 1))).astype(np.float32)
    yield "wide size distribution", sc
    sc = synth.make_scene(P, W, H, seed=21); sc.means3D[::2, 2] *= -1
    yield "half behind the camera", sc


for name, sc in variants():
    tt = hz.to_torch(sc)
    rast = GaussianRasterizer(GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"], tt["proj"], 0,
                                                            tt["campos"], False, False))
    leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "colors")}

    def ours():
        for v in leaves.values(): v.grad = None
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        c, r, o = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], colors_precomp=leaves["colors"],
                       scales=leaves["scales"], rotations=leaves["rotations"])
        torch.autograd.backward([c, o], [gct, got])
    R = refcuda.RefSurfel()

    def ref():
        R.forward(tt["bg"], tt["view"], tt["proj"], tt["campos"], W, H, sc.cam.tanfovx, sc.cam.tanfovy, tt["means3D"], tt["opacities"],
                  tt["scales"], tt["rotations"], colors=tt["colors"])
        R.backward(gct, got)
    a = ev(ours); nr = last_num_rendered(); b = ev(ref, 4, 2)
    print(f"{name:32s} R={nr/1e6:6.2f}M  ours {a:7.3f} ms  reference {b:7.3f} ms  x{b/a:.2f}", flush=True)
