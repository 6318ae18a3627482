"""Developer helper (not a pytest): fwd+bwd time of the 3DGS rasterizer (PLANE=1: the PGSR plane rasterizer with render_geo)
against the reference build on non-benchmark scenes (see gpu_scene_variety_time.py), incl. a thin depth shell and
screen-filling splats."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz, synth
import torch
from oracle import refcuda
PLANE = os.environ.get("PLANE", "0") == "1"
if PLANE:
    from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer
else:
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
P, W, H = 1_000_000, 1600, 900
gc, go = synth.make_upstream_grads(W, H, seed=3, n_others=6, zero_from=6)
gct = torch.from_numpy(gc).cuda()
gam, gpd = torch.from_numpy(np.ascontiguousarray(go[:5])).cuda(), torch.from_numpy(np.ascontiguousarray(go[5:6])).cuda()


def ev(fn, n=6, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def variants():
    rng = np.random.default_rng(9)
    mk = lambda: synth.make_scene(P, W, H, seed=21, scale_dims=3)  # noqa: E731
    yield "uniform (benchmark-like)", mk()
    sc = mk(); z = sc.means3D[:, 2].astype(np.float64)
    znew = np.where(rng.random(P) < 0.95, 6.0 + rng.normal(0.0, 0.01, P), z); k = (znew / z).astype(np.float32)
    sc.means3D[:, 0] *= k; sc.means3D[:, 1] *= k; sc.means3D[:, 2] = znew.astype(np.float32); sc.scales *= k[:, None]
    yield "thin depth shell", sc
    sc = mk(); sel = rng.choice(P, 300, replace=False)
    sc.scales[sel] = (250.0 * sc.means3D[sel, 2] / (1.2 * W))[:, None].astype(np.float32); sc.opacities[sel] = 0.05
    yield "300 screen-filling splats", sc
    sc = mk(); sc.scales *= np.exp(rng.normal(0, 1.0, (P, 1))).astype(np.float32)
    yield "wide size distribution", sc
    sc = mk(); sc.opacities[:] = rng.uniform(0.85, 0.99, (P, 1)).astype(np.float32)
    yield "opaque (early termination)", sc


for name, sc in variants():
    tt = hz.to_torch(sc)
    kw = dict(image_height=H, image_width=W, tanfovx=sc.cam.tanfovx, tanfovy=sc.cam.tanfovy, bg=tt["bg"], scale_modifier=1.0,
              viewmatrix=tt["view"], projmatrix=tt["proj"], sh_degree=0, campos=tt["campos"], prefiltered=False, debug=False)
    if PLANE:
        kw["render_geo"] = True
    rast = GaussianRasterizer(GaussianRasterizationSettings(**kw))
    am = torch.from_numpy(synth.make_all_map(sc)).cuda().requires_grad_(True) if PLANE else None
    leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "colors")}

    def ours():
        for v in leaves.values(): v.grad = None
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        common = dict(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], colors_precomp=leaves["colors"],
                      scales=leaves["scales"], rotations=leaves["rotations"])
        if PLANE:
            am.grad = None
            m2a = torch.zeros_like(leaves["means3D"], requires_grad=True)
            c, r, ob, oam, pd = rast(means2D_abs=m2a, all_map=am, **common)
            torch.autograd.backward([c, oam, pd], [gct, gam, gpd])
        else:
            c, r = rast(**common)
            torch.autograd.backward([c], [gct])
    R = refcuda.RefGauss(plane=PLANE)

    def ref():
        R.forward(tt["bg"], tt["view"], tt["proj"], tt["campos"], W, H, sc.cam.tanfovx, sc.cam.tanfovy, tt["means3D"], tt["opacities"],
                  tt["scales"], tt["rotations"], colors=tt["colors"], all_map=am.detach() if PLANE else None)
        if PLANE:
            R.backward(gct, gam, gpd)
        else:
            R.backward(gct)
    a = ev(ours); b = ev(ref, 3, 1)
    print(f"{name:30s} ours {a:7.3f} ms  reference {b:7.3f} ms  x{b/a:.2f}", flush=True)
