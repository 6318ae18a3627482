"""Developer helper (not a pytest): device times (CUDA events, fwd+bwd, inputs resident) of every variant
against the reference CUDA build (oracle/_ref) on the SURVEY 8(d) configs:
  cfg-A  surfel, P=100k, 800x800, SH degree 3        cfg-3  visible_filter on 2M anchors
  cfg-4  plane (PGSR), P=1M, 1600x900, render_geo    3DGS   P=1M, 1600x900          knn  P=1M / 2M
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz, synth
import torch
from oracle import refcuda


def ev_time(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def surfel_cfgA():
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    P, W, H = 100_000, 800, 800
    sc = synth.make_scene(P, W, H, seed=0, sh=True)
    gc, go = synth.make_upstream_grads(W, H)
    tt = hz.to_torch(sc); gct, got = torch.from_numpy(gc).cuda(), torch.from_numpy(go).cuda()
    rast = GaussianRasterizer(GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"], tt["proj"], sc.sh_degree, tt["campos"], False, False))
    leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}

    def ours():
        for v in leaves.values(): v.grad = None
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        c, r, o = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        torch.autograd.backward([c, o], [gct, got])
    R = refcuda.RefSurfel()

    def ref():
        R.forward(tt["bg"], tt["view"], tt["proj"], tt["campos"], W, H, sc.cam.tanfovx, sc.cam.tanfovy, tt["means3D"], tt["opacities"], tt["scales"], tt["rotations"], shs=tt["shs"], sh_degree=sc.sh_degree)
        R.backward(gct, got)
    a, b = ev_time(ours), ev_time(ref)
    print(f"cfg-A surfel P=100k 800x800 SH3: ours {a:.3f} ms  reference {b:.3f} ms  x{b/a:.2f}  ({P/a/1e3:.1f} vs {P/b/1e3:.1f} M Gaussians/s)")


def ewa(plane, P=1_000_000, W=1600, H=900):
    sc = synth.make_scene(P, W, H, seed=11, scale_dims=3)
    gc, go = synth.make_upstream_grads(W, H, seed=12, n_others=6, zero_from=6)
    tt = hz.to_torch(sc); gct = torch.from_numpy(gc).cuda()
    gam, gpd = torch.from_numpy(np.ascontiguousarray(go[:5])).cuda(), torch.from_numpy(np.ascontiguousarray(go[5:6])).cuda()
    kw = dict(image_height=H, image_width=W, tanfovx=sc.cam.tanfovx, tanfovy=sc.cam.tanfovy, bg=tt["bg"], scale_modifier=1.0, viewmatrix=tt["view"], projmatrix=tt["proj"], sh_degree=0, campos=tt["campos"], prefiltered=False, debug=False)
    leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "colors")}
    if plane:
        from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        kw["render_geo"] = True
        am = torch.from_numpy(synth.make_all_map(sc)).cuda().requires_grad_(True)
    else:
        from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    rast = GaussianRasterizer(GaussianRasterizationSettings(**kw))

    def ours():
        for v in leaves.values(): v.grad = None
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        common = dict(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], colors_precomp=leaves["colors"], scales=leaves["scales"], rotations=leaves["rotations"])
        if plane:
            am.grad = None
            m2a = torch.zeros_like(leaves["means3D"], requires_grad=True)
            c, r, ob, oam, pd = rast(means2D_abs=m2a, all_map=am, **common)
            torch.autograd.backward([c, oam, pd], [gct, gam, gpd])
        else:
            c, r = rast(**common)
            torch.autograd.backward([c], [gct])
    R = refcuda.RefGauss(plane=plane)

    def ref():
        R.forward(tt["bg"], tt["view"], tt["proj"], tt["campos"], W, H, sc.cam.tanfovx, sc.cam.tanfovy, tt["means3D"], tt["opacities"], tt["scales"], tt["rotations"], colors=tt["colors"], all_map=am.detach() if plane else None)
        if plane: R.backward(gct, gam, gpd)
        else: R.backward(gct)
    a, b = ev_time(ours, 10), ev_time(ref, 5)
    name = "cfg-4 plane (PGSR, render_geo)" if plane else "3DGS"
    print(f"{name} P={P} {W}x{H}: ours {a:.3f} ms  reference {b:.3f} ms  x{b/a:.2f}  ({P/a/1e3:.1f} vs {P/b/1e3:.1f} M Gaussians/s)")
    import ctypes, gsr_b200
    L = gsr_b200.lib(); L.gsr_profile_enable(1); ours(); buf = (ctypes.c_float * 16)(); L.gsr_profile_read(buf); L.gsr_profile_enable(0)
    names = ["preprocess_fwd", "scan", "duplicate", "sort", "build_records", "render_fwd", "render_bwd", "preprocess_bwd"]
    print("   kernels: " + " ".join(f"{n}={buf[i]*1e3:.0f}us" for i, n in enumerate(names) if buf[i] >= 0))


def companions():
    from simple_knn._C import distCUDA2
    for n in (1_000_000, 2_000_000):
        pts = torch.from_numpy(synth.make_points(n, seed=5)).cuda()
        a, b = ev_time(lambda: distCUDA2(pts), 5), ev_time(lambda: refcuda.ref_dist2_knn3(pts), 3)
        print(f"distCUDA2 P={n}: ours {a:.2f} ms reference {b:.2f} ms x{b/a:.1f}")
    from scaffold_filter import GaussianRasterizationSettings, GaussianRasterizer
    sc = synth.make_scene(2_000_000, 1600, 1060, seed=3, scale_dims=3)
    tt = hz.to_torch(sc)
    rast = GaussianRasterizer(GaussianRasterizationSettings(sc.cam.H, sc.cam.W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"], tt["proj"], 0, tt["campos"], False, False))
    a = ev_time(lambda: rast.visible_filter(tt["means3D"], tt["scales"], tt["rotations"]), 20)
    b = ev_time(lambda: refcuda.ref_visible_filter(tt["means3D"], tt["scales"], tt["rotations"], tt["view"], tt["proj"], sc.cam.W, sc.cam.H, sc.cam.tanfovx, sc.cam.tanfovy), 10)
    print(f"visible_filter 2M anchors: ours {a*1e3:.0f} us ({2e6*44/a/1e6:.0f} GB/s of 44 B/anchor) reference {b*1e3:.0f} us x{b/a:.1f}")


if __name__ == "__main__":
    surfel_cfgA(); ewa(False); ewa(True); companions()
