"""Developer helper: NaN / Inf in the companions' inputs (distCUDA2, visible_filter, TSDF fusion, SSIM, marching cubes) and odd
image sizes through the surfel rasterizer.  Run under `timeout` (a hang is a finding) and compute-sanitizer."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs-sr_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import harness as hz
import synth
from oracle import refcuda

def stamp(msg):
    torch.cuda.synchronize(); print(msg, flush=True)

# ---- distCUDA2
from simple_knn._C import distCUDA2
rng = np.random.default_rng(3)
for case in ("nan_one", "inf_one", "nan_many", "all_nan", "huge", "all_same"):
    pts = rng.random((50000, 3), dtype=np.float32)
    if case == "nan_one": pts[17, 1] = np.nan
    if case == "inf_one": pts[17, 1] = np.inf
    if case == "nan_many": pts[::7, 0] = np.nan
    if case == "all_nan": pts[:] = np.nan
    if case == "huge": pts[5] = 3e38; pts[6] = -3e38
    if case == "all_same": pts[:] = 0.25
    t = time.time()
    d = distCUDA2(torch.from_numpy(pts).cuda())
    torch.cuda.synchronize()
    dt = time.time() - t
    msg = f"knn {case:9s} {dt*1e3:8.1f} ms finite {float(torch.isfinite(d).float().mean()):.4f}"
    if os.path.exists(os.path.join(hz.ROOT, "oracle", "_ref", "libref_knn.so")) and case in ("nan_one", "inf_one", "huge", "all_same"):
        r = refcuda.ref_dist2_knn3(torch.from_numpy(pts).cuda())
        torch.cuda.synchronize()
        fo, fr = torch.isfinite(d), torch.isfinite(r)
        both = fo & fr
        msg += f" | ref finite {float(fr.float().mean()):.4f} mask eq {bool(torch.equal(fo, fr))} bit-equal on both-finite {bool(torch.equal(d[both], r[both]))}"
    stamp(msg)

# ---- visible_filter
from scaffold_filter import GaussianRasterizationSettings as FS, GaussianRasterizer as FR
W, H, P = 320, 240, 20000
for case in ("nan_mean", "nan_scale", "huge_scale", "nan_quat"):
    sc = synth.make_scene(P, W, H, seed=11, scale_dims=3)
    if case == "nan_mean": sc.means3D[::9, 1] = np.nan
    if case == "nan_scale": sc.scales[::9, 1] = np.nan
    if case == "huge_scale": sc.scales[::9] = 1e30
    if case == "nan_quat": sc.rotations[::9, 2] = np.nan
    tt = hz.to_torch(sc)
    rs = FS(image_height=H, image_width=W, tanfovx=sc.cam.tanfovx, tanfovy=sc.cam.tanfovy, bg=tt["bg"], scale_modifier=1.0,
            viewmatrix=tt["view"], projmatrix=tt["proj"], sh_degree=0, campos=tt["campos"], prefiltered=False, debug=False)
    radii = FR(rs).visible_filter(means3D=tt["means3D"], scales=tt["scales"], rotations=tt["rotations"])
    msg = f"filter {case:10s} visible {int((radii > 0).sum())}"
    if os.path.exists(os.path.join(hz.ROOT, "oracle", "_ref", "libref_filter.so")):
        rr = refcuda.ref_visible_filter(tt["means3D"], tt["scales"], tt["rotations"], tt["view"], tt["proj"], W, H,
                                        sc.cam.tanfovx, sc.cam.tanfovy)
        msg += f" ref {int((rr > 0).sum())} equal {bool(torch.equal(radii, rr))}"
    stamp(msg)

# ---- odd image sizes through the surfel rasterizer (vs the CPU oracle)
for (w, h) in ((1, 1), (1, 37), (16, 16), (17, 15), (5, 200), (4097, 3)):
    sc = synth.make_scene(3000, w, h, seed=4, sigma_px=2.0)
    gc, go = synth.make_upstream_grads(w, h, seed=5)
    out = hz.run_product_surfel(sc, gc, go)
    orc = hz.run_oracle_surfel(sc, gc, go)
    err = np.abs(out["color"] - orc["color"]).max()
    stamp(f"surfel {w}x{h}: visible {int((out['radii']>0).sum())} color err {err:.2e} radii eq {np.array_equal(out['radii'], orc['radii'])}")

# ---- TSDF / SSIM / marching cubes with NaN maps
from gsr_b200.tsdf import TSDFFusion
from tsdf_synth import build_tsdf_case
c = build_tsdf_case("contracted", n=5000)
dm = [d.copy() for d in c["depthmaps"]]
dm[0][10:20, 10:20] = np.nan; dm[1][:] = np.inf
f = TSDFFusion([torch.from_numpy(m) for m in c["projs"]], [torch.from_numpy(d) for d in dm], None, center=c["center"], radius=c["radius"])
s = torch.from_numpy(c["samples"]).cuda(); s[::11] = float("nan")
t = f.compute_unbounded_tsdf(s, True, c["voxel_size"])
stamp(f"tsdf with NaN maps / samples: finite {float(torch.isfinite(t).float().mean()):.4f}")
from gsr_b200.ssim import ssim as fssim
a = torch.rand(3, 100, 130, device="cuda", requires_grad=True); b = torch.rand(3, 100, 130, device="cuda")
with torch.no_grad(): b[1, 5, 5] = float("nan")
v = fssim(a, b); v.backward()
stamp(f"ssim with a NaN pixel: value {float(v)} grad finite {float(torch.isfinite(a.grad).float().mean()):.4f}")
from gsr_b200.mesh import extract_triangle_mesh, post_process_mesh
g = torch.randn(40, 41, 42, device="cuda"); g[::5, ::3, ::2] = float("nan"); g[1::5, ::3, ::2] = float("inf")
m = extract_triangle_mesh(g)
p = post_process_mesh(m, cluster_to_keep=5)
stamp(f"marching cubes with NaN / Inf voxels: V {m.vertices.shape[0]} F {m.triangles.shape[0]} finite verts {float(torch.isfinite(m.vertices).all(dim=1).float().mean()):.4f} post F {p.triangles.shape[0]}")
print("done")
