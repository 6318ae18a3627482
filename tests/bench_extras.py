"""Secondary workloads of the hot path, timed for BOTH arms of bench.py (ours / the unmodified reference CUDA build) so
that every row of DESIGN.md's speed table comes out of a driver-run bench line:

  cfgA_surfel      2DGS surfel rasterizer fwd+bwd, P=100k, 800x800, SH degree 3          (BASELINE config 2)
  gauss_1m         3DGS rasterizer fwd+bwd, P=1M, 1600x900
  cfg4_plane       PGSR plane rasterizer fwd+bwd, P=1M, 1600x900, render_geo             (BASELINE config 4 stand-in)
  visible_filter   scaffold_filter.visible_filter on 2M anchors @ 1600x1060               (BASELINE config 3 prefilter)
  dist2_knn3       simple_knn distCUDA2 on 1M points
  ssim_loss        SSIM loss fwd+bwd, 3x1060x1600 (reference arm: the reference's conv2d formulation in torch, same GPU)
  tsdf_fuse        TSDF fusion, 256^3 samples x 32 views @ 1600x1060 (reference arm: the reference's per-view torch rule)
  mesh_extract     marching cubes + cluster filter of a 512^3 TSDF lattice, all on the device (reference arm: Open3D /
                   skimage are absent -- the numpy restatement oracle/mcubes_oracle.py + mesh_clusters_oracle.py on the host,
                   kind "port", on a 128^3 crop after the device->host copy the reference's unbounded path makes)

Each entry: {"ms": device time per call (CUDA events, inputs resident, median of the timed calls), "value": units/s,
"unit": ...}.  Test/bench infrastructure: imports the drop-in packages for our arm and oracle/refcuda.py (ctypes front-end
of oracle/_ref/*.so) for the reference arm.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz  # noqa: E402
import synth  # noqa: E402


def ev_times(fn, n=10, warm=3):
    """Per-call device times (ms) of n calls after `warm` warm-ups, one event pair per call."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]


def _entry(ms_list, units, unit):
    ms = float(np.median(ms_list))
    return {"ms": ms, "p10_ms": float(np.percentile(ms_list, 10)), "p90_ms": float(np.percentile(ms_list, 90)),
            "value": units / (ms * 1e-3), "unit": unit}


def surfel_cfg_a(impl):
    import torch
    P, W, H = 100_000, 800, 800
    sc = synth.make_scene(P, W, H, seed=0, sh=True)
    gc, go = synth.make_upstream_grads(W, H)
    tt = hz.to_torch(sc)
    gct, got = torch.from_numpy(gc).cuda(), torch.from_numpy(go).cuda()
    if impl == "ours":
        from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        rast = GaussianRasterizer(GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"],
                                                                tt["proj"], sc.sh_degree, tt["campos"], False, False))
        leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}

        def fn():
            for v in leaves.values():
                v.grad = None
            m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
            c, r, o = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
                           scales=leaves["scales"], rotations=leaves["rotations"])
            torch.autograd.backward([c, o], [gct, got])
    else:
        from oracle import refcuda
        R = refcuda.RefSurfel()

        def fn():
            R.forward(tt["bg"], tt["view"], tt["proj"], tt["campos"], W, H, sc.cam.tanfovx, sc.cam.tanfovy, tt["means3D"],
                      tt["opacities"], tt["scales"], tt["rotations"], shs=tt["shs"], sh_degree=sc.sh_degree)
            R.backward(gct, got)
    out = _entry(ev_times(fn, 30, 5), P, "Gaussians/s")
    if impl == "ours":
        # the same forward + backward recorded once into a CUDA graph and replayed: at this size the eager call is bound by
        # the host's launch rate, not by the GPU (the reference forward blocks on a cudaMemcpy and cannot be captured)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            for v in leaves.values():
                v.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
                c, r, o = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
                               scales=leaves["scales"], rotations=leaves["rotations"])
                torch.autograd.backward([c, o], [gct, got])
            e = _entry(ev_times(g.replay, 30, 5), P, "Gaussians/s")
            out["cuda_graph"] = {"ms": e["ms"], "value": e["value"], "unit": e["unit"]}
        except Exception as ex:  # noqa: BLE001
            out["cuda_graph"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    return out


def ewa(impl, plane, P=1_000_000, W=1600, H=900):
    import torch
    sc = synth.make_scene(P, W, H, seed=11, scale_dims=3)
    gc, go = synth.make_upstream_grads(W, H, seed=12, n_others=6, zero_from=6)
    tt = hz.to_torch(sc)
    gct = torch.from_numpy(gc).cuda()
    gam, gpd = torch.from_numpy(np.ascontiguousarray(go[:5])).cuda(), torch.from_numpy(np.ascontiguousarray(go[5:6])).cuda()
    am = torch.from_numpy(synth.make_all_map(sc)).cuda() if plane else None
    if impl == "ours":
        kw = dict(image_height=H, image_width=W, tanfovx=sc.cam.tanfovx, tanfovy=sc.cam.tanfovy, bg=tt["bg"], scale_modifier=1.0,
                  viewmatrix=tt["view"], projmatrix=tt["proj"], sh_degree=0, campos=tt["campos"], prefiltered=False, debug=False)
        leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "colors")}
        if plane:
            from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer
            kw["render_geo"] = True
            am = am.requires_grad_(True)
        else:
            from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        rast = GaussianRasterizer(GaussianRasterizationSettings(**kw))

        def fn():
            for v in leaves.values():
                v.grad = None
            m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
            common = dict(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], colors_precomp=leaves["colors"],
                          scales=leaves["scales"], rotations=leaves["rotations"])
            if plane:
                am.grad = None
                m2a = torch.zeros_like(leaves["means3D"], requires_grad=True)
                c, r, ob, oam, pd = rast(means2D_abs=m2a, all_map=am, **common)
                torch.autograd.backward([c, oam, pd], [gct, gam, gpd])
            else:
                c, r = rast(**common)
                torch.autograd.backward([c], [gct])
    else:
        from oracle import refcuda
        R = refcuda.RefGauss(plane=plane)

        def fn():
            R.forward(tt["bg"], tt["view"], tt["proj"], tt["campos"], W, H, sc.cam.tanfovx, sc.cam.tanfovy, tt["means3D"],
                      tt["opacities"], tt["scales"], tt["rotations"], colors=tt["colors"], all_map=am)
            if plane:
                R.backward(gct, gam, gpd)
            else:
                R.backward(gct)
    return _entry(ev_times(fn, 10, 3), P, "Gaussians/s")


def visible_filter(impl, P=2_000_000, W=1600, H=1060):
    sc = synth.make_scene(P, W, H, seed=3, scale_dims=3)
    tt = hz.to_torch(sc)
    if impl == "ours":
        from scaffold_filter import GaussianRasterizationSettings, GaussianRasterizer
        rast = GaussianRasterizer(GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"],
                                                                tt["proj"], 0, tt["campos"], False, False))
        fn = lambda: rast.visible_filter(tt["means3D"], tt["scales"], tt["rotations"])  # noqa: E731
    else:
        from oracle import refcuda
        fn = lambda: refcuda.ref_visible_filter(tt["means3D"], tt["scales"], tt["rotations"], tt["view"], tt["proj"], W, H,  # noqa: E731
                                                sc.cam.tanfovx, sc.cam.tanfovy)
    return _entry(ev_times(fn, 20, 3), P, "anchors/s")


def dist2_knn3(impl, P=1_000_000):
    import torch
    pts = torch.from_numpy(synth.make_points(P, seed=5)).cuda()
    if impl == "ours":
        from simple_knn._C import distCUDA2
        fn = lambda: distCUDA2(pts)  # noqa: E731
    else:
        from oracle import refcuda
        fn = lambda: refcuda.ref_dist2_knn3(pts)  # noqa: E731
    return _entry(ev_times(fn, 5, 2), P, "points/s")


def ssim_loss(impl, H=1060, W=1600):
    """SSIM loss forward + backward on a 3 x 1060 x 1600 image: the fused kernels against the reference's five 121-tap
    depthwise conv2d + autograd formulation (oracle/ssim_oracle.py restates vanilla_scene.py:32-61 with torch ops) on the
    same GPU."""
    import torch
    x = torch.rand((3, H, W), device="cuda", generator=torch.Generator("cuda").manual_seed(1)).requires_grad_(True)
    y = torch.rand((3, H, W), device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    if impl == "ours":
        from gsr_b200.ssim import ssim
    else:
        from oracle.ssim_oracle import ssim

    def fn():
        x.grad = None
        ssim(x, y).backward()
    return _entry(ev_times(fn, 20, 3), 3 * H * W, "pixel-channels/s")


def tsdf_fuse(impl, Ng=256):
    """TSDF fusion of a 256^3 sample chunk against 32 depth maps @ 1600x1060 (extract_mesh_unbounded's inner call):
    gsr_tsdf_fuse against the reference's per-view torch rule (tests/tsdf_synth.py restates mesh_utils.py:195-246 with the
    same torch ops) on the same GPU."""
    import torch
    from tsdf_synth import build_tsdf_case, torch_rule_unbounded
    c = build_tsdf_case("bench")
    ax = torch.linspace(-1.2, 1.2, Ng, device="cuda")
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
    pts = torch.stack([xx.ravel(), yy.ravel(), zz.ravel()], -1).contiguous()
    projs = [torch.from_numpy(m).cuda() for m in c["projs"]]
    depths = [torch.from_numpy(d).cuda() for d in c["depthmaps"]]
    if impl == "ours":
        from gsr_b200.tsdf import TSDFFusion
        f = TSDFFusion(projs, depths, None, center=c["center"], radius=c["radius"])
        fn = lambda: f.compute_unbounded_tsdf(pts, True, c["voxel_size"])  # noqa: E731
        n, w = 5, 2
    else:
        center = torch.from_numpy(c["center"]).cuda()
        fn = lambda: torch_rule_unbounded(pts, projs, depths, None, center, c["radius"], c["voxel_size"])  # noqa: E731
        n, w = 2, 1
    return _entry(ev_times(fn, n, w), pts.shape[0] * len(projs), "sample-views/s")


def mesh_extract(impl, n=512):
    """extract_triangle_mesh + post_process_mesh (mesh_utils.py:27-49,178; mcube_utils.py:71-80) of a bumpy sphere in a
    n^3 lattice.  Units: lattice voxels per second through both steps."""
    import torch
    ax = torch.arange(n, device="cuda", dtype=torch.float32)
    z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
    c = (n - 1) / 2
    f = torch.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2) - 0.35 * n
    f = f + 3.0 * torch.sin(x * 0.21) * torch.sin(y * 0.17) * torch.sin(z * 0.13)
    del x, y, z
    if impl == "ours":
        from gsr_b200.mesh import extract_triangle_mesh, post_process_mesh
        fn = lambda: post_process_mesh(extract_triangle_mesh(f, voxel_size=0.01), cluster_to_keep=50)  # noqa: E731
        out = _entry(ev_times(fn, 5, 2), n ** 3, "voxels/s")
        m = extract_triangle_mesh(f, voxel_size=0.01)
        out["triangles"] = int(m.triangles.shape[0])
        return out
    import time
    from oracle import mcubes_oracle, mesh_clusters_oracle
    k = 128
    lo = (n - k) // 2
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        crop = f[lo:lo + k, lo:lo + k, :k].contiguous().cpu().numpy()           # the unbounded path's `.cpu().numpy()`
        v, faces, _ = mcubes_oracle.extract(crop, voxel_size=0.01)
        mesh_clusters_oracle.post_process_mesh(v, faces, None, cluster_to_keep=50)
        ts.append((time.perf_counter() - t0) * 1e3)
    out = _entry(ts, k ** 3, "voxels/s")
    out.update(kind="port", sample=f"{k}^3 crop of the {n}^3 lattice, numpy/scipy on the host", triangles=int(faces.shape[0]))
    return out


WORKLOADS = {
    "cfgA_surfel": (lambda impl: surfel_cfg_a(impl), "surfel"),
    "gauss_1m": (lambda impl: ewa(impl, False), "gaussian"),
    "cfg4_plane": (lambda impl: ewa(impl, True), "plane"),
    "visible_filter": (lambda impl: visible_filter(impl), "filter"),
    "dist2_knn3": (lambda impl: dist2_knn3(impl), "knn"),
    # the reference arm of the next two is the reference's own torch formulation (no oracle/_ref library involved)
    "ssim_loss": (lambda impl: ssim_loss(impl), None),
    "tsdf_fuse": (lambda impl: tsdf_fuse(impl), None),
    "mesh_extract": (lambda impl: mesh_extract(impl), None),
}


def run_all(impl):
    """{workload: entry} for one arm; reference workloads whose oracle/_ref library is absent are reported as such."""
    import torch
    out = {}
    for name, (fn, variant) in WORKLOADS.items():
        if impl != "ours":
            from oracle import refcuda
            if variant is not None and not refcuda.available(variant):
                out[name] = {"unavailable": f"oracle/_ref/libref_{variant}.so absent"}
                continue
        try:
            out[name] = fn(impl)
        except Exception as e:  # keep the headline line alive; the failure is visible in the record
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import json
    for arm in (sys.argv[1:] or ["ours", "reference"]):
        print(arm, json.dumps(run_all(arm), indent=1))
