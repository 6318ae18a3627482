"""Developer GPU check (not a pytest) for the EWA family + filter + knn: product vs golden vectors
(reference CUDA build) on the seeded cases, then product vs reference CUDA live at a larger size
with timings.

  python tests/gpu_dev_check_ewa.py [--P 200000 --W 800 --H 800]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz  # noqa: E402
import synth  # noqa: E402
from golden.cases import (FILTER_CASES, GAUSS_CASES, KNN_CASES, build_filter_case, build_gauss_case,  # noqa: E402
                          build_knn_case)

GOLD = os.path.join(hz.ROOT, "tests", "golden")


def golden_cases():
    import torch
    for name in GAUSS_CASES:
        sc, kw = build_gauss_case(name)
        gold = np.load(os.path.join(GOLD, f"gauss_{name}.npz"))
        o = hz.run_product_gauss(sc, **kw)
        torch.cuda.synchronize()
        print(f"== {name}: radii mismatch {(o['radii'] != gold['radii']).sum()}")
        print("   color", hz.rel_linf(o["color"], gold["color"]), hz.rel_linf(o["color"], gold["color"], 1e-3))
        if kw["plane"]:
            print("   observe mismatch", (o["observe"] != gold["observe"]).sum(), "of", (gold["observe"] > 0).sum())
            if kw.get("render_geo", True):
                print("   all_map", hz.rel_linf(o["out_all_map"], gold["out_all_map"]), "plane_depth",
                      hz.rel_linf(o["plane_depth"], gold["plane_depth"]), hz.rel_linf(o["plane_depth"], gold["plane_depth"], 1e-3))
        g = {k[5:]: gold[k] for k in gold.files if k.startswith("grad_") and k != "grad_conic"}
        hz.compare_grads_keys(o["grads"], g, list(g.keys()), names=("product", "ref"))
    from scaffold_filter import GaussianRasterizationSettings, GaussianRasterizer
    for name in FILTER_CASES:
        sc, kw = build_filter_case(name)
        tt = hz.to_torch(sc)
        rs = GaussianRasterizationSettings(sc.cam.H, sc.cam.W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"],
                                           kw.get("scale_modifier", 1.0), tt["view"], tt["proj"], 0, tt["campos"], False, False)
        r = GaussianRasterizer(rs).visible_filter(tt["means3D"], tt["scales"], tt["rotations"]).cpu().numpy()
        gold = np.load(os.path.join(GOLD, f"filter_{name}.npz"))["radii"]
        print(name, "radii mismatch", (r != gold).sum(), "visible mismatch", ((r > 0) != (gold > 0)).sum())
    from simple_knn._C import distCUDA2
    for name in KNN_CASES:
        pts = build_knn_case(name)
        d = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
        gold = np.load(os.path.join(GOLD, f"knn_{name}.npz"))["dist2"]
        print(name, "bit mismatches", (d.view(np.uint32) != gold.view(np.uint32)).sum(), "max rel",
              float(np.max(np.abs(d - gold) / np.maximum(gold, 1e-30))))


def timed(fn, n=5):
    import torch
    fn(); fn()
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.time() - t) / n * 1e3


def live(P, W, H):
    import torch
    from oracle import refcuda
    for plane in (False, True):
        sc = synth.make_scene(P, W, H, seed=11, rotate_camera=True, bg=(0.1, 0.2, 0.3), scale_dims=3)
        gc, go = synth.make_upstream_grads(W, H, seed=12, n_others=6, zero_from=6)
        kw = dict(g_color=gc, plane=plane)
        if plane:
            kw.update(all_map=synth.make_all_map(sc), g_all_map=np.ascontiguousarray(go[:5]),
                      g_plane_depth=np.ascontiguousarray(go[5:6]))
        tt = hz.to_torch(sc)
        mine = hz.run_product_gauss(sc, tt=tt, **kw)
        ref = hz.run_refcuda_gauss(sc, tt=tt, **kw)
        print(f"--- live plane={plane} P={P} {W}x{H}: R_ref={ref['num_rendered']} radii mismatch "
              f"{(mine['radii'] != ref['radii']).sum()}")
        print("   color", hz.rel_linf(mine["color"], ref["color"]), hz.rel_linf(mine["color"], ref["color"], 1e-3))
        if plane:
            print("   observe mismatch", (mine["observe"] != ref["observe"]).sum(), "all_map",
                  hz.rel_linf(mine["out_all_map"], ref["out_all_map"], 1e-3), "plane_depth",
                  hz.rel_linf(mine["plane_depth"], ref["plane_depth"], 1e-3))
        g = {k: v for k, v in ref["grads"].items() if k != "conic"}
        hz.compare_grads_keys(mine["grads"], g, list(g.keys()), names=("product", "refcuda"), outlier_frac=2e-3)
        t_mine = timed(lambda: hz.run_product_gauss(sc, tt=tt, **kw), 3)
        t_ref = timed(lambda: hz.run_refcuda_gauss(sc, tt=tt, **kw), 3)
        print(f"   wall (incl. host copies) product {t_mine:.1f} ms   reference {t_ref:.1f} ms")
    # knn + filter at size
    from simple_knn._C import distCUDA2
    for n, clustered in ((P, False), (P, True), (2_000_000, False)):
        pts = torch.from_numpy(synth.make_points(n, seed=5, clustered=clustered)).cuda()
        d = distCUDA2(pts)
        r = refcuda.ref_dist2_knn3(pts)
        bad = (d.view(torch.int32) != r.view(torch.int32)).sum().item()
        t_m = timed(lambda: distCUDA2(pts), 3)
        t_r = timed(lambda: refcuda.ref_dist2_knn3(pts), 3)
        print(f"knn P={n} clustered={clustered}: bit mismatches {bad}; product {t_m:.2f} ms, reference {t_r:.2f} ms")
    from scaffold_filter import GaussianRasterizationSettings, GaussianRasterizer
    sc = synth.make_scene(2_000_000, 1600, 1060, seed=3, scale_dims=3)
    tt = hz.to_torch(sc)
    rs = GaussianRasterizationSettings(sc.cam.H, sc.cam.W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"],
                                       tt["proj"], 0, tt["campos"], False, False)
    rast = GaussianRasterizer(rs)
    r = rast.visible_filter(tt["means3D"], tt["scales"], tt["rotations"])
    rr = refcuda.ref_visible_filter(tt["means3D"], tt["scales"], tt["rotations"], tt["view"], tt["proj"], sc.cam.W,
                                    sc.cam.H, sc.cam.tanfovx, sc.cam.tanfovy)
    t_m = timed(lambda: rast.visible_filter(tt["means3D"], tt["scales"], tt["rotations"]), 5)
    t_r = timed(lambda: refcuda.ref_visible_filter(tt["means3D"], tt["scales"], tt["rotations"], tt["view"], tt["proj"],
                                                    sc.cam.W, sc.cam.H, sc.cam.tanfovx, sc.cam.tanfovy), 5)
    print(f"filter 2M anchors: radii mismatch {(r != rr).sum().item()} visible mismatch {((r > 0) != (rr > 0)).sum().item()}; "
          f"product {t_m:.2f} ms, reference {t_r:.2f} ms")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=200000)
    ap.add_argument("--W", type=int, default=800)
    ap.add_argument("--H", type=int, default=800)
    ap.add_argument("--no-live", action="store_true")
    a = ap.parse_args()
    golden_cases()
    if not a.no_live:
        live(a.P, a.W, a.H)
