"""Developer GPU profiling helper (not a pytest).

  python tests/gpu_profile.py host   [--P ... --W ... --H ...]   host-side time split of one step
  python tests/gpu_profile.py steps  [--iters N] [--ref]         run N fwd+bwd steps (wrap with ncu)
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


def setup(P, W, H, sh=False):
    import torch
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    sc = synth.make_scene(P, W, H, seed=0, sh=sh)
    gc, go = synth.make_upstream_grads(W, H)
    tt = hz.to_torch(sc)
    gct, got = torch.from_numpy(gc).cuda(), torch.from_numpy(go).cuda()
    rs = GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"], tt["proj"],
                                       sc.sh_degree, tt["campos"], False, False)
    rast = GaussianRasterizer(rs)
    keys = ("means3D", "scales", "rotations", "opacities") + (("shs",) if sh else ("colors",))
    leaves = {k: tt[k].clone().requires_grad_(True) for k in keys}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    return sc, tt, gct, got, rast, leaves, m2d


def product_step(rast, leaves, m2d, gct, got, sync_between=False):
    import torch
    t0 = time.perf_counter()
    color, radii, others = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                colors_precomp=leaves.get("colors"), shs=leaves.get("shs"),
                                scales=leaves["scales"], rotations=leaves["rotations"])
    t1 = time.perf_counter()
    if sync_between:
        torch.cuda.synchronize()
    t2 = time.perf_counter()
    torch.autograd.backward([color, others], [gct, got])
    t3 = time.perf_counter()
    if sync_between:
        torch.cuda.synchronize()
    t4 = time.perf_counter()
    return (t1 - t0, t2 - t1, t3 - t2, t4 - t3)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["host", "steps"])
    ap.add_argument("--P", type=int, default=2000000)
    ap.add_argument("--W", type=int, default=1600)
    ap.add_argument("--H", type=int, default=1060)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--sh", action="store_true")
    a = ap.parse_args()
    import torch
    sc, tt, gct, got, rast, leaves, m2d = setup(a.P, a.W, a.H, a.sh)
    if a.mode == "host":
        for _ in range(3):
            product_step(rast, leaves, m2d, gct, got)
        torch.cuda.synchronize()
        for i in range(5):
            for v in leaves.values():
                v.grad = None
            m2d.grad = None
            r = product_step(rast, leaves, m2d, gct, got, sync_between=True)
            print("fwd call %.3f ms | fwd gpu tail %.3f ms | bwd call %.3f ms | bwd gpu tail %.3f ms" % tuple(x * 1e3 for x in r))
        # async pipeline
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(10):
            product_step(rast, leaves, m2d, gct, got)
        torch.cuda.synchronize()
        print("10 async steps: %.3f ms/step" % ((time.perf_counter() - t0) * 100))
    else:
        if a.ref:
            from oracle.refcuda import RefSurfel
            r = RefSurfel()
            args = (tt["bg"], tt["view"], tt["proj"], tt["campos"], a.W, a.H, sc.cam.tanfovx, sc.cam.tanfovy,
                    tt["means3D"], tt["opacities"], tt["scales"], tt["rotations"])
            for _ in range(a.iters):
                r.forward(*args, colors=tt["colors"], shs=tt["shs"], sh_degree=sc.sh_degree)
                r.backward(gct, got)
            torch.cuda.synchronize()
        else:
            for _ in range(a.iters):
                product_step(rast, leaves, m2d, gct, got)
            torch.cuda.synchronize()
