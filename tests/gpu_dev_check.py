"""Developer GPU check (not a pytest): product vs CPU oracle vs reference CUDA on seeded
scenes, plus fwd/bwd timing of product and reference at a chosen size.

  python tests/gpu_dev_check.py [--P 100000 --W 800 --H 800] [--time-P 2000000 --time-W 1600 --time-H 1060]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz  # noqa: E402
import synth  # noqa: E402


def parity(P, W, H, sh, seed, oracle=True):
    import torch
    sc = synth.make_scene(P, W, H, seed=seed, sh=sh, rotate_camera=True, bg=(0.1, 0.2, 0.3))
    gc, go = synth.make_upstream_grads(W, H, seed=seed + 1)
    go[8:11] = (np.random.default_rng(5).normal(size=(3, H, W)) / (W * H)).astype(np.float32) * 0  # keep 0
    tt = hz.to_torch(sc)
    mine = hz.run_product_surfel(sc, gc, go, tt=tt)
    torch.cuda.synchronize()
    res = {}
    if hz_ref_available():
        ref = hz.run_refcuda_surfel(sc, gc, go, tt=tt)
        print(f"--- P={P} {W}x{H} sh={sh}: R={ref['num_rendered']} visible={(ref['radii'] > 0).sum()}")
        res["fwd_vs_ref"] = hz.compare_forward(mine, ref, ("product", "refcuda"))
        res["grad_vs_ref"] = hz.compare_grads(mine["grads"], ref["grads"], ("product", "refcuda"))
    if oracle:
        t = time.time()
        orc = hz.run_oracle_surfel(sc, gc, go)
        print(f"oracle time {time.time() - t:.1f}s  R={orc['num_rendered']} k_eval={orc['k_eval']}")
        res["fwd_vs_orc"] = hz.compare_forward(mine, orc, ("product", "oracle"))
        res["grad_vs_orc"] = hz.compare_grads(mine["grads"], orc["grads"], ("product", "oracle"))
        if hz_ref_available():
            res["ref_fwd_vs_orc"] = hz.compare_forward(ref, orc, ("refcuda", "oracle"))
            res["ref_grad_vs_orc"] = hz.compare_grads(ref["grads"], orc["grads"], ("refcuda", "oracle"))
            if os.environ.get("GSR_DUMP"):
                np.savez_compressed(f"gpurun_out/dump_P{P}_sh{int(sh)}.npz",
                                    **{f"ref_{k}": v for k, v in ref["grads"].items()},
                                    **{f"orc_{k}": v for k, v in orc["grads"].items()},
                                    **{f"mine_{k}": v for k, v in mine["grads"].items() if v is not None},
                                    ref_color=ref["color"], ref_others=ref["others"], mine_color=mine["color"],
                                    mine_others=mine["others"], ref_radii=ref["radii"], mine_radii=mine["radii"])
    return res


def hz_ref_available():
    from oracle import refcuda
    return refcuda.available("surfel")


def timing(P, W, H, iters=10):
    import torch
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    sc = synth.make_scene(P, W, H, seed=0)
    gc, go = synth.make_upstream_grads(W, H)
    tt = hz.to_torch(sc)
    gct, got = torch.from_numpy(gc).cuda(), torch.from_numpy(go).cuda()
    rs = GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"], tt["proj"],
                                       0, tt["campos"], False, False)
    rast = GaussianRasterizer(rs)
    leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "colors")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)

    def step():
        color, radii, others = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                    colors_precomp=leaves["colors"], scales=leaves["scales"],
                                    rotations=leaves["rotations"])
        torch.autograd.backward([color, others], [gct, got])
        return radii

    for _ in range(3):
        radii = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"[product] P={P} {W}x{H}: fwd+bwd median {ms:.3f} ms  -> {P / ms * 1e3 / 1e6:.1f} M Gaussians/s; visible={(radii > 0).sum().item()}")
    out = {"product_ms": ms}
    if hz_ref_available():
        from oracle.refcuda import RefSurfel
        r = RefSurfel()
        args = (tt["bg"], tt["view"], tt["proj"], tt["campos"], W, H, sc.cam.tanfovx, sc.cam.tanfovy, tt["means3D"],
                tt["opacities"], tt["scales"], tt["rotations"])

        def rstep():
            r.forward(*args, colors=tt["colors"])
            r.backward(gct, got)
        for _ in range(2):
            rstep()
        torch.cuda.synchronize()
        ts, tf = [], []
        for _ in range(max(3, iters // 2)):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _, _, _, R = r.forward(*args, colors=tt["colors"]); torch.cuda.synchronize(); t1 = time.perf_counter()
            r.backward(gct, got); torch.cuda.synchronize(); t2 = time.perf_counter()
            ts.append((t2 - t0) * 1e3); tf.append((t1 - t0) * 1e3)
        rms = float(np.median(ts))
        print(f"[refcuda] P={P} {W}x{H}: fwd+bwd median {rms:.3f} ms (fwd {np.median(tf):.3f}) -> {P / rms * 1e3 / 1e6:.1f} M Gaussians/s; R={R}")
        out.update(ref_ms=rms, ref_fwd_ms=float(np.median(tf)), R=int(R), speedup=rms / ms)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=20000)
    ap.add_argument("--W", type=int, default=320)
    ap.add_argument("--H", type=int, default=240)
    ap.add_argument("--time-P", type=int, default=2000000)
    ap.add_argument("--time-W", type=int, default=1600)
    ap.add_argument("--time-H", type=int, default=1060)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--no-timing", action="store_true")
    ap.add_argument("--out", default="gpurun_out/dev_check.json")
    a = ap.parse_args()
    res = {}
    res["parity_colors"] = parity(a.P, a.W, a.H, sh=False, seed=11, oracle=not a.no_oracle)
    res["parity_sh"] = parity(max(a.P // 4, 100), a.W, a.H, sh=True, seed=12, oracle=not a.no_oracle)
    if not a.no_timing:
        res["timing_small"] = timing(100000, 800, 800)
        res["timing"] = timing(a.time_P, a.time_W, a.time_H)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1, default=str)
