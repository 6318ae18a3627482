"""Generates the golden vectors that pin the CPU oracle: outputs of the UNMODIFIED
reference CUDA kernels (oracle/_ref/libref_surfel.so, built by oracle/build_ref.sh
from /root/reference/submodules/diff-surfel-rasterization) on seeded synthetic
scenes, run on a B200.

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Inputs are NOT stored: every case is regenerated from (generator, seed) by
tests/golden/cases.py, so fixtures stay small.  Outputs stored per case:
forward colour / allmap / radii / num_rendered and all backward gradients.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as hz  # noqa: E402
from golden.cases import (CASES, FILTER_CASES, GAUSS_CASES, KNN_CASES, build_case, build_filter_case,  # noqa: E402
                          build_gauss_case, build_knn_case)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    import torch
    for name in CASES:
        sc, gc, go, kw = build_case(name)
        ref = hz.run_refcuda_surfel(sc, gc, go, **kw)
        torch.cuda.synchronize()
        arrs = dict(color=ref["color"], others=ref["others"], radii=ref["radii"],
                    num_rendered=np.int64(ref["num_rendered"]))
        for k, v in ref["grads"].items():
            arrs["grad_" + k] = v
        np.savez_compressed(os.path.join(out_dir, f"surfel_{name}.npz"), **arrs)
        print(name, "R=", ref["num_rendered"], "visible=", int((ref["radii"] > 0).sum()))


def main_gauss(out_dir):
    """3DGS / plane / filter / knn vectors from oracle/_ref/libref_{gaussian,plane,filter,knn}.so."""
    os.makedirs(out_dir, exist_ok=True)
    import torch
    from oracle import refcuda
    for name in GAUSS_CASES:
        sc, kw = build_gauss_case(name)
        ref = hz.run_refcuda_gauss(sc, **kw)
        arrs = {k: v for k, v in ref.items() if isinstance(v, np.ndarray)}
        arrs["num_rendered"] = np.int64(ref["num_rendered"])
        for k, v in ref["grads"].items():
            arrs["grad_" + k] = v
        np.savez_compressed(os.path.join(out_dir, f"gauss_{name}.npz"), **arrs)
        print(name, "R=", ref["num_rendered"], "visible=", int((ref["radii"] > 0).sum()))
    for name in FILTER_CASES:
        sc, kw = build_filter_case(name)
        tt = hz.to_torch(sc)
        radii = refcuda.ref_visible_filter(tt["means3D"], tt["scales"], tt["rotations"], tt["view"], tt["proj"],
                                           sc.cam.W, sc.cam.H, sc.cam.tanfovx, sc.cam.tanfovy, **kw)
        np.savez_compressed(os.path.join(out_dir, f"filter_{name}.npz"), radii=radii.cpu().numpy())
        print(name, "visible=", int((radii > 0).sum()))
    for name in KNN_CASES:
        pts = build_knn_case(name)
        d = refcuda.ref_dist2_knn3(torch.from_numpy(pts).cuda())
        np.savez_compressed(os.path.join(out_dir, f"knn_{name}.npz"), dist2=d.cpu().numpy())
        print(name, "mean=", float(d.mean()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gauss":
        main_gauss(sys.argv[2] if len(sys.argv) > 2 else os.path.join(hz.ROOT, "gpurun_out", "golden"))
        sys.exit(0)
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(hz.ROOT, "gpurun_out", "golden"))
