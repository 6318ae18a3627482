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
from golden.cases import CASES, build_case  # noqa: E402


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


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(hz.ROOT, "gpurun_out", "golden"))
