"""Generates tests/golden/depth_normal_*.npz with the REFERENCE's own normal_from_depth_image
(/root/reference/gssr/utils/graphics_utils.py:79-146: ndc_2_cam, depth2point_cam, depth2point_world, depth_pcd2normal,
normal_from_depth_image, cut out of the source at generation time and executed verbatim with CPU torch -- the module
itself imports nothing exotic, but is loaded by path so that the gssr package's CUDA-only __init__ imports stay out).
Values and autograd gradients w.r.t. the depth map are stored, without and with the alpha weight PGSR applies
(gssr/scene/pgsr_scene.py:320).

    python tests/golden/make_golden_depth_normal.py      # needs /root/reference
"""
import ast
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from depth_normal_synth import DN_CASES, build_dn_case  # noqa: E402

REF = "/root/reference/gssr/utils/graphics_utils.py"
WANT = ("ndc_2_cam", "depth2point_cam", "depth2point_world", "depth_pcd2normal", "normal_from_depth_image")


def reference_functions():
    src = open(REF).read()
    ns = {"torch": torch, "np": np}
    for fn in ast.parse(src).body:
        if isinstance(fn, ast.FunctionDef) and fn.name in WANT:
            exec(compile(textwrap.dedent(ast.get_source_segment(src, fn)), REF, "exec"), ns)
    return ns


def main():
    ns = reference_functions()
    for name in DN_CASES:
        c = build_dn_case(name)
        K, ext = torch.from_numpy(c["K"]), torch.eye(4)
        out = {}
        for tag, w in (("", None), ("_w", torch.from_numpy(c["weight"]))):
            d = torch.from_numpy(c["depth"]).requires_grad_(True)
            n = ns["normal_from_depth_image"](d, K, ext).permute(2, 0, 1)          # render_normal, pgsr_scene.py:233-237
            if w is not None:
                n = n * w
            n.backward(torch.from_numpy(c["g"]))
            out["normal" + tag], out["grad" + tag] = n.detach().numpy(), d.grad.numpy()
        np.savez_compressed(os.path.join(HERE, f"depth_normal_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()}, "nan:", int(np.isnan(out["grad"]).sum()))


if __name__ == "__main__":
    main()
