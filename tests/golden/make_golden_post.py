"""Generates tests/golden/post_*.npz: the 2DGS post-processing block evaluated with the REFERENCE's own
depths_to_points / depth_to_normal (/root/reference/gssr/utils/point_utils.py:9-37, cut out of the source at generation
time: the module imports cv2 / matplotlib and creates CUDA tensors, so it is executed on CPU torch with torch.arange /
torch.tensor stripped of their device argument and Tensor.cuda() -> identity) composed with the lines of
TwoDGSScene.render that call it (gssr/scene/twodgs_scene.py:88-117, restated here because they sit inside a method that
also calls the rasterizer).  Values and autograd gradients w.r.t. allmap are stored.

    python tests/golden/make_golden_post.py      # needs /root/reference
"""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from post_synth import POST_CASES, build_post_case  # noqa: E402

REF = "/root/reference/gssr/utils/point_utils.py"


def reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    shim = types.SimpleNamespace(**{k: getattr(torch, k) for k in dir(torch) if not k.startswith("__")})
    strip = lambda f: (lambda *a, device=None, **k: f(*a, **k))  # noqa: E731
    shim.arange, shim.tensor = strip(torch.arange), strip(torch.tensor)
    ns = {"torch": shim}
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("depths_to_points", "depth_to_normal"):
            exec(compile(textwrap.dedent(ast.get_source_segment(src, fn)), REF, "exec"), ns)
    return ns


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    ns = reference_functions()
    for name in POST_CASES:
        c = build_post_case(name)
        view = types.SimpleNamespace(world_view_transform=torch.from_numpy(c["wvt"]), full_proj_transform=torch.from_numpy(c["full_proj"]),
                                     image_width=c["W"], image_height=c["H"])
        allmap = torch.from_numpy(c["allmap"]).requires_grad_(True)
        # twodgs_scene.py:88-117
        render_alpha = allmap[1:2]
        render_normal = allmap[2:5]
        render_normal = (render_normal.permute(1, 2, 0) @ (view.world_view_transform[:3, :3].T)).permute(2, 0, 1)
        render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
        render_depth_expected = (allmap[0:1] / render_alpha)
        render_depth_expected = torch.nan_to_num(render_depth_expected, 0, 0)
        surf_depth = render_depth_expected * (1 - c["depth_ratio"]) + (c["depth_ratio"]) * render_depth_median
        surf_normal = ns["depth_to_normal"](view, surf_depth)
        surf_normal = surf_normal.permute(2, 0, 1)
        surf_normal = surf_normal * (render_alpha).detach()
        g = {k: torch.from_numpy(v) for k, v in c["g"].items()}
        torch.autograd.backward([render_normal, surf_depth, surf_normal], [g["normal"], g["depth"], g["surf_normal"]])
        np.savez_compressed(os.path.join(HERE, f"post_{name}.npz"), normal=render_normal.detach().numpy(),
                            depth=surf_depth.detach().numpy(), surf_normal=surf_normal.detach().numpy(), grad=allmap.grad.numpy())
        print(name, "nan grads:", int(torch.isnan(allmap.grad).sum()))


if __name__ == "__main__":
    main()
