"""Generates tests/golden/ssim_*.npz with the REFERENCE's own SSIM code: the methods _ssim / _gaussian / ssim of
VanillaScene (/root/reference/gssr/scene/vanilla_scene.py:32-61) are cut out of the reference source at generation
time (the module itself cannot be imported here: it needs the compiled rasterizer extensions) and executed verbatim on
CPU torch; value and autograd gradient w.r.t. img1 are stored.  Only numeric outputs are committed.

    python tests/golden/make_golden_ssim.py      # needs /root/reference
"""
import ast
import math
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F
from torch.autograd import Variable

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from ssim_synth import SSIM_CASES, build_ssim_case  # noqa: E402

REF = "/root/reference/gssr/scene/vanilla_scene.py"


def reference_methods():
    src = open(REF).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "VanillaScene")
    ns = {"torch": torch, "F": F, "math": math, "Variable": Variable}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("_ssim", "_gaussian", "ssim"):
            exec(compile(textwrap.dedent(ast.get_source_segment(src, fn)), REF, "exec"), ns)
    obj = types.SimpleNamespace()
    for name in ("_ssim", "_gaussian", "ssim"):
        setattr(obj, name, types.MethodType(ns[name], obj))
    return obj


def main():
    ref = reference_methods()
    for name in SSIM_CASES:
        a, b = build_ssim_case(name)
        x = torch.from_numpy(a).requires_grad_(True)
        v = ref.ssim(x, torch.from_numpy(b))
        v.backward()
        np.savez_compressed(os.path.join(HERE, f"ssim_{name}.npz"), value=np.float32(v.item()), grad=x.grad.numpy())
        print(name, a.shape, "ssim =", v.item())


if __name__ == "__main__":
    main()
