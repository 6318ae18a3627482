"""Generates tests/golden/tsdf_*.npz: outputs of the REFERENCE's own TSDF fusion code on the seeded cases
of tests/tsdf_synth.py.

The fusion rule lives in functions nested inside GaussianExtractor.extract_mesh_unbounded
(/root/reference/gssr/utils/mesh_utils.py:187-246), so they cannot be imported; this script cuts their
source text out of the reference file AT GENERATION TIME, executes it verbatim with CPU torch (the only
patch: Tensor.cuda() -> identity, this container has no GPU) and stores the results.  Nothing from the
reference is copied into the repository; only the numeric outputs are committed.

    python tests/golden/make_golden_tsdf.py        # needs /root/reference; writes tests/golden/tsdf_*.npz
"""
from __future__ import annotations

import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tsdf_synth import TSDF_CASES, build_tsdf_case  # noqa: E402

REF = "/root/reference/gssr/utils/mesh_utils.py"


def reference_functions():
    src = open(REF).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.strip().startswith("def contract(x):"))
    end = next(i for i, l in enumerate(src) if l.strip().startswith("normalize = lambda x:"))
    code = textwrap.dedent("\n".join(src[start:end]))
    ns = {"torch": torch, "tqdm": lambda it, **kw: it, "np": np}
    exec(compile(code, REF, "exec"), ns)
    return ns


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self          # no GPU here: the reference's .cuda() calls become no-ops
    _zeros = torch.zeros
    ns = reference_functions()
    for name in TSDF_CASES:
        c = build_tsdf_case(name)
        cams = [types.SimpleNamespace(full_proj_transform=torch.from_numpy(m)) for m in c["projs"]]
        ns["self"] = types.SimpleNamespace(viewpoint_stack=cams, depthmaps=[torch.from_numpy(d) for d in c["depthmaps"]],
                                           rgbmaps=[torch.from_numpy(r) for r in c["rgbmaps"]])
        center, radius = torch.from_numpy(c["center"]), c["radius"]
        unnormalize = lambda x: (x * radius) + center      # mesh_utils.py:249
        inv = (lambda x: unnormalize(ns["uncontract"](x))) if c["contracted"] else None   # :250
        with torch.no_grad():
            tsdf, rgb = ns["compute_unbounded_tsdf"](torch.from_numpy(c["samples"]), inv, c["voxel_size"], return_rgb=True)
        out = os.path.join(HERE, f"tsdf_{name}.npz")
        np.savez_compressed(out, tsdf=tsdf.numpy().astype(np.float32), rgb=rgb.numpy().astype(np.float32))
        print(name, "n=", len(tsdf), "fused fraction=", float((tsdf != 1).float().mean()), "->", out)


if __name__ == "__main__":
    main()
