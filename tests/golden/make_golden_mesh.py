#!/usr/bin/env python3
"""Writes tests/golden/mesh_*.npz: the marching-cubes and cluster-filter restatements (oracle/mcubes_oracle.py,
oracle/mesh_clusters_oracle.py) frozen on seeded lattices.

NOT reference outputs: the reference delegates both steps to Open3D 0.18 / skimage (absent, parity unpinned, see the
oracles' headers).  These vectors pin the CONVENTIONS this repo fixed instead -- the derived 256-case table, vertex
order (owner voxel, axis), face order (cell, table order), winding, the observed-cell rule, colour interpolation, cluster
labels and the filter's order of operations -- so that neither the oracle nor the kernels can drift unnoticed.

    python tests/golden/make_golden_mesh.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import mesh_synth as ms  # noqa: E402
from oracle import mcubes_oracle as mc  # noqa: E402
from oracle import mesh_clusters_oracle as mo  # noqa: E402

CASES = {
    "noise": dict(shape=(9, 10, 11), seed=41, masked=False, level=0.0, origin=(0.5, -1.0, 2.0), voxel=0.25),
    "observed": dict(shape=(12, 9, 10), seed=42, masked=True, level=0.2, origin=(0.0, 0.0, 0.0), voxel=0.05),
}


def build(name):
    c = CASES[name]
    if c["masked"]:
        f, w, rgb = ms.observed_blob(c["shape"], c["seed"])
    else:
        f, w, rgb = ms.noise(c["shape"], c["seed"]), None, None
    return c, f, w, rgb


def main():
    for name in CASES:
        c, f, w, rgb = build(name)
        v, faces, col = mc.extract(f, w, 1.0 if w is not None else None, c["level"], c["origin"], c["voxel"], rgb)
        vroot, troot, ntris, _ = mo.clusters(v, faces)
        pv, pf, pc = mo.post_process_mesh(v, faces, col, cluster_to_keep=2, min_triangles=4)
        out = dict(table=mc.TABLE, ntri=mc.NTRI, verts=v, faces=faces, tri_root=troot, root_ntris=ntris.astype(np.int32),
                   post_verts=pv, post_faces=pf)
        if col is not None:
            out.update(colors=col, post_colors=pc)
        np.savez_compressed(os.path.join(HERE, f"mesh_{name}.npz"), **out)
        print(name, v.shape, faces.shape, "clusters", int((ntris > 0).sum()), "post", pf.shape)


if __name__ == "__main__":
    main()
