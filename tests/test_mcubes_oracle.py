"""CPU: the marching-cubes restatement (oracle/mcubes_oracle.py) has the properties that pin it -- Open3D / skimage, which
the reference calls for this step (mesh_utils.py:178, mcube_utils.py:71-80), are absent, so parity against them is
UNPINNED -- and the product's generated case table (gs-sr_b200/csrc/mc_table.cuh) is the same table."""
import os
import re
import subprocess
import sys

import numpy as np

from oracle import mcubes_oracle as mc
import mesh_synth as ms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gs-sr_b200", "csrc")


def _closed(faces):
    return all(u == [1, 1] for u in mc.edge_use_counts(faces).values())


def _signed_volume(v, f):
    p = v[f].astype(np.float64)
    return float(np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() / 6.0)


def test_case_table_shape_and_symmetries():
    assert mc.MAX_TRIS == 5 and mc.NTRI[0] == 0 and mc.NTRI[255] == 0
    for case in range(256):
        tris = mc.case_triangles(case)
        used = sorted({e for t in tris for e in t})
        crossing = sorted(e for e in range(12) if ((case >> mc.edge_corners(e)[0]) ^ (case >> mc.edge_corners(e)[1])) & 1)
        assert used == crossing                          # every crossed cell edge carries a vertex, no other does
        assert all(len(set(t)) == 3 for t in tris)


def test_sphere_is_closed_outward_and_on_the_level_set():
    f = ms.sphere(40, 12.3)
    v, faces, col = mc.extract(f, origin=(-1.0, 2.0, 0.5), voxel_size=0.25)
    assert col is None and _closed(faces)
    edges = mc.edge_use_counts(faces)
    assert len(v) - len(edges) + len(faces) == 2                          # Euler characteristic of a sphere
    vol = _signed_volume(v, faces)
    assert 0.98 < vol / (4 / 3 * np.pi * (12.3 * 0.25) ** 3) < 1.0      # > 0: normals point to f > 0 (outside)
    centre = np.array([-1.0, 2.0, 0.5]) + 19.5 * 0.25
    assert np.abs(np.linalg.norm(v - centre, axis=1) - 12.3 * 0.25).max() < 0.01 * 0.25 * 12.3


def test_torus_has_genus_one():
    v, faces, _ = mc.extract(ms.torus())
    assert _closed(faces)
    assert len(v) - len(mc.edge_use_counts(faces)) + len(faces) == 0


def test_noise_field_is_watertight_away_from_the_lattice_boundary():
    for shape, seed in (((12, 13, 14), 0), ((20, 9, 11), 1), ((7, 31, 8), 2)):
        f = ms.noise(shape, seed)
        v, faces, _ = mc.extract(f, origin=(1.0, 2.0, 3.0), voxel_size=0.5)
        onb = ms.is_lattice_boundary_vertex(v, (1.0, 2.0, 3.0), 0.5, shape[::-1])
        for (a, b), u in mc.edge_use_counts(faces).items():
            assert u == [1, 1] or (onb[a] and onb[b] and sum(u) == 1)
        assert len(np.unique(faces)) == len(v)           # no unreferenced vertex
        assert (faces[:, 0] != faces[:, 1]).all() and (faces[:, 1] != faces[:, 2]).all()


def test_unobserved_corners_switch_cells_off():
    f, w, rgb = ms.observed_blob((10, 12, 9), 5)
    v, faces, col = mc.extract(f, w, 1.0, rgb=rgb)
    case = mc.cell_cases(f, w, 1.0)
    ok = w > 1
    cells = case != 0
    nz, ny, nx = f.shape
    for k in range(8):
        dx, dy, dz = k & 1, (k >> 1) & 1, (k >> 2) & 1
        sl = (slice(dz, nz - 1 + dz), slice(dy, ny - 1 + dy), slice(dx, nx - 1 + dx))
        assert ok[sl][cells[:nz - 1, :ny - 1, :nx - 1]].all()
    assert len(faces) == int(mc.NTRI[case].sum()) and 0 < len(faces) < int(mc.NTRI[mc.cell_cases(f)].sum())
    assert len(np.unique(faces)) == len(v) and col.shape == v.shape
    assert col.min() >= 0.0 and col.max() <= 1.0        # an interpolation of the corner colours


def test_level_and_degenerate_inputs():
    f = ms.sphere(24, 7.0)
    a = mc.extract(f, level=1.5)
    b = mc.extract((f - np.float32(1.5)).astype(np.float32))
    assert np.array_equal(a[1], b[1]) and np.abs(a[0] - b[0]).max() < 1e-5
    for const in (1.0, -1.0):
        v, faces, _ = mc.extract(np.full((5, 6, 7), const, dtype=np.float32))
        assert v.shape == (0, 3) and faces.shape == (0, 3)
    v, faces, _ = mc.extract(np.array([[[-1.0]]], dtype=np.float32))      # a single voxel: no cell
    assert len(v) == 0 and len(faces) == 0
    f = np.ones((3, 3, 3), dtype=np.float32)
    f[1, 1, 1] = 0.0                                                       # exactly on the level: not inside
    assert len(mc.extract(f)[1]) == 0
    f[1, 1, 1] = -1.0
    v, faces, _ = mc.extract(f)                                            # one inside corner: an octahedron
    assert len(v) == 6 and len(faces) == 8 and _closed(faces)


def test_generated_product_table_is_the_oracle_table():
    text = open(os.path.join(CSRC, "mc_table.cuh")).read()
    regen = subprocess.run([sys.executable, os.path.join(CSRC, "gen_mc_table.py")], capture_output=True, text=True, check=True)
    assert regen.stdout == text                                            # the committed header is the generator's output
    ntri = [int(x) for x in re.search(r"MC_NTRI\[256\] = \{(.*?)\};", text, re.S).group(1).replace("\n", " ").split(",") if x.strip()]
    words = [int(x.strip().rstrip("ul"), 16) for x in re.search(r"MC_TRIS\[256\] = \{(.*?)\};", text, re.S).group(1).split(",") if x.strip()]
    assert len(ntri) == 256 and len(words) == 256
    for case in range(256):
        assert ntri[case] == mc.NTRI[case]
        for t in range(ntri[case]):
            tri = [(words[case] >> (12 * t + 4 * k)) & 15 for k in range(3)]
            assert tri == [int(e) for e in mc.TABLE[case, 3 * t:3 * t + 3]]
        assert words[case] >> (12 * ntri[case]) == 0


def test_oracles_reproduce_the_frozen_conventions():
    """tests/golden/mesh_*.npz (make_golden_mesh.py): frozen outputs of the two restatements -- the reference has nothing
    to pin them to (Open3D / skimage absent), so the repo's own conventions are pinned against drift."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_mesh import CASES, build
    from oracle import mesh_clusters_oracle as mo
    for name in CASES:
        g = np.load(os.path.join(ROOT, "tests", "golden", f"mesh_{name}.npz"))
        assert np.array_equal(g["table"], mc.TABLE) and np.array_equal(g["ntri"], mc.NTRI)
        c, f, w, rgb = build(name)
        v, faces, col = mc.extract(f, w, 1.0 if w is not None else None, c["level"], c["origin"], c["voxel"], rgb)
        assert np.array_equal(v, g["verts"]) and np.array_equal(faces, g["faces"])
        _, troot, ntris, _ = mo.clusters(v, faces)
        assert np.array_equal(troot, g["tri_root"]) and np.array_equal(ntris, g["root_ntris"])
        pv, pf, pc = mo.post_process_mesh(v, faces, col, cluster_to_keep=2, min_triangles=4)
        assert np.array_equal(pv, g["post_verts"]) and np.array_equal(pf, g["post_faces"])
        if col is not None:
            assert np.array_equal(col, g["colors"]) and np.array_equal(pc, g["post_colors"])
