"""GPU: gsr_mc_count / gsr_mc_emit (through gsr_b200.mesh -> C ABI) reproduce the numpy restatement
(oracle/mcubes_oracle.py) bit for bit -- vertices, colours and faces, in the same order -- and keep the size-independent
properties of the extractor at a full 512^3 lattice (closed, outward, Euler characteristic).  The reference delegates this
step to Open3D / skimage (mesh_utils.py:178, mcube_utils.py:71-80; absent, parity unpinned)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import mesh_synth as ms  # noqa: E402


def _run(f, w=None, min_w=None, level=0.0, origin=(0.0, 0.0, 0.0), voxel=1.0, rgb=None):
    from gsr_b200.mesh import extract_triangle_mesh
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m = extract_triangle_mesh(t(f), t(w), min_w, level, origin, voxel, t(rgb))
    torch.cuda.synchronize()
    return m.numpy()


def _same(got, want):
    for g, w in zip(got, want):
        if w is None:
            assert g is None
        else:
            assert g.dtype == w.dtype and g.shape == w.shape and np.array_equal(g, w)


@pytest.mark.parametrize("shape,seed", [((12, 13, 14), 0), ((33, 32, 31), 1), ((5, 70, 300), 2), ((130, 9, 11), 3), ((2, 2, 2), 4),
                                        ((3, 5, 7), 5), ((9, 33, 2), 6), ((64, 64, 64), 7)])
def test_noise_fields_match_the_oracle_bit_for_bit(shape, seed):
    from oracle import mcubes_oracle as mc
    f = ms.noise(shape, seed)
    _same(_run(f, origin=(1.0, -2.0, 3.5), voxel=0.37), mc.extract(f, origin=(1.0, -2.0, 3.5), voxel_size=0.37))


@pytest.mark.parametrize("level", [0.0, 0.3])
def test_observed_mask_colours_and_level_match_the_oracle(level):
    from oracle import mcubes_oracle as mc
    f, w, rgb = ms.observed_blob((40, 37, 45), 7)
    _same(_run(f, w, 1.0, level, (0.5, 0.25, -4.0), 0.02, rgb), mc.extract(f, w, 1.0, level, (0.5, 0.25, -4.0), 0.02, rgb))
    _same(_run(f, w, 3.0, level), mc.extract(f, w, 3.0, level))


def test_unaligned_lattices_take_the_scalar_loads():
    """tsdf / weight views that start 4 bytes into an allocation (no 16-byte alignment), with nx a multiple of 4."""
    from gsr_b200.mesh import extract_triangle_mesh
    from oracle import mcubes_oracle as mc
    shape = (9, 10, 16)
    f, w, _ = ms.observed_blob(shape, 21)
    n = f.size
    fb, wb = torch.zeros(n + 1, device="cuda"), torch.zeros(n + 1, device="cuda")
    fb[1:].copy_(torch.from_numpy(f).reshape(-1))
    wb[1:].copy_(torch.from_numpy(w).reshape(-1))
    ft, wt = fb[1:].view(shape), wb[1:].view(shape)
    assert ft.data_ptr() % 16 == 4 and ft.is_contiguous()
    m = extract_triangle_mesh(ft, wt, 1.0)
    torch.cuda.synchronize()
    _same(m.numpy(), mc.extract(f, w, 1.0))


def test_smooth_surfaces_match_the_oracle():
    from oracle import mcubes_oracle as mc
    for f in (ms.sphere(40, 12.3), ms.torus()):
        _same(_run(f, voxel=0.25), mc.extract(f, voxel_size=0.25))


def test_empty_results_and_argument_errors():
    from gsr_b200.mesh import extract_triangle_mesh
    for const in (1.0, -1.0):
        v, faces, col = _run(np.full((5, 6, 7), const, dtype=np.float32))
        assert v.shape == (0, 3) and faces.shape == (0, 3) and col is None
    v, faces, _ = _run(np.array([[[-1.0]]], dtype=np.float32))
    assert len(v) == 0 and len(faces) == 0
    f = torch.zeros((4, 4, 4), device="cuda")
    with pytest.raises(ValueError):
        extract_triangle_mesh(f[0])
    with pytest.raises(ValueError):
        extract_triangle_mesh(f, weight=f)
    with pytest.raises(ValueError):
        extract_triangle_mesh(f, rgb=torch.zeros((4, 4, 4), device="cuda"))
    with pytest.raises(RuntimeError):
        extract_triangle_mesh(torch.zeros((4, 4, 4)))


def test_full_size_lattice_keeps_the_closed_surface_properties():
    """512^3 (the chunk the reference's unbounded path meshes at a time, mcube_utils.py:34): two nested spheres and a
    torus; checked on the device through size-independent properties."""
    from gsr_b200.mesh import extract_triangle_mesh
    n = 512
    ax = torch.arange(n, device="cuda", dtype=torch.float32)
    z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
    c = (n - 1) / 2
    d = torch.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2)
    shell = torch.abs(d - 180.3) - 12.7                                   # two nested spheres (a thick shell)
    q = torch.sqrt((x - c) ** 2 + (y - c) ** 2) - 90.2
    tor = torch.sqrt(q * q + (z - c) ** 2) - 30.6
    f = torch.minimum(shell, tor)
    del z, y, x, d, q, shell, tor
    m = extract_triangle_mesh(f, voxel_size=0.01)
    V, Fn = m.vertices.shape[0], m.triangles.shape[0]
    tri = m.triangles.long()
    assert int(tri.min()) == 0 and int(tri.max()) == V - 1
    assert torch.unique(tri).numel() == V
    # every directed edge once, its reverse once
    a = torch.cat([tri[:, 0], tri[:, 1], tri[:, 2]])
    b = torch.cat([tri[:, 1], tri[:, 2], tri[:, 0]])
    fwd, bwd = torch.unique(a * V + b), torch.unique(b * V + a)
    assert fwd.numel() == 3 * Fn and torch.equal(fwd, bwd)
    E = 3 * Fn // 2
    assert V - E + Fn == 2 + 2 + 0                                        # two spheres + one torus
    p = m.vertices.double()[tri]
    vol = float((p[:, 0] * torch.cross(p[:, 1], p[:, 2], dim=1)).sum() / 6.0)
    want = (4 / 3 * np.pi * (193.0 ** 3 - 167.6 ** 3) + 2 * np.pi ** 2 * 90.2 * 30.6 ** 2) * 1e-6
    assert 0.995 < vol / want < 1.001                                     # positive: outward


def test_bounded_volume_extracts_its_mesh():
    from gsr_b200.tsdf import BoundedTSDFVolume
    from oracle import mcubes_oracle as mc
    rng = np.random.default_rng(3)
    vol = BoundedTSDFVolume((-0.4, -0.3, 0.2), 0.02, (30, 28, 26), 0.1, 5.0, with_rgb=True)
    f, w, rgb = ms.observed_blob((26, 28, 30), 11)
    vol.tsdf.copy_(torch.from_numpy(f))
    vol.weight.copy_(torch.from_numpy(w))
    vol.rgb.copy_(torch.from_numpy(rgb))
    vol._fresh = False
    m = vol.extract_triangle_mesh()
    torch.cuda.synchronize()
    org = np.array([-0.4, -0.3, 0.2], dtype=np.float32)
    _same(m.numpy(), mc.extract(f, w, 1.0, 0.0, org, np.float32(0.02), rgb))


def test_marching_cubes_with_contraction_meshes_the_chunked_lattice_in_one_piece():
    """mcube_utils.py:17-110 with two chunks per axis: the lattice is the reference's (shared chunk boundaries once), the
    mesh equals the oracle's on the same samples, is closed across the chunk seams and points outwards."""
    from gsr_b200.mesh import marching_cubes_with_contraction
    from oracle import mcubes_oracle as mc
    seen = []

    def field(p):
        return (p - torch.tensor([0.1, -0.05, 0.2], device=p.device)).norm(dim=-1) - 0.6 + 0.05 * torch.sin(9 * p[:, 0])

    def sdf(p):
        assert p.is_cuda and p.shape[1] == 3 and p.shape[0] <= 40_000
        seen.append(p)
        return field(p)

    lo, hi = (-1.0, -0.9, -0.8), (0.9, 1.0, 1.1)
    mesh = marching_cubes_with_contraction(sdf, resolution=64, bounding_box_min=lo, bounding_box_max=hi, crop=32, points_per_call=40_000)
    pts = torch.cat(seen)
    G = 2 * 31 + 1
    assert pts.shape[0] == G ** 3
    lattice = torch.cat([field(q) for q in seen]).reshape(G, G, G)
    xs = pts.reshape(G, G, G, 3)[:, 0, 0, 0].cpu().numpy()
    assert xs[0] == np.float32(lo[0]) and xs[-1] == np.float32(hi[0]) and (np.diff(xs) > 0).all()       # the seam sample once
    ov, of, _ = mc.extract(lattice.cpu().numpy())
    spacing = np.array([(hi[k] - lo[k]) / 2 / 31 for k in range(3)], dtype=np.float32)
    want_v = ov[:, [2, 1, 0]] * spacing + np.array(lo, dtype=np.float32)
    gv, gf, _ = mesh.numpy()
    assert np.array_equal(gf, of[:, [0, 2, 1]]) and np.array_equal(gv, want_v.astype(np.float32))
    assert all(u == [1, 1] for u in mc.edge_use_counts(gf).values())
    p = gv[gf].astype(np.float64)
    assert np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() > 0
    # the contraction hook and the clip
    far = marching_cubes_with_contraction(field, resolution=64, bounding_box_min=lo, bounding_box_max=hi, crop=32,
                                          inv_contraction=lambda v: v * 100.0, max_range=32.0)
    assert float(far.vertices.abs().max()) == 32.0 and torch.equal(far.triangles, mesh.triangles)
    empty = marching_cubes_with_contraction(lambda q: torch.ones(q.shape[0], device=q.device), resolution=32, crop=32)
    assert empty.vertices.shape == (0, 3) and empty.triangles.shape == (0, 3)


def test_extract_mesh_unbounded_runs_the_whole_chain():
    """TSDFFusion.extract_mesh_unbounded (mesh_utils.py:181-277): contracted lattice -> fusion -> marching cubes -> colours."""
    from gsr_b200.mesh import post_process_mesh
    from gsr_b200.tsdf import TSDFFusion
    from tsdf_synth import build_tsdf_case
    c = build_tsdf_case("contracted", n=16)
    f = TSDFFusion([torch.from_numpy(m) for m in c["projs"]], [torch.from_numpy(d) for d in c["depthmaps"]],
                   [torch.from_numpy(r) for r in c["rgbmaps"]], center=c["center"], radius=c["radius"])
    xyz = torch.from_numpy(c["center"]).cuda() + 0.5 * c["radius"] * torch.randn(5000, 3, device="cuda")
    mesh = f.extract_mesh_unbounded(resolution=128, gaussians_xyz=xyz, crop=64)
    V, F = mesh.vertices.shape[0], mesh.triangles.shape[0]
    assert F > 1000 and mesh.vertex_colors.shape == (V, 3)
    assert torch.isfinite(mesh.vertices).all() and float(mesh.vertex_colors.min()) >= 0.0 and float(mesh.vertex_colors.max()) <= 1.0 + 1e-6
    assert int(mesh.triangles.min()) == 0 and int(mesh.triangles.max()) == V - 1
    # vertices are the uncontracted lattice crossings: mapping them back gives tsdf ~ 0 there (on the [-1, 1] scale of the truncated field)
    back = f.compute_unbounded_tsdf(f.contract_normalized(mesh.vertices), True, f.radius * 2 / 128)
    assert float(back.abs().median()) < 0.2
    post = post_process_mesh(mesh, cluster_to_keep=3)
    assert 0 < post.triangles.shape[0] <= F


@pytest.mark.parametrize("name", ["noise", "observed"])
def test_product_reproduces_the_frozen_conventions(name):
    """The committed vectors of tests/golden/make_golden_mesh.py through the C ABI: mesh and cleaned mesh, bit for bit."""
    import os
    import sys
    from gsr_b200.mesh import extract_triangle_mesh, post_process_mesh
    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold_dir)
    from make_golden_mesh import build
    g = np.load(os.path.join(gold_dir, f"mesh_{name}.npz"))
    c, f, w, rgb = build(name)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m = extract_triangle_mesh(t(f), t(w), 1.0 if w is not None else None, c["level"], c["origin"], c["voxel"], t(rgb))
    v, faces, col = m.numpy()
    assert np.array_equal(v, g["verts"]) and np.array_equal(faces, g["faces"])
    pv, pf, pc = post_process_mesh(m, cluster_to_keep=2, min_triangles=4).numpy()
    assert np.array_equal(pv, g["post_verts"]) and np.array_equal(pf, g["post_faces"])
    if rgb is not None:
        assert np.array_equal(col, g["colors"]) and np.array_equal(pc, g["post_colors"])
