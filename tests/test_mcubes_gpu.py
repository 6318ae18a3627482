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
