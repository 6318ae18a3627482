"""GPU: gsr_mesh_clusters / gsr_mesh_keep_clusters / gsr_mesh_filter_* (through gsr_b200.mesh -> C ABI) against the numpy +
scipy restatement of the reference's post_process_mesh (mesh_utils.py:27-49; Open3D absent, parity unpinned): integer
results exactly, cluster areas to rounding; and the whole mesh path volume -> mesh -> cleaned mesh -> PLY on the device."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import mesh_synth as ms  # noqa: E402
from test_mesh_clusters_oracle import _blobs  # noqa: E402


def _mesh(v, f, c=None):
    from gsr_b200.mesh import TriangleMesh
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return TriangleMesh(t(v), t(f), t(c))


@pytest.mark.parametrize("field", ["blobs", "noise", "torus"])
def test_clusters_match_the_oracle(field):
    from gsr_b200.mesh import _clusters, cluster_connected_triangles
    from oracle import mcubes_oracle as mc
    from oracle import mesh_clusters_oracle as mo
    f = {"blobs": _blobs(), "noise": ms.noise((40, 41, 42), 5), "torus": ms.torus()}[field]
    v, faces, _ = mc.extract(f)
    vroot, troot, ntris, area = _clusters(_mesh(v, faces), True)
    ovroot, otroot, ontris, oarea = mo.clusters(v, faces)
    assert np.array_equal(vroot.cpu().numpy(), ovroot) and np.array_equal(troot.cpu().numpy(), otroot)
    assert np.array_equal(ntris.cpu().numpy().astype(np.int64), ontris)
    assert np.allclose(area.cpu().numpy(), oarea, rtol=1e-9, atol=1e-12)
    tc, n, a = cluster_connected_triangles(_mesh(v, faces))
    roots = np.nonzero(ontris)[0]
    assert np.array_equal(n, ontris[roots]) and np.allclose(a, oarea[roots], rtol=1e-9)
    assert np.array_equal(roots[tc.cpu().numpy()], otroot)


def test_a_long_chain_of_triangles_is_one_cluster():
    """A strip whose vertex ids run against the hooking direction: deep union-find trees before flattening."""
    from gsr_b200.mesh import _clusters
    n = 200_000
    rng = np.random.default_rng(0)
    perm = rng.permutation(n + 2).astype(np.int32)
    f = np.stack([perm[:-2], perm[1:-1], perm[2:]], axis=1)
    f = f[rng.permutation(n)]
    v = rng.random((n + 2, 3)).astype(np.float32)
    vroot, troot, ntris, _ = _clusters(_mesh(v, f), False)
    assert int(vroot.max()) == 0 and int(troot.max()) == 0 and int(ntris[0]) == n and int(ntris[1:].max()) == 0


@pytest.mark.parametrize("keep_n", [1, 2, 1000])
def test_post_process_mesh_matches_the_oracle(keep_n):
    from gsr_b200.mesh import post_process_mesh
    from oracle import mcubes_oracle as mc
    from oracle import mesh_clusters_oracle as mo
    v, faces, _ = mc.extract(_blobs())
    col = np.random.default_rng(1).random(v.shape).astype(np.float32)
    got = post_process_mesh(_mesh(v, faces, col), cluster_to_keep=keep_n).numpy()
    want = mo.post_process_mesh(v, faces, col, cluster_to_keep=keep_n)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)
    got = post_process_mesh(_mesh(v, faces), cluster_to_keep=keep_n).numpy()
    assert got[2] is None and np.array_equal(got[1], want[1])


def test_filter_order_of_operations_and_empty_meshes():
    from gsr_b200.mesh import post_process_mesh, remove_triangles_by_mask
    v = np.arange(18, dtype=np.float32).reshape(6, 3)
    f = np.array([[0, 1, 2], [2, 3, 3], [3, 4, 5], [1, 2, 4]], dtype=np.int32)
    v2, f2, _ = remove_triangles_by_mask(_mesh(v, f), torch.tensor([1, 1, 0, 1], dtype=torch.bool, device="cuda")).numpy()
    assert np.array_equal(v2, v[[0, 1, 2, 3, 4]]) and np.array_equal(f2, [[0, 1, 2], [1, 2, 4]])
    v3, f3, _ = remove_triangles_by_mask(_mesh(v, f), torch.zeros(4, dtype=torch.uint8, device="cuda")).numpy()
    assert v3.shape == (0, 3) and f3.shape == (0, 3)
    empty = _mesh(np.zeros((0, 3), dtype=np.float32), np.zeros((0, 3), dtype=np.int32))
    out = post_process_mesh(empty)
    assert out.vertices.shape == (0, 3) and out.triangles.shape == (0, 3)
    bad = _mesh(v, np.array([[0, 1, 7]], dtype=np.int32))
    with pytest.raises(RuntimeError, match="outside"):
        post_process_mesh(bad)


def test_volume_to_cleaned_ply_on_the_device(tmp_path):
    """TSDF lattice -> marching cubes -> cluster filter -> PLY, nothing but the final arrays leaves the GPU."""
    from gsr_b200.mesh import extract_triangle_mesh, post_process_mesh
    from oracle import mcubes_oracle as mc
    from oracle import mesh_clusters_oracle as mo
    f = _blobs(seed=4, n=48)
    rgb = np.random.default_rng(2).random(f.shape + (3,)).astype(np.float32)
    mesh = extract_triangle_mesh(torch.from_numpy(f).cuda(), voxel_size=0.05, origin=(-1.0, -1.0, -1.0), rgb=torch.from_numpy(rgb).cuda())
    post = post_process_mesh(mesh, cluster_to_keep=3)
    v, faces, col = mc.extract(f, origin=(-1.0, -1.0, -1.0), voxel_size=0.05, rgb=rgb)
    wv, wf, wc = mo.post_process_mesh(v, faces, col, cluster_to_keep=3)
    gv, gf, gc = post.numpy()
    assert np.array_equal(gv, wv) and np.array_equal(gf, wf) and np.array_equal(gc, wc)
    path = tmp_path / "fuse_post.ply"
    post.write_ply(str(path))
    raw = path.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    assert f"element vertex {len(wv)}".encode() in head and f"element face {len(wf)}".encode() in head
    assert len(body) == len(wv) * (3 * 8 + 3) + len(wf) * (1 + 3 * 4)
    xyz = np.frombuffer(body[:27], dtype="<f8", count=3)
    assert np.allclose(xyz, wv[0].astype(np.float64))
