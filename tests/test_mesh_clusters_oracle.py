"""CPU: the mesh clean-up restatement (oracle/mesh_clusters_oracle.py) against hand-checkable cases, and vertex-connected
clusters == Open3D's edge-connected clusters on marching-cubes output."""
import numpy as np

from oracle import mcubes_oracle as mc
from oracle import mesh_clusters_oracle as mo
import mesh_synth as ms


def _blobs(seed=0, n=36):
    """A few separate closed blobs of different sizes in one lattice."""
    rng = np.random.default_rng(seed)
    x, y, z = ms.lattice_coords(n, n, n)
    f = np.full((n, n, n), 10.0, dtype=np.float32)
    for c, r in (((8, 8, 8), 5.2), ((25, 9, 10), 3.1), ((10, 26, 24), 6.3), ((27, 27, 8), 1.2), ((26, 26, 27), 2.4)):
        f = np.minimum(f, (np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - np.float32(r)).astype(np.float32))
    return f + 0.01 * rng.standard_normal(f.shape).astype(np.float32)


def test_clusters_of_separate_blobs():
    v, f, _ = mc.extract(_blobs())
    vroot, troot, ntris, area = mo.clusters(v, f)
    roots = np.nonzero(ntris)[0]
    assert len(roots) == 5 and ntris.sum() == len(f)
    assert (vroot[roots] == roots).all() and (vroot <= np.arange(len(v))).all()
    assert (troot == vroot[f[:, 1]]).all() and (troot == vroot[f[:, 2]]).all()
    # area of a sphere of radius r, within the marching-cubes discretisation
    big = roots[np.argmax(ntris[roots])]
    assert abs(area[big] / (4 * np.pi * 6.3 ** 2) - 1) < 0.03
    # the same partition as edge adjacency
    e = mo.edge_connected_clusters(f)
    pairs = {(int(a), int(b)) for a, b in zip(troot, e)}
    assert len(pairs) == 5


def test_vertex_and_edge_clusters_agree_on_marching_cubes_noise():
    v, f, _ = mc.extract(ms.noise((14, 15, 16), 3))
    _, troot, _, _ = mo.clusters(v, f)
    e = mo.edge_connected_clusters(f)
    assert len({(int(a), int(b)) for a, b in zip(troot, e)}) == len(set(troot.tolist())) == len(set(e.tolist()))


def test_post_process_keeps_the_largest_clusters():
    v, f, _ = mc.extract(_blobs())
    col = np.random.default_rng(1).random(v.shape).astype(np.float32)
    _, troot, ntris, _ = mo.clusters(v, f)
    sizes = np.sort(ntris[ntris > 0])
    v2, f2, c2 = mo.post_process_mesh(v, f, col, cluster_to_keep=2, min_triangles=1)
    assert len(f2) == sizes[-1] + sizes[-2] and len(np.unique(f2)) == len(v2) and c2.shape == v2.shape
    assert np.array_equal(v2[f2], v[f[ntris[troot] >= sizes[-2]]])          # same triangles, same order, re-indexed
    v3, f3, _ = mo.post_process_mesh(v, f, None, cluster_to_keep=1000)       # fewer clusters than asked: the >= 50 rule
    assert len(f3) == sizes[sizes >= 50].sum()


def test_filter_order_of_operations():
    v = np.arange(18, dtype=np.float32).reshape(6, 3)
    f = np.array([[0, 1, 2], [2, 3, 3], [3, 4, 5], [1, 2, 4]], dtype=np.int32)
    v2, f2, _ = mo.remove_triangles_by_mask(v, f, None, [1, 1, 0, 1])
    # vertex 3 is referenced only by the degenerate triangle: it stays (unreferenced vertices go first), the triangle goes
    assert np.array_equal(v2, v[[0, 1, 2, 3, 4]]) and np.array_equal(f2, [[0, 1, 2], [1, 2, 4]])
    v3, f3, _ = mo.remove_triangles_by_mask(v, f, None, [0, 0, 0, 0])
    assert v3.shape == (0, 3) and f3.shape == (0, 3)
