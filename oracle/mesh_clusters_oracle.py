"""TEST INFRASTRUCTURE -- CPU restatement (numpy + scipy) of the mesh clean-up GS-SR runs after marching cubes,
/root/reference/gssr/utils/mesh_utils.py:27-49 (post_process_mesh), which calls Open3D 0.18.0 (requirements.txt:6, absent
here: PARITY UNPINNED against Open3D itself):
    mesh.cluster_connected_triangles()      triangles joined through shared mesh edges, (cluster per triangle, triangles
                                            per cluster, area per cluster)
    mesh.remove_triangles_by_mask(mask); mesh.remove_unreferenced_vertices(); mesh.remove_degenerate_triangles()
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product
(gs-sr_b200/csrc/mesh_clusters.cu) never does.

Clusters here are the connected components of the VERTEX graph (scipy.sparse.csgraph.connected_components), labelled by
their smallest vertex id -- what the product computes.  They equal Open3D's edge-connected clusters unless two sheets
touch in a single vertex; `edge_connected_clusters` computes Open3D's definition so tests can check the two agree on
marching-cubes output.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def clusters(verts, faces):
    """vertex_root (V,), tri_root (F,), root_ntris (V,), root_area (V,) float64."""
    verts = np.asarray(verts, dtype=np.float32)
    faces = np.asarray(faces, dtype=np.int64)
    V = verts.shape[0]
    a = np.concatenate([faces[:, 0], faces[:, 0]])
    b = np.concatenate([faces[:, 1], faces[:, 2]])
    g = coo_matrix((np.ones(a.shape[0], dtype=np.int8), (a, b)), shape=(V, V))
    _, lab = connected_components(g, directed=False)
    first = np.full(lab.max() + 1 if V else 0, V, dtype=np.int64)
    np.minimum.at(first, lab, np.arange(V))
    vroot = first[lab] if V else np.zeros(0, dtype=np.int64)
    troot = vroot[faces[:, 0]] if faces.shape[0] else np.zeros(0, dtype=np.int64)
    ntris = np.bincount(troot, minlength=V)[:V] if V else np.zeros(0, dtype=np.int64)
    p = verts.astype(np.float64)[faces]
    area = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1) if faces.shape[0] else np.zeros(0)
    rarea = np.bincount(troot, weights=area, minlength=V)[:V] if V else np.zeros(0)
    return vroot.astype(np.int32), troot.astype(np.int32), ntris.astype(np.int64), rarea


def edge_connected_clusters(faces):
    """Open3D's definition: label per triangle, triangles adjacent when they share an (undirected) edge."""
    faces = np.asarray(faces, dtype=np.int64)
    F = faces.shape[0]
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
    e.sort(axis=1)
    tri = np.tile(np.arange(F), 3)
    key = e[:, 0] * (faces.max() + 1 if F else 1) + e[:, 1]
    order = np.argsort(key, kind="stable")
    key, tri = key[order], tri[order]
    same = key[1:] == key[:-1]
    g = coo_matrix((np.ones(int(same.sum()), dtype=np.int8), (tri[:-1][same], tri[1:][same])), shape=(F, F))
    return connected_components(g, directed=False)[1]


def remove_triangles_by_mask(verts, faces, colors, keep):
    """Order-preserving: triangles with keep == 0 go, then unreferenced vertices, then triangles with a repeated index."""
    verts, faces = np.asarray(verts), np.asarray(faces)
    keep = np.asarray(keep).astype(bool)
    f = faces[keep]
    ref = np.zeros(verts.shape[0], dtype=bool)
    ref[f.ravel()] = True
    new = np.cumsum(ref) - 1
    f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]
    return verts[ref], new[f].astype(np.int32), (None if colors is None else np.asarray(colors)[ref])


def post_process_mesh(verts, faces, colors=None, cluster_to_keep=1000, min_triangles=50):
    """mesh_utils.py:27-49 (with the product's documented behaviour when there are fewer clusters than cluster_to_keep)."""
    _, troot, ntris, _ = clusters(verts, faces)
    sizes = np.sort(ntris[ntris > 0])
    n_cluster = int(sizes[-cluster_to_keep]) if sizes.shape[0] >= cluster_to_keep else 0
    n_cluster = max(n_cluster, min_triangles)
    keep = ntris[troot] >= n_cluster
    return remove_triangles_by_mask(verts, faces, colors, keep)
