"""Triangle mesh out of a fused TSDF lattice -- host-side mirror of what the reference does with its volume after the
integration loop (/root/reference/gssr/utils/mesh_utils.py:178, /root/reference/extract_mesh_split.py:119-128,
/root/reference/extract_mesh.py:125-134):

    mesh = volume.extract_triangle_mesh()                     # Open3D ScalableTSDFVolume
    o3d.io.write_triangle_mesh(path, mesh)

and of the host marching cubes of the unbounded path (gssr/utils/mcube_utils.py:71-80).  Here:

    vol = BoundedTSDFVolume(...); vol.integrate(...); vol.reduce_to(0)
    mesh = vol.extract_triangle_mesh()                        # or extract_triangle_mesh(tsdf, ...)
    mesh.write_ply(path)
    mesh_post = post_process_mesh(mesh, cluster_to_keep=1000) # mesh_utils.py:27-49

Marching cubes runs on the GPU the volume lives on (``gsr_mc_count`` / ``gsr_mc_emit``, gs-sr_b200/csrc/mcubes.cu); the
lattice never travels to the host.  Open3D and skimage are absent third-party dependencies without vectors in the
reference: parity against them is UNPINNED (conventions follow Open3D's extractor, include/gsr_b200.h).  No CPU /
PyTorch fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import check, lib
from ._torch_util import f32c, on_device, stream_ptr


@dataclass
class TriangleMesh:
    """The three arrays of o3d.geometry.TriangleMesh the reference's mesh path touches, as device tensors."""
    vertices: torch.Tensor                          # (V, 3) float32
    triangles: torch.Tensor                         # (F, 3) int32, normals towards tsdf > level (free space)
    vertex_colors: Optional[torch.Tensor] = None    # (V, 3) float32 in [0, 1]

    def numpy(self):
        c = None if self.vertex_colors is None else self.vertex_colors.cpu().numpy()
        return self.vertices.cpu().numpy(), self.triangles.cpu().numpy(), c

    def write_ply(self, path):
        """Binary little-endian PLY with the element / property layout o3d.io.write_triangle_mesh produces for a mesh
        with vertex colours (double x y z, uchar red green blue, list uchar uint vertex_indices)."""
        v, f, c = self.numpy()
        props = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
        head = ["ply", "format binary_little_endian 1.0", "comment Created by gsr_b200", f"element vertex {len(v)}",
                "property double x", "property double y", "property double z"]
        if c is not None:
            props += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
            head += ["property uchar red", "property uchar green", "property uchar blue"]
        head += [f"element face {len(f)}", "property list uchar uint vertex_indices", "end_header"]
        vert = np.empty(len(v), dtype=props)
        vert["x"], vert["y"], vert["z"] = v[:, 0], v[:, 1], v[:, 2]
        if c is not None:
            c8 = np.clip(c * 255.0, 0.0, 255.0).astype(np.uint8)
            vert["red"], vert["green"], vert["blue"] = c8[:, 0], c8[:, 1], c8[:, 2]
        face = np.empty(len(f), dtype=[("n", "u1"), ("i", "<u4", (3,))])
        face["n"] = 3
        face["i"] = f.astype(np.uint32)
        with open(path, "wb") as fh:
            fh.write(("\n".join(head) + "\n").encode("ascii"))
            fh.write(vert.tobytes())
            fh.write(face.tobytes())


def _i32c(t, name, dev):
    if not t.is_cuda or t.device != dev:
        raise RuntimeError(f"{name} must be a CUDA tensor on {dev}")
    return t.to(torch.int32).contiguous()


@torch.no_grad()
def _clusters(mesh, with_area):
    """(vertex_root, tri_root, root_ntris, root_area) device arrays of gsr_mesh_clusters."""
    dev = mesh.vertices.device
    v = f32c(mesh.vertices, "vertices", dev)
    f = _i32c(mesh.triangles, "triangles", dev)
    nv, nf = v.shape[0], f.shape[0]
    vroot = torch.empty(nv, dtype=torch.int32, device=dev)
    troot = torch.empty(nf, dtype=torch.int32, device=dev)
    ntri = torch.empty(nv, dtype=torch.int32, device=dev)
    area = torch.empty(nv, dtype=torch.float64, device=dev) if with_area else None
    with on_device(dev):
        check(lib().gsr_mesh_clusters(nv, nf, v.data_ptr() if with_area and nv else None, f.data_ptr() if nf else None,
                                      vroot.data_ptr() if nv else None, troot.data_ptr() if nf else None,
                                      ntri.data_ptr() if nv else None, area.data_ptr() if with_area and nv else None,
                                      stream_ptr(dev)), "gsr_mesh_clusters")
    return vroot, troot, ntri, area


@torch.no_grad()
def _cluster_sizes(root_ntris, capacity=1 << 16):
    """The non-zero entries of root_ntris as a host array (gathered on the device: a few kB travel, not the whole array)."""
    dev = root_ntris.device
    nv = root_ntris.shape[0]
    if nv == 0:
        return np.zeros(0, dtype=np.int64)
    n = ctypes.c_longlong(0)
    while True:
        buf = torch.empty(capacity, dtype=torch.int32, device=dev)
        with on_device(dev):
            check(lib().gsr_mesh_cluster_sizes(nv, root_ntris.data_ptr(), buf.data_ptr(), capacity, ctypes.byref(n),
                                               stream_ptr(dev)), "gsr_mesh_cluster_sizes")
        if n.value <= capacity - 1:
            return buf[:n.value].cpu().numpy().astype(np.int64)
        capacity = n.value + 2


@torch.no_grad()
def cluster_connected_triangles(mesh):
    """``mesh.cluster_connected_triangles()`` as the reference calls it (mesh_utils.py:34): (triangle_clusters (F,) int32
    device tensor, cluster_n_triangles (C,) numpy int64, cluster_area (C,) numpy float64).  Clusters are numbered by
    their smallest vertex id (Open3D: by their first triangle); post_process_mesh only uses the sizes."""
    _, troot, ntri, area = _clusters(mesh, True)
    n_host = ntri.cpu().numpy()
    roots = np.nonzero(n_host)[0]
    dense = np.full(n_host.shape[0] + 1, -1, dtype=np.int32)
    dense[roots] = np.arange(roots.shape[0], dtype=np.int32)
    tri_clusters = torch.from_numpy(dense).to(troot.device)[troot.long()]
    return tri_clusters, n_host[roots].astype(np.int64), area.cpu().numpy()[roots]


@torch.no_grad()
def remove_triangles_by_mask(mesh, triangles_to_keep):
    """Drops the triangles with a zero in `triangles_to_keep` ((F,) bool / uint8 device tensor), then the vertices nothing
    references, then triangles with a repeated index -- remove_triangles_by_mask / remove_unreferenced_vertices /
    remove_degenerate_triangles of mesh_utils.py:43-45 in one pass; survivors keep their order."""
    dev = mesh.vertices.device
    v = f32c(mesh.vertices, "vertices", dev)
    c = f32c(mesh.vertex_colors, "vertex_colors", dev)
    f = _i32c(mesh.triangles, "triangles", dev)
    keep = triangles_to_keep.to(device=dev, dtype=torch.uint8).contiguous()
    nv, nf = v.shape[0], f.shape[0]
    if keep.shape != (nf,):
        raise ValueError("triangles_to_keep must have one entry per triangle")
    L = lib()
    nvo, nfo = ctypes.c_longlong(0), ctypes.c_longlong(0)
    with on_device(dev):
        work = torch.empty(L.gsr_mesh_filter_workspace_bytes(nv, nf), dtype=torch.uint8, device=dev)
        check(L.gsr_mesh_filter_count(nv, nf, f.data_ptr() if nf else None, keep.data_ptr() if nf else None, work.data_ptr(),
                                      ctypes.byref(nvo), ctypes.byref(nfo), stream_ptr(dev)), "gsr_mesh_filter_count")
        vo = torch.empty((nvo.value, 3), dtype=torch.float32, device=dev)
        fo = torch.empty((nfo.value, 3), dtype=torch.int32, device=dev)
        co = torch.empty((nvo.value, 3), dtype=torch.float32, device=dev) if c is not None else None
        check(L.gsr_mesh_filter_emit(nv, nf, v.data_ptr() if nv else None, c.data_ptr() if c is not None and nv else None,
                                     f.data_ptr() if nf else None, work.data_ptr(), vo.data_ptr(),
                                     co.data_ptr() if co is not None and nv else None, fo.data_ptr(), stream_ptr(dev)),
              "gsr_mesh_filter_emit")
    return TriangleMesh(vo, fo, co)


@torch.no_grad()
def post_process_mesh(mesh, cluster_to_keep=1000, min_triangles=50):
    """``post_process_mesh`` of the reference (mesh_utils.py:27-49): keep the `cluster_to_keep` largest connected clusters,
    none smaller than 50 triangles; drop the vertices they leave behind.  One divergence: with fewer than
    `cluster_to_keep` clusters the reference's ``np.sort(...)[-cluster_to_keep]`` raises IndexError; here every cluster of
    at least `min_triangles` triangles is kept."""
    dev = mesh.vertices.device
    _, troot, ntri, _ = _clusters(mesh, False)
    sizes = np.sort(_cluster_sizes(ntri))
    n_cluster = int(sizes[-cluster_to_keep]) if sizes.shape[0] >= cluster_to_keep else 0
    n_cluster = max(n_cluster, int(min_triangles))
    nf = troot.shape[0]
    keep = torch.empty(nf, dtype=torch.uint8, device=dev)
    with on_device(dev):
        check(lib().gsr_mesh_keep_clusters(nf, troot.data_ptr() if nf else None, ntri.data_ptr() if nf else None, n_cluster,
                                           keep.data_ptr() if nf else None, stream_ptr(dev)), "gsr_mesh_keep_clusters")
    return remove_triangles_by_mask(mesh, keep)


@torch.no_grad()
def extract_triangle_mesh(tsdf, weight=None, min_weight=None, level=0.0, origin=(0.0, 0.0, 0.0), voxel_size=1.0, rgb=None):
    """Marching cubes of a (nz, ny, nx) lattice at `level`.

    weight / min_weight: a cell yields triangles only when all eight corners have weight > min_weight (Open3D's
    "observed" test; BoundedTSDFVolume starts at weight 1, so min_weight = 1 there).  Without them every cell counts
    (what skimage does for the unbounded path).  rgb: (nz, ny, nx, 3) colours, interpolated onto the vertices."""
    if tsdf.dim() != 3:
        raise ValueError("tsdf must be a (nz, ny, nx) lattice")
    if not tsdf.is_cuda:
        raise RuntimeError("extract_triangle_mesh needs CUDA tensors (gsr_b200 has no CPU path)")
    if (weight is None) != (min_weight is None):
        raise ValueError("weight and min_weight go together")
    dev = tsdf.device
    nz, ny, nx = tsdf.shape
    f = f32c(tsdf, "tsdf", dev)
    w = f32c(weight, "weight", dev)
    c = f32c(rgb, "rgb", dev)
    if w is not None and w.shape != f.shape:
        raise ValueError("weight must have the shape of tsdf")
    if c is not None and tuple(c.shape) != (nz, ny, nx, 3):
        raise ValueError("rgb must be (nz, ny, nx, 3)")
    L = lib()
    org = (ctypes.c_float * 3)(*[float(v) for v in origin])
    nv, nt = ctypes.c_longlong(0), ctypes.c_longlong(0)
    with on_device(dev):
        work = torch.empty(L.gsr_mc_workspace_bytes(nx, ny, nz), dtype=torch.uint8, device=dev)
        check(L.gsr_mc_count(nx, ny, nz, f.data_ptr(), w.data_ptr() if w is not None else None,
                             float(min_weight) if min_weight is not None else 0.0, float(level), work.data_ptr(),
                             ctypes.byref(nv), ctypes.byref(nt), stream_ptr(dev)), "gsr_mc_count")
        verts = torch.empty((nv.value, 3), dtype=torch.float32, device=dev)
        faces = torch.empty((nt.value, 3), dtype=torch.int32, device=dev)
        colors = torch.empty((nv.value, 3), dtype=torch.float32, device=dev) if c is not None else None
        if nv.value or nt.value:
            check(L.gsr_mc_emit(nx, ny, nz, f.data_ptr(), c.data_ptr() if c is not None else None, float(level), org,
                                float(voxel_size), work.data_ptr(), verts.data_ptr(),
                                colors.data_ptr() if colors is not None else None, faces.data_ptr(), stream_ptr(dev)),
                  "gsr_mc_emit")
    return TriangleMesh(verts, faces, colors)


def _chunked_axis(lo, hi, n_chunks, crop, device):
    """The sample coordinates of one axis as the reference lays them out (mcube_utils.py:36-52): `n_chunks` chunks between
    lo and hi, each ``torch.linspace(chunk_min, chunk_max, crop)``; neighbouring chunks share their boundary sample, which
    appears once here."""
    edges = np.linspace(lo, hi, n_chunks + 1)
    parts = []
    for i in range(n_chunks):
        ax = torch.linspace(float(edges[i]), float(edges[i + 1]), crop, device=device)
        parts.append(ax if i == n_chunks - 1 else ax[:-1])
    return torch.cat(parts)


@torch.no_grad()
def marching_cubes_with_contraction(sdf, resolution=512, bounding_box_min=(-1.0, -1.0, -1.0), bounding_box_max=(1.0, 1.0, 1.0),
                                    level=0.0, inv_contraction=None, max_range=32.0, crop=512, device=None,
                                    points_per_call=256 ** 3):
    """``marching_cubes_with_contraction`` of the reference (gssr/utils/mcube_utils.py:17-110): sample `sdf` (a callable on
    (n, 3) CUDA points, n <= points_per_call as in :57-63) on the same lattice points, mesh the level set, map the
    vertices through `inv_contraction` and clip them to +-max_range.

    The reference meshes one 512^3 chunk at a time on the host (device->host copy, skimage, trimesh concatenate +
    merge_vertices).  Here the whole lattice -- (resolution/crop * (crop-1) + 1)^3 samples, 4.3 GB at resolution 1024 -- stays
    in HBM and is meshed in one piece on the GPU: no seams to merge, no host copy."""
    if resolution % crop != 0:
        raise ValueError("resolution must be a multiple of crop (the reference asserts resolution % 512 == 0)")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    n_chunks = resolution // crop
    ax = [_chunked_axis(bounding_box_min[k], bounding_box_max[k], n_chunks, crop, device) for k in range(3)]
    G = ax[0].shape[0]
    lattice = torch.empty((G, G, G), dtype=torch.float32, device=device)          # [ix, iy, iz], the reference's volume layout
    step = max(1, int(points_per_call) // (G * G))
    for i0 in range(0, G, step):
        i1 = min(G, i0 + step)
        xx, yy, zz = torch.meshgrid(ax[0][i0:i1], ax[1], ax[2], indexing="ij")
        pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
        lattice[i0:i1] = sdf(pts).reshape(i1 - i0, G, G)
        del xx, yy, zz, pts
    # the extractor's fastest axis is this layout's z: it sees the lattice as (nz', ny', nx') = (x, y, z) and returns
    # lattice coordinates (x', y', z') = (iz, iy, ix); renaming the axes back is a reflection, so the winding flips too
    m = extract_triangle_mesh(lattice, level=level)
    del lattice
    spacing = torch.tensor([(bounding_box_max[k] - bounding_box_min[k]) / n_chunks / (crop - 1) for k in range(3)],
                           dtype=torch.float32, device=device)
    origin = torch.tensor([float(v) for v in bounding_box_min], dtype=torch.float32, device=device)
    verts = m.vertices[:, [2, 1, 0]] * spacing + origin
    faces = m.triangles[:, [0, 2, 1]].contiguous()
    if inv_contraction is not None:
        verts = inv_contraction(verts).clamp_(-max_range, max_range)
    return TriangleMesh(verts.contiguous(), faces)
