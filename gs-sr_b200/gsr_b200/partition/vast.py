"""VastGaussian scene partition on the CPU -- SURVEY.md section 8(f4), BASELINE config 1.

Restates, array-at-a-time, the four partition steps and the tile writer GS-SR runs before per-tile training
(/root/reference/gssr/utils/vastgaussian_utils.py, /root/reference/split_scene.py):

  1. camera_position_based_region_division   (:89-149)  cameras sorted along x then y into num_col x num_row tiles, or a
                                                        median split of the longer axis until a tile has < max_num_images
  2. position_based_data_selection           (:152-181) every image / point whose camera centre / xyz lies in the tile box
                                                        grown by `ratio` (the stored box stays the un-grown one)
  3. visibility_based_camera_selection       (:184-286) outside cameras that see the tile's 3-D bounding box: area of the
                                                        projected box's convex hull inside the image > threshold, and the
                                                        camera is closer than 1.2 x the tile's own mean camera distance
  4. coverage_based_point_selection          (:289-303) the tile's points become all points observed by its cameras
  + transform_colmap (:15-77), box.txt and the tile_%04d COLMAP models of split_scene.py:55-82.

Everything is numpy on the host (no torch, no GPU): the reference's one CUDA call here, distCUDA2 for the z-range of a
tile, is served by knn_cpu.dist2_knn3_cpu; its scipy ConvexHull + shapely intersection by a monotone-chain hull and a
Sutherland-Hodgman clip below.  Reference quirks that decide tile membership are kept and named where they occur:
  * grid mode drops the trailing  N mod num_col  cameras of the x-ordering (and likewise per column) (:123-137);
  * the camera-to-box distance of step 3 is an L1 distance -- sum(sqrt(d^2)) per corner (:264);
  * points behind a camera are projected like any other (no depth test) (:214-226).
Results are compared with the reference's own functions on the 64-camera synthetic model in tests/test_partition_cpu.py
(golden tile membership / boxes produced by tests/golden/make_golden_partition.py).
"""
from __future__ import annotations

import os
import shutil
from typing import Dict, List, Optional

import numpy as np

from .colmap_io import Image, Point3D, qvecs2rotmats, read_model, rotmat2qvec, write_model
from .knn_cpu import dist2_knn3_cpu


# ---- camera geometry ------------------------------------------------------------------------------------------------
def w2c_matrices(images: Dict[int, Image]):
    """(ids, (N,4,4) world-to-camera matrices) in dict order (get_w2c_matrix, :184-191)."""
    ids = np.fromiter(images.keys(), dtype=np.int64, count=len(images))
    M = np.zeros((len(ids), 4, 4))
    if len(ids):
        M[:, :3, :3] = qvecs2rotmats(np.stack([images[i].qvec for i in ids]))
        M[:, :3, 3] = np.stack([images[i].tvec for i in ids])
    M[:, 3, 3] = 1.0
    return ids, M


def camera_centers(images: Dict[int, Image]):
    """(ids, (N,3) camera centres) = inverse(w2c)[:3, 3] (get_cam_center, :79-86; same LAPACK inverse per matrix)."""
    ids, M = w2c_matrices(images)
    C = np.linalg.inv(M)[:, :3, 3] if len(ids) else np.zeros((0, 3))
    return ids, C


def _bbox_xy(C):
    return np.array([C[:, 0].min(), C[:, 0].max(), C[:, 1].min(), C[:, 1].max()])


# ---- step 1 -------------------------------------------------------------------------------------------------------------
def camera_position_based_region_division(images, num_col: Optional[int] = None, num_row: Optional[int] = None,
                                          max_num_images: int = 150) -> List[dict]:
    """-> [{"images": [Image...], "box": [mx, Mx, my, My]}] in the reference's tile order."""
    ids, C = camera_centers(images)
    groups: List[np.ndarray] = []          # index arrays into ids / C
    if num_col is None or num_row is None:
        def split(sel):
            ext = _bbox_xy(C[sel])
            axis = 0 if (ext[1] - ext[0]) > (ext[3] - ext[2]) else 1
            order = sel[np.argsort(C[sel, axis], kind="stable")]
            half = len(order) // 2
            for part in (order[:half], order[half:]):
                if len(part) < max_num_images:
                    groups.append(part)
                else:
                    split(part)
        split(np.arange(len(ids)))
    else:
        n = len(ids)
        per_col = n // num_col
        by_x = np.argsort(C[:, 0], kind="stable")
        for i in range(num_col):
            stop = (i + 1) * per_col if (i + 1) * per_col < n else n
            col = by_x[i * per_col: stop]
            m = len(col)
            per_tile = m // num_row
            by_y = col[np.argsort(C[col, 1], kind="stable")]
            for j in range(num_row):
                stop_j = (j + 1) * per_tile if (j + 1) * per_tile < m else m
                groups.append(by_y[j * per_tile: stop_j])
    return [{"images": [images[int(ids[k])] for k in g], "box": _bbox_xy(C[g])} for g in groups]


# ---- step 2 -------------------------------------------------------------------------------------------------------------
def _points_xyz(points3D: Dict[int, Point3D]):
    ids = np.fromiter(points3D.keys(), dtype=np.int64, count=len(points3D))
    xyz = np.stack([points3D[i].xyz for i in ids]) if len(ids) else np.zeros((0, 3))
    return ids, xyz


def _in_box_xy(P, box):
    return (P[:, 0] >= box[0]) & (P[:, 0] <= box[1]) & (P[:, 1] >= box[2]) & (P[:, 1] <= box[3])


def position_based_data_selection(tiles, images, points3D, ratio: float = 0.2) -> List[dict]:
    img_ids, C = camera_centers(images)
    pt_ids, X = _points_xyz(points3D)
    out = []
    for tile in tiles:
        mx, Mx, my, My = tile["box"]
        dw, dh = (Mx - mx) * ratio / 2.0, (My - my) * ratio / 2.0
        grown = np.array([mx - dw, Mx + dw, my - dh, My + dh])
        out.append({"images": [images[int(i)] for i in img_ids[_in_box_xy(C, grown)]], "box": tile["box"],
                    "points3D": [points3D[int(i)] for i in pt_ids[_in_box_xy(X, grown)]]})
    return out


# ---- step 3 -------------------------------------------------------------------------------------------------------------
def _intrinsics(cam):
    if cam.model == "SIMPLE_PINHOLE":
        fx = fy = cam.params[0]
    elif cam.model == "PINHOLE":
        fx, fy = cam.params[0], cam.params[1]
    else:
        raise AssertionError("Colmap camera model not handled: only undistorted datasets (PINHOLE or SIMPLE_PINHOLE cameras) "
                             "supported!")
    return np.array([[fx, 0.0, cam.width / 2.0], [0.0, fy, cam.height / 2.0], [0, 0, 1]])


def convex_hull_2d(pts):
    """Andrew's monotone chain; vertices in counter-clockwise order, collinear points dropped."""
    P = np.unique(np.asarray(pts, np.float64), axis=0)
    if len(P) < 3:
        return P
    P = P[np.lexsort((P[:, 1], P[:, 0]))]

    def half(seq):
        h = []
        for p in seq:
            while len(h) >= 2 and ((h[-1][0] - h[-2][0]) * (p[1] - h[-2][1]) - (h[-1][1] - h[-2][1]) * (p[0] - h[-2][0])) <= 0:
                h.pop()
            h.append(p)
        return h
    lower, upper = half(P), half(P[::-1])
    return np.array(lower[:-1] + upper[:-1])


def clip_polygon_to_rect(poly, x0, y0, x1, y1):
    """Sutherland-Hodgman clip of a convex polygon against an axis-aligned rectangle."""
    pts = [tuple(p) for p in poly]
    for axis, bound, keep_less in ((0, x0, False), (0, x1, True), (1, y0, False), (1, y1, True)):
        if not pts:
            break
        nxt = []
        for a, b in zip(pts, pts[1:] + pts[:1]):
            ina = a[axis] <= bound if keep_less else a[axis] >= bound
            inb = b[axis] <= bound if keep_less else b[axis] >= bound
            if ina != inb:
                t = (bound - a[axis]) / (b[axis] - a[axis])
                cross = (a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1]))
            if ina and inb:
                nxt.append(b)
            elif ina and not inb:
                nxt.append(cross)
            elif not ina and inb:
                nxt.append(cross); nxt.append(b)
        pts = nxt
    return np.array(pts).reshape(-1, 2)


def polygon_area(poly):
    if len(poly) < 3:
        return 0.0
    x, y = poly[:, 0], poly[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


def tile_bounding_box(tile, knn=dist2_knn3_cpu):
    """The 8 homogeneous corners of the tile's 3-D box: xy from the stored box, z from its points after dropping those
    whose nearest-neighbour distance is a 3-sigma outlier (:221-232).  float32 like the reference's CUDA tensors."""
    mx, Mx, my, My = tile["box"]
    xyz = np.array([p.xyz for p in tile["points3D"]]).astype(np.float32)
    dist = knn(xyz).astype(np.float32)
    keep = (dist > (dist.mean() - 3 * dist.std())) & (dist < (dist.mean() + 3 * dist.std()))
    z = xyz[keep][:, -1]
    mz, Mz = float(z.min()), float(z.max())
    return np.array([[x, y, zz, 1.0] for x in (mx, Mx) for y in (my, My) for zz in (mz, Mz)])


def visibility_based_camera_selection(tiles, images, cameras, threshod: float = 0.25, knn=dist2_knn3_cpu) -> List[dict]:
    img_ids, W2C = w2c_matrices(images)
    C = np.linalg.inv(W2C)[:, :3, 3]
    row = {int(i): k for k, i in enumerate(img_ids)}
    K = {cid: _intrinsics(cam) for cid, cam in cameras.items() if cam.model in ("SIMPLE_PINHOLE", "PINHOLE")}
    out = []
    for tile in tiles:
        inside = {im.id for im in tile["images"]}
        bb = tile_bounding_box(tile, knn)                                  # (8,4)
        if not tile["images"]:
            raise ValueError("visibility_based_camera_selection: a tile without cameras (the reference fails here too)")
        own = np.stack([C[row[im.id]] for im in tile["images"]])
        # mean over the tile's cameras of the distance to their farthest box corner, x 1.2 (:238-242)
        d_corner = np.sqrt(((own[None, :, :] - bb[:, None, :3]) ** 2).sum(-1))       # (8, n)
        md = d_corner.max(axis=0).mean() * 1.2
        added = []
        for iid in img_ids:
            iid = int(iid)
            if iid in inside:
                continue
            extr = images[iid]
            cam = cameras[extr.camera_id]
            if extr.camera_id not in K:
                _intrinsics(cam)                                           # raises like the reference
            pc = (W2C[row[iid]] @ bb.T).T
            pc = pc[:, :3] / pc[:, 3:4]
            uv = (K[extr.camera_id] @ pc.T).T
            uv = uv[:, :2] / uv[:, 2:3]
            hull = convex_hull_2d(uv)
            area = polygon_area(clip_polygon_to_rect(hull, 0.0, 0.0, float(cam.width), float(cam.height))) if len(hull) >= 3 else 0.0
            ratio = area / (cam.width * cam.height)
            d = np.mean(np.sum(np.abs(bb[:, :3] - C[row[iid]][None, :]), axis=1))   # L1 per corner, as the reference (:264)
            if ratio > threshod and d < md:
                added.append(extr)
        out.append({"images": added + tile["images"], "box": tile["box"], "points3D": tile["points3D"]})
    return out


# ---- step 4 -------------------------------------------------------------------------------------------------------------
def coverage_based_point_selection(tiles, points3D) -> List[dict]:
    out = []
    for tile in tiles:
        seen = [im.point3D_ids[im.point3D_ids != -1] for im in tile["images"]]
        ids = np.unique(np.concatenate(seen)) if seen else np.zeros((0,), np.int64)
        out.append({"images": tile["images"], "box": tile["box"], "points3D": [points3D[int(i)] for i in ids]})
    return out


# ---- alignment ------------------------------------------------------------------------------------------------------
def transform_colmap(input_model: str, output_model: str, transform_file: str, output_format: str = ".txt"):
    """Rotate a model so that world z is the ground normal: the 4x4 of `transform_file` with translation and per-row
    scale removed is applied to every pose and point (:15-77)."""
    with open(transform_file, "r") as f:
        P = np.vstack([np.array([float(v) for v in line.strip().split(" ")]) for line in f.readlines()])
    assert P.shape == (4, 4), "transform matrix should be 4x4"
    P[:3, -1] = 0
    R = P[:3, :3]
    P[:3, :3] = R / np.sqrt(np.sum(R * R, axis=1))[:, np.newaxis]
    cameras, images, points3D = read_model(path=input_model, ext="")
    Pinv = np.linalg.inv(P)
    ids, W2C = w2c_matrices(images)
    new_w2c = W2C @ Pinv
    images_new = {}
    for k, i in enumerate(ids):
        e = images[int(i)]
        images_new[int(i)] = Image(id=e.id, qvec=rotmat2qvec(new_w2c[k, :3, :3]), tvec=new_w2c[k, :3, -1], camera_id=e.camera_id,
                                   name=e.name, xys=e.xys, point3D_ids=e.point3D_ids)
    points3D_new = {}
    for i, p in points3D.items():
        xyz = (P[:3, :3] @ p.xyz[:, np.newaxis] + P[:3, 3:4]).flatten()
        points3D_new[i] = Point3D(id=p.id, xyz=xyz, rgb=p.rgb, error=p.error, image_ids=p.image_ids, point2D_idxs=p.point2D_idxs)
    os.makedirs(output_model, exist_ok=True)
    write_model(cameras, images_new, points3D_new, path=output_model, ext=output_format)
    return cameras, images_new, points3D_new


# ---- split_scene.py -------------------------------------------------------------------------------------------------
def write_box(path, box):
    with open(path, "w") as f:
        f.write("mx Mx my My\n")
        f.write(f"{box[0]} {box[1]} {box[2]} {box[3]}")


def read_box(path):
    with open(path, "r") as f:
        lines = f.read().split("\n")
    return np.array([float(v) for v in lines[1].split()])


def partition_scene(cameras, images, points3D, num_col=None, num_row=None, max_num_images=200, extend_ratio=0.1,
                    visibility_threshold=0.5, knn=dist2_knn3_cpu) -> List[dict]:
    """The four steps in the order and with the defaults of SceneSpliter.main (split_scene.py:14-53)."""
    tiles = camera_position_based_region_division(images, num_col, num_row, max_num_images)
    tiles = position_based_data_selection(tiles, images, points3D, ratio=extend_ratio)
    tiles = visibility_based_camera_selection(tiles, images, cameras, threshod=visibility_threshold, knn=knn)
    return coverage_based_point_selection(tiles, points3D)


def write_tiles(tiles, cameras, output_path, source_path=None, copy_images=True) -> List[str]:
    """tile_%04d/sparse/0/{cameras,images,points3D}.txt + box.txt (+ images/ copied from <source_path>/images), as
    split_scene.py:55-82 lays them out for train_split.py."""
    dirs = []
    for i, tile in enumerate(tiles):
        name = "tile_%04d" % i
        sparse = os.path.join(output_path, name, "sparse", "0")
        os.makedirs(sparse, exist_ok=True)
        write_model(cameras, {im.id: im for im in tile["images"]}, {p.id: p for p in tile["points3D"]}, path=sparse, ext=".txt")
        write_box(os.path.join(output_path, name, "box.txt"), tile["box"])
        if copy_images:
            dst = os.path.join(output_path, name, "images")
            if os.path.exists(dst):
                shutil.rmtree(dst)
            os.makedirs(dst, exist_ok=True)
            for im in tile["images"]:
                shutil.copy(os.path.join(source_path, "images", im.name), os.path.join(dst, im.name))
        dirs.append(os.path.join(output_path, name))
    return dirs


def split_scene(source_path, output_path=None, num_col=None, num_row=None, max_num_images=200, extend_ratio=0.1,
                visibility_threshold=0.5, transform_file=None, copy_images=True) -> List[str]:
    """split_scene.py's SceneSpliter.main: read <source>/sparse/0 (optionally re-aligned), partition, write the tiles."""
    output_path = output_path or source_path
    os.makedirs(output_path, exist_ok=True)
    if transform_file is not None:
        cameras, images, points3D = transform_colmap(os.path.join(source_path, "sparse/0"), os.path.join(output_path, "sparse/aligned"),
                                                     transform_file)
    else:
        cameras, images, points3D = read_model(path=os.path.join(source_path, "sparse/0"))
    tiles = partition_scene(cameras, images, points3D, num_col, num_row, max_num_images, extend_ratio, visibility_threshold)
    return write_tiles(tiles, cameras, output_path, source_path, copy_images)


def list_tiles(source_path) -> List[str]:
    """Tile directories in index order.  (train_split.py:15-16 pairs an UNSORTED os.listdir with index-named configs,
    SURVEY quirk Q9; sorting is what it means.)"""
    return [os.path.join(source_path, t) for t in sorted(os.listdir(source_path)) if t.startswith("tile_")]


def tiles_for_rank(tile_dirs, rank, world_size) -> List[str]:
    """Tile i -> GPU i mod world_size (north_star: one VastGaussian tile per GPU; train_split.py trains them in a loop)."""
    return [t for i, t in enumerate(tile_dirs) if i % world_size == rank]
