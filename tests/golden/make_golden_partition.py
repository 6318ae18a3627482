"""Generates tests/golden/partition_*.json|npz with the REFERENCE's own partition / COLMAP-I/O functions.

Run here (needs /root/reference; no GPU):   python tests/golden/make_golden_partition.py

The reference modules are imported from /root/reference unmodified.  Three of their imports are absent from this image
and are shimmed IN THIS SCRIPT ONLY (they are inputs to the golden data, not product code):
  * shapely (Polygon / box / .intersection(...).area): replaced by an independent exact convex-polygon intersection
    built on fractions-free float64 half-plane clipping written below (not the product's clipper);
  * simple_knn._C.distCUDA2 (a CUDA extension): replaced by a brute-force float32 3-NN in numpy;
  * torch.Tensor.cuda(): identity (the reference moves the points to the GPU only to call distCUDA2).
Outputs: tile membership (image ids, point ids, boxes) after every step for the 2x2 grid split and the quadtree split of
the 64-camera synthetic model, SHA-256 of the files the reference's writers produce for it, and a tiny model written by
the reference in both formats (tests/golden/colmap_tiny/).
"""
import hashlib, json, os, sys, types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")

# ---- shims ------------------------------------------------------------------------------------------------------------
import torch  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self


def _brute_knn(points):
    p = points.detach().cpu().numpy().astype(np.float32)
    out = np.zeros(len(p), np.float32)
    for i in range(len(p)):
        d = p - p[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(np.float32)
        d2[i] = np.inf
        s = np.sort(d2)[:3]
        out[i] = np.float32(np.float32(s[0] + s[1]) + s[2]) / np.float32(3)
    return torch.from_numpy(out)


knn_mod = types.ModuleType("simple_knn"); knn_c = types.ModuleType("simple_knn._C"); knn_c.distCUDA2 = _brute_knn
knn_mod._C = knn_c
sys.modules["simple_knn"], sys.modules["simple_knn._C"] = knn_mod, knn_c


class _Poly:
    def __init__(self, pts):
        self.pts = [tuple(map(float, p)) for p in pts]

    @property
    def area(self):
        a = 0.0
        for (x0, y0), (x1, y1) in zip(self.pts, self.pts[1:] + self.pts[:1]):
            a += x0 * y1 - x1 * y0
        return abs(a) / 2

    def intersection(self, other):
        # clip self against every edge of the (convex, counter-clockwise) other polygon
        out = self.pts
        o = other.pts
        if sum(x0 * y1 - x1 * y0 for (x0, y0), (x1, y1) in zip(o, o[1:] + o[:1])) < 0:
            o = o[::-1]
        for (ax, ay), (bx, by) in zip(o, o[1:] + o[:1]):
            side = lambda p: (bx - ax) * (p[1] - ay) - (by - ay) * (p[0] - ax)  # noqa: E731
            nxt = []
            for p, q in zip(out, out[1:] + out[:1]):
                sp, sq = side(p), side(q)
                if sp >= 0:
                    nxt.append(p)
                if (sp >= 0) != (sq >= 0):
                    t = sp / (sp - sq)
                    nxt.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
            out = nxt
            if not out:
                break
        return _Poly(out)


geom = types.ModuleType("shapely.geometry")
geom.Polygon = _Poly
geom.box = lambda x0, y0, x1, y1: _Poly([(x0, y0), (x1, y0), (x1, y1), (x0, y1)])
sys.modules["shapely"] = types.ModuleType("shapely"); sys.modules["shapely.geometry"] = geom

from gssr.utils import colmap_read_write_model as ref_io  # noqa: E402
from gssr.utils import vastgaussian_utils as ref_vast  # noqa: E402

import partition_synth  # noqa: E402


def to_ref(cameras, images, points):
    cams = {i: ref_io.Camera(id=i, model=m, width=w, height=h, params=p) for i, (m, w, h, p) in cameras.items()}
    imgs = {i: ref_io.Image(id=i, qvec=q, tvec=t, camera_id=c, name=n, xys=xy, point3D_ids=pid)
            for i, (q, t, c, n, xy, pid) in images.items()}
    pts = {i: ref_io.Point3D(id=i, xyz=x, rgb=rgb, error=e, image_ids=ii, point2D_idxs=jj) for i, (x, rgb, e, ii, jj) in points.items()}
    return cams, imgs, pts


def snapshot(tiles):
    ident = lambda im: int(im["image"].id) if isinstance(im, dict) else int(im.id)   # step 1 keeps {"image", "center"} dicts  # noqa: E731
    return [{"images": [ident(im) for im in t["images"]], "points3D": [int(p.id) for p in t.get("points3D", [])],
             "box": [float(v) for v in t["box"]]} for t in tiles]


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def main():
    import tempfile
    cams, imgs, pts = to_ref(*partition_synth.make_model())
    gold = {}
    for tag, kw in (("grid2x2", dict(num_col=2, num_row=2)), ("quadtree", dict(num_col=None, num_row=None, max_num_images=20))):
        t1 = ref_vast.camera_position_based_region_division(imgs, **kw)
        t2 = ref_vast.position_based_data_selection(t1, imgs, pts, ratio=0.1)
        t3 = ref_vast.visibility_based_camera_selection(t2, imgs, cams, threshod=0.5)
        t4 = ref_vast.coverage_based_point_selection(t3, pts)
        gold[tag] = {"step1": snapshot(t1), "step2": snapshot(t2), "step3": snapshot(t3), "step4": snapshot(t4)}
        print(tag, "tiles:", len(t4), "images/tile:", [len(t["images"]) for t in t4], "added by visibility:",
              [len(a["images"]) - len(b["images"]) for a, b in zip(t3, t2)])
    # step 3 on its own: the L1 distance rule (:264) keeps cameras OUTSIDE a grid tile out almost always, so the hull /
    # clip / area-ratio logic is exercised on tiles from which every second camera has been removed (those cameras sit
    # above the tile: the area ratio decides)
    t1 = ref_vast.camera_position_based_region_division(imgs, num_col=2, num_row=2)
    t2 = ref_vast.position_based_data_selection(t1, imgs, pts, ratio=0.8)
    thin = [{"images": t["images"][::2], "box": t["box"], "points3D": t["points3D"]} for t in t2]
    gold["thinned"] = {"ratio": 0.8, "input": snapshot(thin)}
    for thr in (0.25, 0.4):
        t3 = ref_vast.visibility_based_camera_selection(thin, imgs, cams, threshod=thr)
        gold["thinned"][f"thr{thr}"] = snapshot(t3)
        print("thinned thr", thr, "added", [len(a["images"]) - len(b["images"]) for a, b in zip(t3, thin)])
    # the reference's writers on the full model: hashes only (files are MBs)
    with tempfile.TemporaryDirectory() as d:
        for ext in (".txt", ".bin"):
            ref_io.write_model(cams, imgs, pts, d, ext=ext)
        gold["sha256"] = {f: sha(os.path.join(d, f)) for f in sorted(os.listdir(d))}
    with open(os.path.join(HERE, "partition_64cam.json"), "w") as f:
        json.dump(gold, f)
    # a tiny model written by the reference in both formats, kept as files
    tiny = os.path.join(HERE, "colmap_tiny")
    os.makedirs(tiny, exist_ok=True)
    tc, ti, tp = to_ref(*partition_synth.make_model(n_side=2, n_points=12, seed=3))
    for ext in (".txt", ".bin"):
        ref_io.write_model(tc, ti, tp, tiny, ext=ext)
    # rotmat2qvec / qvec2rotmat of the reference on fixed inputs
    rng = np.random.default_rng(5)
    q = rng.normal(size=(16, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    R = np.stack([ref_io.qvec2rotmat(v) for v in q])
    np.savez(os.path.join(HERE, "partition_quat.npz"), q=q, R=R, q_back=np.stack([ref_io.rotmat2qvec(m) for m in R]))
    print("written", os.listdir(tiny))


if __name__ == "__main__":
    main()
