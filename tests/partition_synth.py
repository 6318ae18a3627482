"""Synthetic COLMAP model for BASELINE config 1: a grid of nadir-looking cameras over a mostly
planar point cloud, with 2-D observations and tracks filled in consistently.  Plain Python/numpy structures (tuples of
arrays); tests/test_partition_cpu.py and tests/golden/make_golden_partition.py turn them into the product's / the
reference's record types."""
import numpy as np


def make_model(n_side=8, n_points=3000, seed=0, width=640, height=480, focal=420.0, spacing=4.0, height_above=14.0):
    """-> cameras {id: (model, W, H, params)}, images {id: (qvec, tvec, cam_id, name, xys, point3D_ids)},
    points {id: (xyz, rgb, error, image_ids, point2D_idxs)}; ids start at 1, image order is shuffled."""
    rng = np.random.default_rng(seed)
    cameras = {1: ("PINHOLE", width, height, np.array([focal, focal * 1.01, width / 2.0, height / 2.0]))}
    extent = spacing * (n_side - 1)
    pts = np.stack([rng.uniform(-3, extent + 3, n_points), rng.uniform(-3, extent + 3, n_points),
                    rng.normal(0.0, 0.15, n_points)], axis=1)
    far = rng.choice(n_points, 6, replace=False)
    pts[far, 2] += rng.uniform(15, 30, 6)                      # isolated high points: dropped by the 3-sigma kNN filter
    order = rng.permutation(n_side * n_side)
    images, obs_of_point = {}, {i + 1: ([], []) for i in range(n_points)}
    for rank, k in enumerate(order):
        gx, gy = k % n_side, k // n_side
        c = np.array([gx * spacing + rng.normal(0, 0.2), gy * spacing + rng.normal(0, 0.2), height_above + rng.normal(0, 0.3)])
        # camera looks down: x right, y "down" = -world y, z = -world z, plus a small tilt
        a, b = rng.normal(0, 0.04, 2)
        Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        R = Rx @ Ry @ np.diag([1.0, -1.0, -1.0])
        t = -R @ c
        pc = pts @ R.T + t
        u = focal * pc[:, 0] / pc[:, 2] + width / 2.0
        v = focal * 1.01 * pc[:, 1] / pc[:, 2] + height / 2.0
        vis = np.where((pc[:, 2] > 0.1) & (u >= 0) & (u < width) & (v >= 0) & (v < height))[0]
        vis = vis[rng.random(len(vis)) < 0.7]
        xys = np.stack([u[vis], v[vis]], axis=1)
        pids = (vis + 1).astype(np.int64)
        # a few observations without a 3-D point
        extra = rng.integers(0, 4)
        xys = np.concatenate([xys, rng.uniform(0, 100, (extra, 2))])
        pids = np.concatenate([pids, -np.ones(extra, np.int64)])
        iid = rank + 1
        for j, pid in enumerate(pids):
            if pid > 0:
                obs_of_point[int(pid)][0].append(iid); obs_of_point[int(pid)][1].append(j)
        images[iid] = (rot_to_qvec(R), t, 1, f"img_{gy:02d}_{gx:02d}.jpg", xys, pids)
    points = {}
    for pid in range(1, n_points + 1):
        points[pid] = (pts[pid - 1], rng.integers(0, 256, 3), float(rng.uniform(0.1, 2.0)),
                       np.array(obs_of_point[pid][0], np.int64), np.array(obs_of_point[pid][1], np.int64))
    return cameras, images, points


def rot_to_qvec(R):
    """Rotation matrix -> (w, x, y, z), Shepperd's method (independent of the code under test)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q if q[0] >= 0 else -q
