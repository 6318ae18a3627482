"""CPU tests of the host-side logic of the Python mirrors (no kernel launches)."""
import math

import numpy as np
import pytest
import torch


def test_tsdf_view_grouping_keeps_order_and_budget():
    from gsr_b200.tsdf import TSDFFusion
    f = TSDFFusion.__new__(TSDFFusion)
    f.nviews, f.views_per_launch = 32, None
    f._map_bytes = [1600 * 1060 * 4] * 32
    g1, g4 = f._launch_groups(1), f._launch_groups(4)
    for groups, planes in ((g1, 1), (g4, 4)):
        assert groups[0][0] == 0 and sum(n for _, n in groups) == 32
        assert all(groups[i][0] + groups[i][1] == groups[i + 1][0] for i in range(len(groups) - 1))      # consecutive, ordered
        assert all(n * 1600 * 1060 * 4 * planes <= TSDFFusion.MAP_BUDGET_BYTES[planes] for _, n in groups)
    assert max(n for _, n in g1) == 17 and max(n for _, n in g4) == 8      # 6.47 MiB per 1600x1060 plane
    f.views_per_launch = 5
    assert [n for _, n in f._launch_groups(1)] == [5, 5, 5, 5, 5, 5, 2]
    f.nviews, f._map_bytes = 0, []
    assert f._launch_groups(1) == [(0, 0)]
    f.nviews, f._map_bytes, f.views_per_launch = 2, [10 ** 9, 10 ** 9], None         # maps larger than the budget: one per launch
    assert f._launch_groups(1) == [(0, 1), (1, 1)]


def test_ssim_window_equals_the_reference_formula():
    from gsr_b200 import ssim as fused
    from oracle import ssim_oracle
    w1 = torch.tensor(list(fused._WINDOW))
    g = torch.Tensor([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    assert torch.equal(w1, g / g.sum())
    w2 = ssim_oracle.window(1)[0, 0]
    assert torch.allclose(torch.outer(w1, w1), w2, rtol=0, atol=1e-9)
    assert abs(float(w1.sum()) - 1.0) < 1e-6


def test_post_camera_constants_match_depths_to_points():
    """K / rays_o / R handed to the post-processing kernels reproduce depths_to_points (point_utils.py:9-22) exactly."""
    import synth
    from gsr_b200.surfel_post import _camera_constants
    from oracle import surfel_post_oracle as po
    W, H = 97, 33
    q = np.array([0.9, 0.2, -0.3, 0.1]); q /= np.linalg.norm(q)
    cam = synth.make_camera(W, H, R=synth.quat_to_rot(q), t=np.array([0.3, -0.2, 0.5]))
    wvt, fp = torch.from_numpy(cam.viewmatrix), torch.from_numpy(cam.projmatrix)
    c = _camera_constants(wvt, fp, W, H)
    K, o, R = c[:9].reshape(3, 3), c[9:12], c[12:].reshape(3, 3)
    depth = torch.rand(1, H, W) + 1.0
    pts = po.depths_to_points(wvt, fp, W, H, depth).reshape(H, W, 3)
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    rays = torch.stack([xs, ys, torch.ones_like(xs)], -1) @ K
    mine = depth[0][..., None] * rays + o
    assert torch.allclose(mine, pts, rtol=1e-5, atol=1e-5)
    assert torch.equal(R, wvt[:3, :3])


def test_python_mirrors_reject_cpu_tensors_without_touching_the_gpu():
    from gsr_b200.ssim import ssim
    from gsr_b200.surfel_post import surfel_postprocess
    with pytest.raises(RuntimeError, match="no CPU path"):
        ssim(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        surfel_postprocess(torch.zeros(11, 8, 8), torch.eye(4), torch.eye(4))
    with pytest.raises(NotImplementedError):
        ssim(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8), window_size=7)


def test_per_gaussian_shape_checks():
    """A tensor left behind by a densify / prune desync must raise on the host, not read out of bounds on the device."""
    import torch
    from gsr_b200._torch_util import check_per_gaussian
    P = 5
    ok = dict(opacities=(torch.zeros(P, 1), [(1,), ()]), rotations=(torch.zeros(P, 4), [(4,)]), sh=(torch.zeros(P, 16, 3), [(None, 3)]),
              colors_precomp=(torch.zeros(0), [(3,)]), scales=(None, [(2,)]))
    check_per_gaussian(P, **ok)
    check_per_gaussian(P, opacities=(torch.zeros(P), [(1,), ()]))
    with pytest.raises(RuntimeError, match="rotations has 4 rows, expected 5"):
        check_per_gaussian(P, rotations=(torch.zeros(4, 4), [(4,)]))
    with pytest.raises(RuntimeError, match="sh has shape"):
        check_per_gaussian(P, sh=(torch.zeros(P, 16, 4), [(None, 3)]))
    with pytest.raises(RuntimeError, match="scales has shape"):
        check_per_gaussian(P, scales=(torch.zeros(P, 3), [(2,)]))
