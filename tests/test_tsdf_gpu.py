"""GPU: gsr_tsdf_fuse (through gsr_b200.tsdf.TSDFFusion -> C ABI) against the golden vectors of the reference's
own fusion code, against the numpy oracle on fresh seeds, and size-independent identities at a full
256^3 chunk (the unit GS-SR's marching cubes evaluates, gssr/utils/mcube_utils.py:60)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tsdf_synth import TSDF_CASES, build_tsdf_case  # noqa: E402
from test_tsdf_oracle_golden import TOL_GPU, check_against_golden  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_product(c, return_rgb=True, samples=None):
    from gsr_b200.tsdf import TSDFFusion
    f = TSDFFusion([torch.from_numpy(m) for m in c["projs"]], [torch.from_numpy(d) for d in c["depthmaps"]],
                   [torch.from_numpy(r) for r in c["rgbmaps"]], center=c["center"], radius=c["radius"])
    s = torch.from_numpy(c["samples"] if samples is None else samples).cuda()
    out = f.compute_unbounded_tsdf(s, True if c["contracted"] else None, c["voxel_size"], return_rgb=return_rgb)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", TSDF_CASES)
def test_product_matches_reference_golden(name):
    c = build_tsdf_case(name)
    gold = np.load(os.path.join(GOLD, f"tsdf_{name}.npz"))
    tsdf, rgb = run_product(c)
    check_against_golden(tsdf.cpu().numpy(), rgb.cpu().numpy(), gold, tol=TOL_GPU)
    only = run_product(c, return_rgb=False)
    assert torch.equal(only, tsdf)                       # colour fusion never changes the tsdf


def test_product_matches_oracle_fresh_seed():
    from oracle import tsdf_oracle
    c = build_tsdf_case("ragged")
    rng = np.random.default_rng(77)
    c["samples"] = rng.uniform(-1.9, 1.9, size=(50000, 3)).astype(np.float32)
    tsdf, rgb = run_product(c)
    ot, orgb = tsdf_oracle.compute_unbounded_tsdf(c["samples"], c["contracted"], c["center"], c["radius"], c["voxel_size"],
                                                  c["projs"], c["depthmaps"], c["rgbmaps"], return_rgb=True)
    check_against_golden(tsdf.cpu().numpy(), rgb.cpu().numpy(), dict(tsdf=ot, rgb=orgb), tol=TOL_GPU)


def test_empty_and_errors():
    from gsr_b200.tsdf import TSDFFusion
    c = build_tsdf_case("contracted", n=16)
    f = TSDFFusion([torch.from_numpy(m) for m in c["projs"]], [torch.from_numpy(d) for d in c["depthmaps"]], None,
                   center=c["center"], radius=c["radius"])
    assert f.compute_unbounded_tsdf(torch.empty((0, 3), device="cuda"), True, 0.01).shape == (0,)
    with pytest.raises(RuntimeError):
        f.compute_unbounded_tsdf(torch.zeros((4, 2), device="cuda"), True, 0.01)
    with pytest.raises(RuntimeError):
        f.compute_unbounded_tsdf(torch.zeros((4, 3), device="cuda"), True, 0.01, return_rgb=True)
    with pytest.raises(RuntimeError):
        f.compute_unbounded_tsdf(torch.zeros((4, 3)), True, 0.01)
    g = TSDFFusion([], [], None)
    assert torch.all(g.compute_unbounded_tsdf(torch.rand((100, 3), device="cuda"), None, 0.01) == 1)


def test_full_chunk_identities():
    """256^3 lattice chunk x 32 views at 1600x1060: (1) tsdf in [-1, 1]; (2) a permutation of the samples permutes
    the result (no cross-sample state); (3) fusing the views in two streamed halves through the C ABI's
    init=0 continuation gives bit-identical results to one call."""
    import ctypes
    from gsr_b200 import check, lib
    from gsr_b200.tsdf import TSDFFusion
    c = build_tsdf_case("bench")
    Ng = 256
    ax = torch.linspace(-1.2, 1.2, Ng, device="cuda")
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
    pts = torch.stack([xx.ravel(), yy.ravel(), zz.ravel()], -1).contiguous()
    f = TSDFFusion([torch.from_numpy(m) for m in c["projs"]], [torch.from_numpy(d) for d in c["depthmaps"]],
                   [torch.from_numpy(r) for r in c["rgbmaps"]], center=c["center"], radius=c["radius"])
    t = f.compute_unbounded_tsdf(pts, True, c["voxel_size"])
    assert torch.isfinite(t).all() and t.min() >= -1 and t.max() <= 1
    assert (t < 1).float().mean() > 0.001
    perm = torch.randperm(pts.shape[0], device="cuda")
    t2 = f.compute_unbounded_tsdf(pts[perm].contiguous(), True, c["voxel_size"])
    assert torch.equal(t2, t[perm])
    # streamed halves
    L = lib()
    n = pts.shape[0]
    ts = torch.empty(n, device="cuda"); ws = torch.empty(n, device="cuda")
    half = f.nviews // 2
    s = torch.cuda.current_stream().cuda_stream
    check(L.gsr_tsdf_fuse(n, pts.data_ptr(), 1, f._center, f.radius, float(c["voxel_size"]), half, f._views.data_ptr(), 1,
                          ts.data_ptr(), ws.data_ptr(), None, s), "gsr_tsdf_fuse")
    check(L.gsr_tsdf_fuse(n, pts.data_ptr(), 1, f._center, f.radius, float(c["voxel_size"]), f.nviews - half,
                          f._views.data_ptr() + 96 * half, 0, ts.data_ptr(), ws.data_ptr(), None, s), "gsr_tsdf_fuse")
    assert torch.equal(ts, t)


# ---- bounded volume (extract_mesh_bounded / extract_mesh_split) ------------------------------------------------------------
GRID = dict(origin=(-0.8, -0.8, -0.8), voxel_size=0.05, dims=(33, 30, 29), sdf_trunc=0.2, depth_trunc=4.0)


def _views(c, sl=slice(None)):
    return ([torch.from_numpy(m) for m in c["projs"][sl]], [torch.from_numpy(d) for d in c["depthmaps"][sl]],
            [torch.from_numpy(r) for r in c["rgbmaps"][sl]])


@pytest.mark.parametrize("with_rgb", [False, True])
def test_bounded_volume_matches_oracle(with_rgb):
    """gsr_tsdf_integrate_grid against the numpy restatement of the same rule (the reference's torch rule on a bounded
    lattice + the depth <= 0 / > depth_trunc masks of mesh_utils.py:160-170; Open3D itself is absent: parity unpinned)."""
    from gsr_b200.tsdf import BoundedTSDFVolume
    from oracle import tsdf_oracle
    c = build_tsdf_case("ragged")
    vol = BoundedTSDFVolume(with_rgb=with_rgb, **GRID)
    p, d, r = _views(c)
    vol.integrate(p, d, r if with_rgb else None)
    ot, ow, orgb = tsdf_oracle.integrate_grid(projs=c["projs"], depthmaps=c["depthmaps"], rgbmaps=c["rgbmaps"] if with_rgb else None, **GRID)
    assert vol.tsdf.shape == (29, 30, 33)
    t, w = vol.tsdf.cpu().numpy(), vol.weight.cpu().numpy()
    flips = (w != ow)                                  # a sample within rounding of the truncation / frustum / depth mask
    assert flips.mean() <= 1e-3
    assert np.abs(t - ot)[~flips].max() <= 1e-4
    if with_rgb:
        assert np.abs(vol.rgb.cpu().numpy() - orgb)[~flips].max() <= 1e-4
    assert (w > 1).mean() > 0.05 and (t < 0).any() and (t > 0.5).any()          # the surface is inside the volume


def test_bounded_volume_streams_views_and_combines_like_two_gpus():
    """Views integrated in several calls == one call; two volumes fused from disjoint view sets (one VastGaussian tile per
    GPU) combined with the reduce algebra == one volume fused from all views."""
    from gsr_b200.tsdf import BoundedTSDFVolume, combine_partial_volumes
    c = build_tsdf_case("world_rgb", n=8)
    p, d, r = _views(c)
    full = BoundedTSDFVolume(with_rgb=True, **GRID).integrate(p, d, r)
    streamed = BoundedTSDFVolume(with_rgb=True, **GRID)
    for sl in (slice(0, 2), slice(2, 5)):
        streamed.integrate(*_views(c, sl))
    assert torch.equal(streamed.tsdf, full.tsdf) and torch.equal(streamed.weight, full.weight) and torch.equal(streamed.rgb, full.rgb)
    a = BoundedTSDFVolume(with_rgb=True, **GRID).integrate(*_views(c, slice(0, 3)))
    b = BoundedTSDFVolume(with_rgb=True, **GRID).integrate(*_views(c, slice(3, 5)))
    t, w, rgb = combine_partial_volumes([(a.tsdf, a.weight, a.rgb), (b.tsdf, b.weight, b.rgb)])
    assert torch.equal(w, full.weight)
    assert float((t - full.tsdf).abs().max()) <= 2e-6 and float((rgb - full.rgb).abs().max()) <= 2e-6
    empty = BoundedTSDFVolume(with_rgb=False, **GRID).integrate([], [])
    assert float(empty.tsdf.min()) == 1.0 and float(empty.weight.max()) == 1.0
    with pytest.raises(ValueError):
        BoundedTSDFVolume(with_rgb=False, **GRID).integrate(p, d, r)
