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
