"""GPU: NaN / Inf in the companions' inputs and extreme image shapes -- no fault, no hang, and the reference build's answer
where the reference build travelled (tests/gpu_poison_probe2.py is the exploratory version)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]

import harness as hz  # noqa: E402
import synth  # noqa: E402

REF = os.path.join(hz.ROOT, "oracle", "_ref")


@pytest.mark.parametrize("case", ["nan_one", "inf_one", "nan_many", "all_nan", "huge", "all_same"])
def test_dist2_knn3_with_non_finite_points(case):
    from oracle import refcuda
    from simple_knn._C import distCUDA2
    pts = np.random.default_rng(3).random((50000, 3), dtype=np.float32)
    if case == "nan_one": pts[17, 1] = np.nan
    if case == "inf_one": pts[17, 1] = np.inf
    if case == "nan_many": pts[::7, 0] = np.nan
    if case == "all_nan": pts[:] = np.nan
    if case == "huge": pts[5] = 3e38; pts[6] = -3e38
    if case == "all_same": pts[:] = 0.25
    t = torch.from_numpy(pts).cuda()
    d = distCUDA2(t)
    torch.cuda.synchronize()
    assert d.shape == (50000,)
    if case in ("nan_one", "inf_one", "huge", "all_same") and os.path.exists(os.path.join(REF, "libref_knn.so")):
        r = refcuda.ref_dist2_knn3(t)                      # (the reference itself loops / faults on the other cases)
        fo, fr = torch.isfinite(d), torch.isfinite(r)
        assert torch.equal(fo, fr) and torch.equal(d[fo].view(torch.int32), r[fo].view(torch.int32))


@pytest.mark.parametrize("case", ["nan_mean", "nan_scale", "huge_scale", "nan_quat"])
def test_visible_filter_with_non_finite_anchors(case):
    from oracle import refcuda
    from scaffold_filter import GaussianRasterizationSettings, GaussianRasterizer
    W, H, P = 320, 240, 20000
    sc = synth.make_scene(P, W, H, seed=11, scale_dims=3)
    if case == "nan_mean": sc.means3D[::9, 1] = np.nan
    if case == "nan_scale": sc.scales[::9, 1] = np.nan
    if case == "huge_scale": sc.scales[::9] = 1e30
    if case == "nan_quat": sc.rotations[::9, 2] = np.nan
    tt = hz.to_torch(sc)
    rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=sc.cam.tanfovx, tanfovy=sc.cam.tanfovy, bg=tt["bg"],
                                       scale_modifier=1.0, viewmatrix=tt["view"], projmatrix=tt["proj"], sh_degree=0,
                                       campos=tt["campos"], prefiltered=False, debug=False)
    radii = GaussianRasterizer(rs).visible_filter(means3D=tt["means3D"], scales=tt["scales"], rotations=tt["rotations"])
    torch.cuda.synchronize()
    if os.path.exists(os.path.join(REF, "libref_filter.so")):
        want = refcuda.ref_visible_filter(tt["means3D"], tt["scales"], tt["rotations"], tt["view"], tt["proj"], W, H,
                                          sc.cam.tanfovx, sc.cam.tanfovy)
        assert torch.equal(radii, want)


@pytest.mark.parametrize("w,h", [(1, 1), (1, 37), (17, 15), (5, 200), (200, 5), (4097, 3), (3, 4097), (2000, 16)])
def test_extreme_image_shapes_match_the_reference_build(w, h):
    from oracle import refcuda
    if not refcuda.available("surfel"):
        pytest.skip("oracle/_ref/libref_surfel.so did not travel")
    sc = synth.make_scene(3000, w, h, seed=4, sigma_px=2.0)
    gc, go = synth.make_upstream_grads(w, h, seed=5)
    out = hz.run_product_surfel(sc, gc, go)
    ref = hz.run_refcuda_surfel(sc, gc, go)
    assert np.array_equal(out["radii"], ref["radii"])
    assert hz.rel_linf(out["color"], ref["color"]) <= 1e-4
    for k in out["grads"]:
        assert hz.rel_linf(out["grads"][k], ref["grads"][k]) <= 1e-3, k


def test_image_space_companions_with_non_finite_inputs():
    """TSDF fusion with NaN / Inf depth maps and NaN samples, SSIM with a NaN pixel, marching cubes + cluster filter over a
    lattice with NaN / Inf voxels: no fault; what does not depend on the poisoned values stays finite."""
    from gsr_b200.mesh import extract_triangle_mesh, post_process_mesh
    from gsr_b200.ssim import ssim
    from gsr_b200.tsdf import TSDFFusion
    from tsdf_synth import build_tsdf_case
    c = build_tsdf_case("contracted", n=5000)
    dm = [d.copy() for d in c["depthmaps"]]
    dm[0][10:20, 10:20] = np.nan
    dm[1][:] = np.inf
    f = TSDFFusion([torch.from_numpy(m) for m in c["projs"]], [torch.from_numpy(d) for d in dm], None, center=c["center"],
                   radius=c["radius"])
    s = torch.from_numpy(c["samples"]).cuda()
    s[::11] = float("nan")
    t = f.compute_unbounded_tsdf(s, True, c["voxel_size"])
    torch.cuda.synchronize()
    assert t.shape == (5000,) and float(t[torch.isfinite(t)].abs().max()) <= 1.0
    a = torch.rand(3, 100, 130, device="cuda", requires_grad=True)
    b = torch.rand(3, 100, 130, device="cuda")
    b[1, 5, 5] = float("nan")
    ssim(a, b).backward()
    torch.cuda.synchronize()
    bad = ~torch.isfinite(a.grad)
    assert not bad[0].any() and not bad[2].any() and int(bad[1].sum()) <= 21 * 21      # the NaN stays inside its two windows
    g = torch.randn(40, 41, 42, device="cuda")
    g[::5, ::3, ::2] = float("nan")
    g[1::5, ::3, ::2] = float("inf")
    m = extract_triangle_mesh(g)
    p = post_process_mesh(m, cluster_to_keep=5)
    torch.cuda.synchronize()
    V = m.vertices.shape[0]
    assert int(m.triangles.min()) == 0 and int(m.triangles.max()) == V - 1 and 0 < p.triangles.shape[0] <= m.triangles.shape[0]
