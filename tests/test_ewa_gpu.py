"""GPU parity tests (pytest -m gpu) for the EWA family (diff_gaussian_rasterization, diff_plane_rasterization),
scaffold_filter and simple_knn: the product path (drop-in Python API -> C ABI -> hand-written sm_100a kernels)
against the golden vectors captured from the unmodified reference CUDA kernels, the CPU oracle on seeded scenes,
the reference CUDA build itself when oracle/_ref travelled, and size-independent identities at full size.
Tolerances: tests/test_ewa_oracle_golden.py."""
import os

import numpy as np
import pytest
import torch

import harness as hz
import synth
from golden.cases import (FILTER_CASES, GAUSS_CASES, KNN_CASES, build_filter_case, build_gauss_case,
                          build_knn_case)
from test_ewa_oracle_golden import check_ewa_forward, check_ewa_grads, ewa_grad_keys, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GAUSS_CASES)
def test_ewa_product_matches_golden_reference_vectors(name):
    gold = load("gauss", name)
    sc, kw = build_gauss_case(name)
    out = hz.run_product_gauss(sc, **kw)
    check_ewa_forward(out, gold, kw)
    check_ewa_grads(out["grads"], gold, ewa_grad_keys(sc, kw))


@pytest.mark.parametrize("plane,P,W,H,sh,seed", [(False, 20000, 320, 240, False, 31), (True, 20000, 333, 177, False, 32),
                                                 (True, 5000, 320, 240, True, 33)])
def test_ewa_product_matches_oracle(plane, P, W, H, sh, seed):
    sc = synth.make_scene(P, W, H, seed=seed, sh=sh, rotate_camera=True, bg=(0.1, 0.2, 0.3), scale_dims=3)
    gc, go = synth.make_upstream_grads(W, H, seed=seed + 1, n_others=6, zero_from=6)
    kw = dict(g_color=gc, plane=plane)
    if plane:
        kw.update(all_map=synth.make_all_map(sc), g_all_map=np.ascontiguousarray(go[:5]),
                  g_plane_depth=np.ascontiguousarray(go[5:6]))
    out = hz.run_product_gauss(sc, **kw)
    orc = hz.run_oracle_gauss(sc, **kw)
    check_ewa_forward(out, orc, kw, radii_slack=max(1, P // 2000))
    check_ewa_grads(out["grads"], orc["grads"], ewa_grad_keys(sc, kw))


@pytest.mark.parametrize("plane", [False, True])
def test_ewa_product_matches_reference_cuda_build(plane):
    from oracle import refcuda
    if not refcuda.available("plane" if plane else "gaussian"):
        pytest.skip("oracle/_ref reference build not present")
    P, W, H = 150000, 800, 600
    sc = synth.make_scene(P, W, H, seed=41, sh=not plane, rotate_camera=True, scale_dims=3)
    gc, go = synth.make_upstream_grads(W, H, seed=42, n_others=6, zero_from=6)
    kw = dict(g_color=gc, plane=plane)
    if plane:
        kw.update(all_map=synth.make_all_map(sc), g_all_map=np.ascontiguousarray(go[:5]),
                  g_plane_depth=np.ascontiguousarray(go[5:6]))
    tt = hz.to_torch(sc)
    out = hz.run_product_gauss(sc, tt=tt, **kw)
    ref = hz.run_refcuda_gauss(sc, tt=tt, **kw)
    check_ewa_forward(out, ref, kw, radii_slack=max(1, P // 2000))
    check_ewa_grads(out["grads"], ref["grads"], ewa_grad_keys(sc, kw))


@pytest.mark.parametrize("plane", [False, True])
def test_ewa_used_bits_path_equals_cull_path(plane):
    """Same contract as the surfel test: the backward that walks the forward's marks (P < 2^23) and the one that repeats
    the cull test (forced with no_used_bits) visit the same pairs."""
    import gsr_b200
    sc = synth.make_scene(200000, 640, 400, seed=45, scale_dims=3)
    gc, go = synth.make_upstream_grads(640, 400, seed=46, n_others=6, zero_from=6)
    kw = dict(g_color=gc, plane=plane)
    if plane:
        kw.update(all_map=synth.make_all_map(sc), g_all_map=np.ascontiguousarray(go[:5]), g_plane_depth=np.ascontiguousarray(go[5:6]))
    tt = hz.to_torch(sc)
    a = hz.run_product_gauss(sc, tt=tt, **kw)
    gsr_b200.lib().gsr_set_option(b"no_used_bits", 1)
    try:
        b = hz.run_product_gauss(sc, tt=tt, **kw)
    finally:
        gsr_b200.lib().gsr_set_option(b"no_used_bits", 0)
    assert np.array_equal(a["color"], b["color"]) and np.array_equal(a["radii"], b["radii"])
    if plane:
        assert np.array_equal(a["observe"], b["observe"]) and np.array_equal(a["out_all_map"], b["out_all_map"])
    for k in a["grads"]:
        if a["grads"][k] is not None and b["grads"][k] is not None and a["grads"][k].size:
            x, y = a["grads"][k].astype(np.float64), b["grads"][k].astype(np.float64)
            assert np.abs(x - y).max() <= 1e-5 * max(np.abs(y).max(), 1e-30), k


@pytest.mark.parametrize("seed", list(range(8)))
def test_ewa_random_configuration_sweep(seed):
    """Seeded sweep (3DGS for even seeds, plane + render_geo for odd): image shapes, splat sizes, opacity ranges, points
    behind the camera; radii / out_observe equal the reference build, colour matches the oracle."""
    from oracle import refcuda
    rng = np.random.default_rng(5000 + seed)
    plane = bool(seed & 1)
    W = int(rng.choice([16, 23, 64, 97, 130, 256])); H = int(rng.choice([16, 31, 48, 75, 128]))
    P = int(rng.choice([1, 7, 200, 1500, 4000]))
    sc = synth.make_scene(P, W, H, seed=6000 + seed, scale_dims=3, sigma_px=float(rng.choice([0.3, 1.0, 3.0, 9.0, 25.0])),
                          opacity_sigma=float(rng.choice([0.5, 1.5, 4.0])), rotate_camera=bool(rng.integers(0, 2)),
                          behind_fraction=float(rng.choice([0.0, 0.3])), bg=tuple(rng.uniform(0, 1, 3)))
    gc, go = synth.make_upstream_grads(W, H, seed=7000 + seed, n_others=6, zero_from=6)
    kw = dict(g_color=gc, plane=plane)
    if plane:
        kw.update(all_map=synth.make_all_map(sc), g_all_map=np.ascontiguousarray(go[:5]), g_plane_depth=np.ascontiguousarray(go[5:6]))
    out = hz.run_product_gauss(sc, **kw)
    orc = hz.run_oracle_gauss(sc, **kw)
    assert hz.rel_linf(out["color"], orc["color"], 2e-3) <= 1e-4
    assert ((out["radii"] > 0) != (orc["radii"] > 0)).sum() <= max(1, P // 500)
    for k in ("opacities", "means3D", "colors"):
        assert np.isfinite(out["grads"][k]).all()
    if refcuda.available("plane" if plane else "gaussian"):
        ref = hz.run_refcuda_gauss(sc, **kw)
        assert (out["radii"] != ref["radii"]).sum() <= max(1, P // 2000)
        if plane:
            assert (out["observe"] != ref["observe"]).sum() <= max(1, P // 500)
        if (ref["radii"] > 0).sum() >= 50:
            g = {k: v for k, v in ref["grads"].items() if k in ("opacities", "means3D", "colors", "scales")}
            res = hz.compare_grads_keys(out["grads"], g, list(g.keys()), verbose=False, outlier_frac=2e-3)
            for k, (linf, l2) in res.items():
                assert linf <= 1e-3 and l2 <= 1e-3, (k, linf, l2)


def test_ewa_culling_never_changes_results():
    """no_cull evaluates every (pixel, splat) pair of the reference's tile lists; the culled path must give
    bit-identical images and last contributors (identical blend order)."""
    import gsr_b200
    sc = synth.make_scene(30000, 320, 240, seed=51, rotate_camera=True, scale_dims=3)
    gc, _ = synth.make_upstream_grads(320, 240, seed=52)
    kw = dict(g_color=gc, plane=True, all_map=synth.make_all_map(sc), render_geo=True,
              g_all_map=np.zeros((5, 240, 320), np.float32), g_plane_depth=np.zeros((1, 240, 320), np.float32))
    a = hz.run_product_gauss(sc, **kw)
    assert gsr_b200.lib().gsr_set_option(b"no_cull", 1) == 0
    try:
        b = hz.run_product_gauss(sc, **kw)
    finally:
        gsr_b200.lib().gsr_set_option(b"no_cull", 0)
    for k in ("color", "out_all_map", "plane_depth", "radii", "observe"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("opacities", "colors"):
        assert np.allclose(a["grads"][k], b["grads"][k], rtol=1e-4, atol=1e-9), k


@pytest.mark.parametrize("plane", [False, True])
@pytest.mark.parametrize("cap", [500, 30_000_000])
def test_ewa_binning_capacity_prediction_paths(plane, cap):
    """No stream sync for num_rendered (include/gsr_b200.h): a forced capacity far below R takes the clamped run + exact
    re-run (out_observe, accumulated with atomics, must start from zero again), one far above R lets the speculative run
    stand; results are bit-identical to the default path."""
    import gsr_b200
    sc = synth.make_scene(40000, 320, 240, seed=61, rotate_camera=True, scale_dims=3)
    gc, _ = synth.make_upstream_grads(320, 240, seed=62)
    kw = dict(g_color=gc, plane=plane)
    if plane:
        kw.update(all_map=synth.make_all_map(sc), render_geo=True, g_all_map=np.zeros((5, 240, 320), np.float32),
                  g_plane_depth=np.zeros((1, 240, 320), np.float32))
    base = hz.run_product_gauss(sc, **kw)
    L = gsr_b200.lib()
    L.gsr_set_option(b"force_capacity", cap)
    try:
        forced = hz.run_product_gauss(sc, **kw)
    finally:
        L.gsr_set_option(b"force_capacity", 0)
    for k in ("color", "radii") + (("out_all_map", "plane_depth", "observe") if plane else ()):
        assert np.array_equal(base[k], forced[k]), k
    for k in ("opacities", "colors", "means3D"):
        assert hz.rel_linf(forced["grads"][k], base["grads"][k]) <= 2e-5, k


def test_ewa_edge_cases_empty_culled_and_strided():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    sc = synth.make_scene(64, 48, 32, seed=61, scale_dims=3, bg=(0.2, 0.4, 0.6))
    tt = hz.to_torch(sc)
    rs = GaussianRasterizationSettings(sc.cam.H, sc.cam.W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"],
                                       tt["proj"], 0, tt["campos"], False, False)
    rast = GaussianRasterizer(rs)
    # P == 0: background-free zero image like the reference (nothing is launched)
    e = torch.zeros((0, 3), device="cuda")
    color, radii = rast(means3D=e, means2D=e, opacities=torch.zeros((0, 1), device="cuda"), colors_precomp=e,
                        scales=e, rotations=torch.zeros((0, 4), device="cuda"))
    assert color.shape == (3, 32, 48) and float(color.abs().max()) == 0 and radii.numel() == 0
    # everything behind the camera: pure background, zero grads
    m = tt["means3D"].clone(); m[:, 2] = -1.0
    m.requires_grad_(True)
    color, radii = rast(means3D=m, means2D=torch.zeros_like(m), opacities=tt["opacities"], colors_precomp=tt["colors"],
                        scales=tt["scales"], rotations=tt["rotations"])
    assert int((radii > 0).sum()) == 0
    assert torch.allclose(color, tt["bg"][:, None, None].expand_as(color))
    color.sum().backward()
    assert float(m.grad.abs().max()) == 0
    # strided scales view (scaling[:, :3] of a (P,6) tensor, scaffold_scene.py) == contiguous copy
    wide = torch.cat([tt["scales"], torch.rand_like(tt["scales"])], dim=1)
    c1, r1 = rast(means3D=tt["means3D"], means2D=torch.zeros_like(tt["means3D"]), opacities=tt["opacities"],
                  colors_precomp=tt["colors"], scales=wide[:, :3], rotations=tt["rotations"])
    c2, r2 = rast(means3D=tt["means3D"], means2D=torch.zeros_like(tt["means3D"]), opacities=tt["opacities"],
                  colors_precomp=tt["colors"], scales=tt["scales"], rotations=tt["rotations"])
    assert torch.equal(c1, c2) and torch.equal(r1, r2)
    # prefiltered violation raises instead of trapping the context
    rs2 = rs._replace(prefiltered=True)
    with pytest.raises(RuntimeError, match="prefiltered"):
        GaussianRasterizer(rs2)(means3D=m.detach(), means2D=torch.zeros_like(m), opacities=tt["opacities"],
                                colors_precomp=tt["colors"], scales=tt["scales"], rotations=tt["rotations"])


def test_ewa_full_size_identities():
    """1 M planar Gaussians at 1600x900 (SURVEY 8(d) config 4 stand-in): linearity of the blend in colours and
    background, all_map[3] (the blended 1.0) == 1 - T, finite outputs, |grad| sums dominate signed sums."""
    from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    P, W, H = 1_000_000, 1600, 900
    sc = synth.make_scene(P, W, H, seed=71, scale_dims=3)
    tt = hz.to_torch(sc)
    am = torch.from_numpy(synth.make_all_map(sc)).cuda()
    bg0 = torch.zeros(3, device="cuda")

    def render(colors, bg):
        rs = GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, bg, 1.0, tt["view"], tt["proj"], 0,
                                           tt["campos"], False, True, False)
        m2 = torch.zeros_like(tt["means3D"], requires_grad=True)
        m2a = torch.zeros_like(tt["means3D"], requires_grad=True)
        out = GaussianRasterizer(rs)(means3D=tt["means3D"], means2D=m2, means2D_abs=m2a, opacities=tt["opacities"],
                                     colors_precomp=colors, scales=tt["scales"], rotations=tt["rotations"], all_map=am)
        return out, m2, m2a

    (c1, radii, obs, amap, pd), m2, m2a = render(tt["colors"], bg0)
    (c2, *_), _, _ = render(2.0 * tt["colors"], bg0)
    (c3, *_), _, _ = render(tt["colors"], torch.ones(3, device="cuda"))
    assert torch.isfinite(c1).all() and torch.isfinite(amap).all()
    assert torch.allclose(c2, 2.0 * c1, rtol=1e-5, atol=1e-6)
    alpha = amap[3].detach()
    assert float(alpha.min()) >= 0 and float(alpha.max()) <= 1 + 1e-5
    assert torch.allclose(c3, c1 + (1 - alpha)[None], rtol=1e-5, atol=2e-6)
    assert int(obs.sum()) > 0 and int(obs.min()) >= 0 and int(obs[radii == 0].sum()) == 0
    (c1.sum() + amap.sum()).backward()
    assert (m2.grad.abs() <= m2a.grad * (1 + 1e-4) + 1e-9).all()


@pytest.mark.parametrize("name", FILTER_CASES)
def test_filter_product_matches_golden(name):
    from scaffold_filter import GaussianRasterizationSettings, GaussianRasterizer
    sc, kw = build_filter_case(name)
    tt = hz.to_torch(sc)
    rs = GaussianRasterizationSettings(sc.cam.H, sc.cam.W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"],
                                       kw.get("scale_modifier", 1.0), tt["view"], tt["proj"], 0, tt["campos"], False, False)
    wide = torch.cat([tt["scales"], torch.rand_like(tt["scales"])], dim=1)   # (P,6)[:, :3] as the callers pass it
    r = GaussianRasterizer(rs).visible_filter(tt["means3D"], wide[:, :3], tt["rotations"])
    assert r.dtype == torch.int32
    assert np.array_equal(r.cpu().numpy(), load("filter", name)["radii"])
    assert GaussianRasterizer(rs).visible_filter(torch.zeros((0, 3), device="cuda")).numel() == 0


def test_filter_product_equals_rasterizer_radii_at_full_size():
    """2 M anchors at 1600x1060: visible_filter radii == radii of a full 3DGS forward."""
    from scaffold_filter import GaussianRasterizationSettings, GaussianRasterizer
    import diff_gaussian_rasterization as dgr
    sc = synth.make_scene(2_000_000, 1600, 1060, seed=81, scale_dims=3)
    tt = hz.to_torch(sc)
    args = (sc.cam.H, sc.cam.W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"], tt["proj"], 0, tt["campos"],
            False, False)
    r = GaussianRasterizer(GaussianRasterizationSettings(*args)).visible_filter(tt["means3D"], tt["scales"], tt["rotations"])
    _, radii = dgr.GaussianRasterizer(dgr.GaussianRasterizationSettings(*args))(
        means3D=tt["means3D"], means2D=torch.zeros_like(tt["means3D"]), opacities=tt["opacities"],
        colors_precomp=tt["colors"], scales=tt["scales"], rotations=tt["rotations"])
    assert torch.equal(r, radii)
    from oracle import refcuda
    if os.path.exists(os.path.join(hz.ROOT, "oracle", "_ref", "libref_filter.so")):
        rr = refcuda.ref_visible_filter(tt["means3D"], tt["scales"], tt["rotations"], tt["view"], tt["proj"], sc.cam.W,
                                        sc.cam.H, sc.cam.tanfovx, sc.cam.tanfovy)
        assert int((r != rr).sum()) <= 20 and int(((r > 0) != (rr > 0)).sum()) <= 2


@pytest.mark.parametrize("name", KNN_CASES)
def test_knn_product_matches_golden_bit_exactly(name):
    from simple_knn._C import distCUDA2
    pts = build_knn_case(name)
    d = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    assert np.array_equal(d.view(np.uint32), load("knn", name)["dist2"].view(np.uint32))


@pytest.mark.parametrize("P,clustered", [(1, False), (2, False), (3, False), (4, False), (33, False), (1025, True),
                                         (40000, True), (50000, False)])
def test_knn_product_matches_oracle_bit_exactly(P, clustered):
    from oracle import oracle as orc
    from simple_knn._C import distCUDA2
    pts = synth.make_points(P, seed=90 + P % 7, clustered=clustered)
    d = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    want = orc.dist2_knn3(pts)
    assert np.array_equal(d.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("case", ["far_outlier", "two_outliers_and_duplicates", "slab"])
def test_knn_with_outliers_matches_oracle_bit_exactly(case):
    """SfM clouds have far outliers: the Morton grid is laid over a robust box (mean +- 4 sigma of trimmed moments, knn.cu),
    points outside fall into the border cells.  The order is only a heuristic -- the result must stay the exact 3-NN."""
    from oracle import oracle as orc
    from simple_knn._C import distCUDA2
    rng = np.random.default_rng(17)
    pts = rng.random((30000, 3), dtype=np.float32)
    if case == "far_outlier":
        pts[123] = 1e4
    elif case == "two_outliers_and_duplicates":
        pts[5] = (-3e3, 2e3, 10.0); pts[6] = (4e3, -1e3, -2e3); pts[100:200] = pts[300]
    else:
        pts[:, 2] *= 1e-4; pts[7, 2] = 50.0
    d = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    want = orc.dist2_knn3(pts)
    assert np.array_equal(d.view(np.uint32), want.view(np.uint32))


def test_knn_full_size_properties():
    """2 M points: invariance under permutation of the input order (a size-independent property of an exact
    k-NN), bit-equality with the reference CUDA build when it travelled, degenerate axis."""
    from simple_knn._C import distCUDA2
    pts = torch.from_numpy(synth.make_points(2_000_000, seed=99)).cuda()
    d = distCUDA2(pts)
    perm = torch.randperm(pts.shape[0], device="cuda")
    d2 = distCUDA2(pts[perm])
    assert torch.equal(d[perm], d2)
    assert float(d.min()) >= 0 and torch.isfinite(d).all()
    from oracle import refcuda
    if os.path.exists(os.path.join(hz.ROOT, "oracle", "_ref", "libref_knn.so")):
        assert torch.equal(refcuda.ref_dist2_knn3(pts).view(torch.int32), d.view(torch.int32))
    flat = pts[:100000].clone(); flat[:, 2] = 1.5    # zero extent along z (the reference divides by zero here)
    df = distCUDA2(flat)
    assert torch.isfinite(df).all()
    sub = flat[:3000].cpu().numpy()
    from oracle import oracle as orc
    assert np.array_equal(distCUDA2(flat[:3000]).cpu().numpy().view(np.uint32), orc.dist2_knn3(sub).view(np.uint32))
