"""CPU tests: the EWA / filter / knn oracle (oracle/gauss_oracle.c) against the golden vectors captured
from the UNMODIFIED reference CUDA kernels on a B200 (tests/golden/make_golden.py gauss).

Tolerances (north_star: forward <= 1e-4 rel L-inf, gradients <= 1e-3 rel):
  * images (colour, all_map, plane depth): rel L-inf (relative to the channel's max) <= 1e-4 after excluding
    the 1e-3 fraction of pixels with the largest deviation (alpha-within-an-ulp-of-1/255 flips, see
    tests/test_oracle_golden.py); raw <= 2e-2;
  * radii, num_rendered: exact; out_observe: exact up to T > 0.5 flips (<= 1e-3 of the entries, off by <= 2);
  * gradients: rel L-inf and rel L2 <= 1e-3 after excluding the 2e-3 fraction of Gaussians with the largest
    deviation, raw rel L2 <= 5e-2;
  * visible_filter radii: exact; distCUDA2: BIT-exact (the oracle evaluates the reference build's
    fma(dz,dz, fma(dx,dx, dy*dy)) sequence).
"""
import os

import numpy as np
import pytest

import harness as hz
from golden.cases import (FILTER_CASES, GAUSS_CASES, KNN_CASES, build_filter_case, build_gauss_case,
                          build_knn_case)
from test_oracle_golden import FWD_RAW_TOL, FWD_TOL, GRAD_L2_TOL, GRAD_RAW_L2_TOL, GRAD_TOL, grad_errors

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(prefix, name):
    path = os.path.join(GOLD, f"{prefix}_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {path} missing")
    return np.load(path)


def check_ewa_forward(out, gold, kw, radii_slack=0):
    if "num_rendered" in out and radii_slack == 0:
        assert int(out["num_rendered"]) == int(gold["num_rendered"])
    assert (out["radii"] != gold["radii"]).sum() <= radii_slack
    imgs = [("color", out["color"], gold["color"])]
    if kw["plane"]:
        d = np.abs(out["observe"].astype(np.int64) - gold["observe"].astype(np.int64))
        assert (d != 0).mean() <= 1e-3 and d.max() <= 2, ("observe", (d != 0).sum(), d.max())
        if kw.get("render_geo", True):
            imgs += [(f"all_map[{c}]", out["out_all_map"][c], gold["out_all_map"][c]) for c in range(5)]
            imgs += [("plane_depth", out["plane_depth"], gold["plane_depth"])]
        else:
            assert np.abs(out["out_all_map"]).max() == 0 and np.abs(out["plane_depth"]).max() == 0
    for name, a, b in imgs:
        assert hz.rel_linf(a, b, 1e-3) <= FWD_TOL, (name, hz.rel_linf(a, b, 1e-3))
        if name != "plane_depth":   # a ratio of blended sums: unbounded where the blended normal is ~ perpendicular
            assert hz.rel_linf(a, b) <= FWD_RAW_TOL, (name, hz.rel_linf(a, b))


def ewa_grad_keys(sc, kw):
    keys = ["means2D", "opacities", "means3D"]
    keys += ["shs"] if sc.shs is not None else ["colors"]
    keys += ["cov3D"] if "cov3D_precomp" in kw else ["scales", "rotations"]
    if kw["plane"]:
        keys += ["means2D_abs"]
        if kw.get("render_geo", True):
            keys += ["all_map"]
    return keys


def check_ewa_grads(grads, gold, keys):
    for k in keys:
        g = gold["grad_" + k] if not isinstance(gold, dict) else gold[k]
        if g.size == 0:
            continue
        a = np.asarray(grads[k]).reshape(g.shape)
        linf, l2, raw = grad_errors(a, g)
        assert linf <= GRAD_TOL and l2 <= GRAD_L2_TOL and raw <= GRAD_RAW_L2_TOL, (k, linf, l2, raw)


@pytest.mark.parametrize("name", GAUSS_CASES)
def test_ewa_oracle_matches_reference_cuda(name):
    gold = load("gauss", name)
    sc, kw = build_gauss_case(name)
    out = hz.run_oracle_gauss(sc, **kw)
    check_ewa_forward(out, gold, kw)
    check_ewa_grads(out["grads"], gold, ewa_grad_keys(sc, kw) + ["conic"] + ([] if "cov3D_precomp" in kw else ["cov3D"]))


def test_plane_without_geo_equals_gaussian_colour():
    """L/ with render_geo=False blends exactly what G/ blends (same kernels minus the all_map terms)."""
    sc, kw = build_gauss_case("g_colors")
    a = hz.run_oracle_gauss(sc, **kw)
    kw2 = dict(kw, plane=True, render_geo=False)
    b = hz.run_oracle_gauss(sc, **kw2)
    assert np.array_equal(a["color"], b["color"]) and np.array_equal(a["radii"], b["radii"])
    for k in ("means2D", "means3D", "opacities", "colors", "scales", "rotations"):
        assert np.array_equal(a["grads"][k], b["grads"][k]), k
    # |.| sums dominate the signed sums
    assert (np.abs(b["grads"]["means2D"]) <= b["grads"]["means2D_abs"] * (1 + 1e-5) + 1e-12).all()


def test_ewa_oracle_double_build_agrees():
    sc, kw = build_gauss_case("p_geo")
    a = hz.run_oracle_gauss(sc, **kw)
    b = hz.run_oracle_gauss(sc, double=True, **kw)
    check_ewa_forward(a, b, kw)
    check_ewa_grads(a["grads"], b["grads"], ewa_grad_keys(sc, kw))


def test_ewa_oracle_empty_and_all_culled():
    import synth
    sc = synth.make_scene(50, 48, 32, seed=3, scale_dims=3)
    sc.means3D[:, 2] = -1.0
    gc, _ = synth.make_upstream_grads(48, 32)
    out = hz.run_oracle_gauss(sc, g_color=gc)
    assert out["num_rendered"] == 0 and (out["radii"] == 0).all()
    assert np.allclose(out["color"], sc.cam.bg[:, None, None])
    assert all(np.abs(v).max() == 0 for v in out["grads"].values() if v.size)


@pytest.mark.parametrize("name", FILTER_CASES)
def test_filter_oracle_matches_reference_cuda(name):
    from oracle import oracle as orc
    sc, kw = build_filter_case(name)
    r = orc.visible_filter(sc.cam, sc.means3D, sc.scales, sc.rotations, **kw)
    gold = load("filter", name)["radii"]
    assert np.array_equal(r, gold)
    assert 0 < (gold > 0).sum() and ((gold > 0).sum() < gold.size or name == "f_basic")


def test_filter_oracle_is_the_3dgs_preprocess():
    """visible_filter radii == radii of the full 3DGS forward on the same anchors (F/forward.cu:267-342 is
    G/forward.cu:155-256 truncated)."""
    from oracle import oracle as orc
    sc, kw = build_gauss_case("g_ragged")
    full = hz.run_oracle_gauss(sc, **kw)
    assert np.array_equal(orc.visible_filter(sc.cam, sc.means3D, sc.scales, sc.rotations), full["radii"])


@pytest.mark.parametrize("name", KNN_CASES)
def test_knn_oracle_matches_reference_cuda_bit_exactly(name):
    from oracle import oracle as orc
    pts = build_knn_case(name)
    d = orc.dist2_knn3(pts)
    gold = load("knn", name)["dist2"]
    assert np.array_equal(d.view(np.uint32), gold.view(np.uint32))


def test_knn_oracle_against_numpy_bruteforce():
    from oracle import oracle as orc
    import synth
    pts = synth.make_points(600, seed=9, clustered=True)
    d = orc.dist2_knn3(pts)
    D = ((pts[:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(D, np.inf)
    want = np.sort(D, axis=1)[:, :3].mean(axis=1)
    assert np.allclose(d, want, rtol=1e-5, atol=1e-12)
    assert (d[want == 0] == 0).all()   # triple duplicates give exactly 0
