"""GPU parity at the HEADLINE sizes (pytest -m gpu): product (drop-in API -> C ABI -> sm_100a kernels) against the
UNMODIFIED reference CUDA kernels (oracle/_ref, compiled from /root/reference by oracle/build_ref.sh) on identical inputs:

  * cfg-B  surfel, 2M x 1600x1060, precomputed colours -- contiguous AND stride-3 `scales` (bench.py's workload)
  * 3DGS   1M x 1600x900
  * cfg-4  plane (PGSR), 1M x 1600x900, render_geo, two views

north_star tolerances, asserted RAW: forward <= 1e-4 rel L-inf (relative to the channel's max), gradients <= 1e-3 rel
L-inf (relative to the tensor's max) and rel L2.  No blanket outlier trim: a pixel / Gaussian may exceed the tolerance only
when the cause is demonstrated independently of the deviation itself --

  pixel     one of ITS blend decisions (alpha vs 1/255, T(1-alpha) vs 1e-4, depth vs 0.2, ray-splat vs low-pass branch;
            T vs 0.5 for the median-selected channels) came within BAND of its threshold in the product's own arithmetic
            (gsr_b200.audit.decision_margins re-walks the forward's record stream, bit-checked against the forward),
            i.e. the two float32 implementations took different sides of a hard threshold ("flip");
  Gaussian  it overlaps such a flipped pixel, or it is seen edge-on (|cos(normal, view ray)| < EDGE_COS, surfels only): the
            ray-splat intersection divides by that cosine twice, float32 gradients of such splats are ill-conditioned in
            ANY implementation (the reference deviates from a float64 evaluation by the same amount).

The number of excused pixels / Gaussians is itself asserted (<= 1e-4 of the population) and printed.
"""
import numpy as np
import pytest

import harness as hz
import synth

pytestmark = pytest.mark.gpu

FWD_TOL, GRAD_TOL = 1e-4, 1e-3
DIST_TOL = 2e-3            # distortion channel: difference of O(1) blended moments, channel max ~1e-3 (measured 4.2e-4 at cfg-B)
BAND = 2e-4                # a decision within this relative distance of its threshold may flip between implementations (measured <= 5.8e-5)
EDGE_COS = 1e-2
MAX_EXCUSED = 1e-4


def _need_ref(variant):
    from oracle import refcuda
    if not refcuda.available(variant):
        pytest.skip(f"oracle/_ref/libref_{variant}.so not present")


def _rel_dev(a, b):
    """per-element |a-b| / max|b| as float32 (full-size arrays: avoid float64 temporaries)"""
    scale = max(float(np.abs(b).max()), 1e-30)
    return np.abs(a - b) / np.float32(scale)


def _screen_xy(sc):
    m = sc.means3D.astype(np.float64)
    ph = np.concatenate([m, np.ones((m.shape[0], 1))], 1) @ sc.cam.projmatrix.astype(np.float64)
    w = ph[:, 3:4] + 1e-7
    ndc = ph[:, :2] / w
    return ((ndc[:, 0] + 1.0) * sc.cam.W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * sc.cam.H - 1.0) * 0.5


def _touching(sc, radii, flipped_pix):
    """Gaussians whose screen rectangle (centre +- radius, one pixel of slack) contains a flipped pixel."""
    touch = np.zeros(sc.P, bool)
    if flipped_pix.size == 0:
        return touch
    x, y = _screen_xy(sc)
    r = radii.astype(np.float64) + 1.5
    for p in flipped_pix:
        px, py = p % sc.cam.W, p // sc.cam.W
        touch |= (np.abs(x - px) <= r) & (np.abs(y - py) <= r) & (radii > 0)
    return touch


def _edge_on(sc):
    """|cos| between the surfel normal and the view ray through its centre."""
    q = sc.rotations.astype(np.float64)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    n = np.stack([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], 1)       # R[:, 2]
    v = sc.means3D.astype(np.float64) - sc.cam.campos.astype(np.float64)[None]
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return np.abs((n * v).sum(1))


def check_forward(out, ref, chans, sel_chans=(), idx_chan=None, dist_chan=None, label=""):
    """chans: [(name, ours, ref)] blended channels held to FWD_TOL raw outside flips.  Returns the flipped pixel ids."""
    assert out["audit_mismatches"] == 0, "the audit did not reproduce the forward"
    mg = out["margins"].reshape(out["margins"].shape[0], -1)
    blend_margin = np.minimum(mg[0], mg[1])
    if mg.shape[0] == 5:
        blend_margin = np.minimum(blend_margin, np.minimum(mg[3], mg[4]))
    N = blend_margin.size
    dev = np.zeros(N, np.float32)
    for name, a, b in chans:
        dev = np.maximum(dev, _rel_dev(a, b).ravel())
    flipped = dev > FWD_TOL
    nf = int(flipped.sum())
    print(f"[{label}] pixels over {FWD_TOL:g}: {nf} of {N} ({nf / N:.2e}); their decision margins: "
          f"{np.sort(blend_margin[flipped])[-5:] if nf else []}; max dev outside: {dev[~flipped].max():.2e}; "
          f"marginal pixels (margin < {BAND:g}): {(blend_margin < BAND).mean():.2e}")
    assert nf <= MAX_EXCUSED * N, (label, nf)
    assert (blend_margin[flipped] < BAND).all(), (label, "a deviating pixel has no marginal decision", blend_margin[flipped].max())
    excused = flipped.copy()
    if idx_chan is not None:       # selection channels: a T > 0.5 flip swaps whole values
        sel_margin = np.minimum(blend_margin, mg[2])
        mis = (idx_chan[0] != idx_chan[1]).ravel()
        print(f"[{label}] median-index mismatches: {int(mis.sum())} ({mis.mean():.2e})")
        assert mis.sum() <= MAX_EXCUSED * N and (sel_margin[mis] < BAND).all(), (label, sel_margin[mis].max())
        excused |= mis
        for name, a, b in sel_chans:
            d = _rel_dev(a, b).ravel()
            assert d[~excused].max() <= FWD_TOL, (label, name, d[~excused].max())
    if dist_chan is not None:
        d = _rel_dev(*dist_chan).ravel()
        print(f"[{label}] distortion channel: max rel dev {d[~excused].max():.2e} of a channel max {np.abs(dist_chan[1]).max():.2e}")
        assert d[~excused].max() <= DIST_TOL, (label, d[~excused].max())
        assert np.abs(dist_chan[0] - dist_chan[1]).ravel()[~excused].max() <= FWD_TOL * 0.1   # vs the O(1) moments it is formed from
    return np.where(excused)[0]


def check_grads(sc, out, ref, keys, flipped_pix, edge_cos=None, label=""):
    touch = _touching(sc, out["radii"], flipped_pix)
    explained = touch | (edge_cos < EDGE_COS if edge_cos is not None else False)
    P = sc.P
    for k in keys:
        a = out["grads"][k]; b = np.asarray(ref["grads"][k]).reshape(a.shape)
        d = _rel_dev(a, b).reshape(P, -1).max(1)
        bad = d > GRAD_TOL
        ok = ~(bad & explained)
        l2 = float(np.linalg.norm((a - b)[ok].astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))
        print(f"[{label}] grad {k:11s}: over {GRAD_TOL:g}: {int(bad.sum())} ({bad.mean():.1e}; touching a flipped pixel "
              f"{int((bad & touch).sum())}, edge-on {int((bad & ~touch & explained).sum())}, unexplained {int((bad & ~explained).sum())}); "
              f"max elsewhere {d[ok].max():.2e}; rel L2 {l2:.2e}")
        assert (bad & ~explained).sum() == 0, (label, k, d[bad & ~explained].max())
        assert bad.sum() <= MAX_EXCUSED * P, (label, k, int(bad.sum()))
        assert l2 <= GRAD_TOL, (label, k, l2)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg_b():
    _need_ref("surfel")
    P, W, H = 2_000_000, 1600, 1060
    sc = synth.make_scene(P, W, H, seed=0)
    gc, go = synth.make_upstream_grads(W, H, seed=1)
    tt = hz.to_torch(sc)
    out = hz.run_product_surfel(sc, gc, go, tt=tt, audit=True)
    ref = hz.run_refcuda_surfel(sc, gc, go, tt=tt)
    ref.pop("ref", None)
    return sc, gc, go, tt, out, ref


def _surfel_forward_check(out, ref, label):
    assert (out["radii"] != ref["radii"]).sum() == 0
    O, Ro = out["others"], ref["others"]
    chans = [(f"color{c}", out["color"][c], ref["color"][c]) for c in range(3)] + [(f"others{c}", O[c], Ro[c]) for c in range(5)]
    sel = [("median_depth", O[5], Ro[5])] + [(f"median_normal{c}", O[c], Ro[c]) for c in (8, 9, 10)]
    return check_forward(out, ref, chans, sel, idx_chan=(O[7], Ro[7]), dist_chan=(O[6], Ro[6]), label=label)


def test_cfg_b_surfel_matches_reference_build(cfg_b):
    sc, gc, go, tt, out, ref = cfg_b
    flipped = _surfel_forward_check(out, ref, "cfg-B surfel")
    check_grads(sc, out, ref, ("means3D", "means2D", "colors", "opacities", "scales", "rotations"), flipped, _edge_on(sc), "cfg-B surfel")


def test_cfg_b_strided_scales_view(cfg_b):
    """GS-SR hands scaffold/octree-2DGS `scaling[:, :2]` (stride 3): same images bit for bit, same gradients up to the
    order of the float atomics, and the padding column gets no gradient."""
    sc, gc, go, tt, out, ref = cfg_b
    st = hz.run_product_surfel(sc, gc, go, tt=tt, strided_scales=True)
    assert np.array_equal(st["color"], out["color"]) and np.array_equal(st["others"], out["others"])
    assert np.array_equal(st["radii"], out["radii"])
    g3 = st["grads"]["scales"]
    assert g3.shape == (sc.P, 3) and np.abs(g3[:, 2]).max() == 0
    assert hz.rel_linf(g3[:, :2], out["grads"]["scales"]) <= 2e-5
    for k in ("means3D", "means2D", "colors", "opacities", "rotations"):
        assert hz.rel_linf(st["grads"][k], out["grads"][k]) <= 2e-5, k


@pytest.mark.parametrize("option", ["no_cull", "no_used_bits"])
def test_cfg_b_culling_and_used_bits_do_not_change_results(cfg_b, option):
    """Full size: without the contribution boxes every pair of a tile is evaluated like the reference does; without the
    forward's "blended" marks the backward repeats the cull test.  Images bit-identical, gradients equal up to the order of
    the float atomics."""
    import gsr_b200
    sc, gc, go, tt, out, ref = cfg_b
    L = gsr_b200.lib()
    L.gsr_set_option(option.encode(), 1)
    try:
        alt = hz.run_product_surfel(sc, gc, go, tt=tt)
    finally:
        L.gsr_set_option(option.encode(), 0)
    assert np.array_equal(alt["color"], out["color"]) and np.array_equal(alt["others"], out["others"])
    assert np.array_equal(alt["radii"], out["radii"])
    for k in ("means3D", "means2D", "colors", "opacities", "scales", "rotations"):
        assert hz.rel_linf(alt["grads"][k], out["grads"][k]) <= 2e-5, (option, k)


def test_gaussian_1m_matches_reference_build():
    _need_ref("gaussian")
    P, W, H = 1_000_000, 1600, 900
    sc = synth.make_scene(P, W, H, seed=5, scale_dims=3)
    gc, _ = synth.make_upstream_grads(W, H, seed=6)
    tt = hz.to_torch(sc)
    out = hz.run_product_gauss(sc, gc, tt=tt, audit=True)
    ref = hz.run_refcuda_gauss(sc, gc, tt=tt)
    assert (out["radii"] != ref["radii"]).sum() <= 2          # ceil() of a value within an ulp of an integer
    flipped = check_forward(out, ref, [(f"color{c}", out["color"][c], ref["color"][c]) for c in range(3)], label="3DGS 1M")
    check_grads(sc, out, ref, ("means3D", "means2D", "colors", "opacities", "scales", "rotations"), flipped, None, "3DGS 1M")


@pytest.mark.parametrize("view", [0, 1])
def test_cfg4_plane_matches_reference_build(view):
    """BASELINE config 4 stand-in: PGSR plane rasterizer, 1M x 1600x900, render_geo, reference + neighbour view."""
    _need_ref("plane")
    P, W, H = 1_000_000, 1600, 900
    sc = synth.make_scene(P, W, H, seed=7, scale_dims=3, rotate_camera=bool(view))
    gc, go = synth.make_upstream_grads(W, H, seed=8 + view, n_others=6, zero_from=6)
    kw = dict(g_color=gc, plane=True, all_map=synth.make_all_map(sc), g_all_map=np.ascontiguousarray(go[:5]),
              g_plane_depth=np.ascontiguousarray(go[5:6]))
    tt = hz.to_torch(sc)
    out = hz.run_product_gauss(sc, tt=tt, audit=True, **kw)
    ref = hz.run_refcuda_gauss(sc, tt=tt, **kw)
    assert (out["radii"] != ref["radii"]).sum() <= 2
    chans = [(f"color{c}", out["color"][c], ref["color"][c]) for c in range(3)] + \
            [(f"all_map{c}", out["out_all_map"][c], ref["out_all_map"][c]) for c in range(5)]
    label = f"cfg-4 plane view {view}"
    flipped = check_forward(out, ref, chans, label=label)
    # plane depth = A4 / -(A0 rx + A1 ry + A2): a quotient of blended channels; compare where the denominator is not ~0
    den = np.abs(ref["out_all_map"][2]).ravel()
    pd = _rel_dev(out["plane_depth"], ref["plane_depth"]).ravel()
    okpix = np.ones(pd.size, bool); okpix[flipped] = False
    okpix &= den > 0.05
    print(f"[{label}] plane depth max rel dev {pd[okpix].max():.2e}")
    assert pd[okpix].max() <= FWD_TOL * 10
    # out_observe: integer count of pixels with T > 0.5 per Gaussian; a T-vs-0.5 flip moves one count by one
    dobs = (out["observe"] != ref["observe"])
    mg = out["margins"].reshape(3, -1)
    print(f"[{label}] out_observe mismatches {int(dobs.sum())}; pixels with a marginal T vs 0.5 decision {(mg[2] < BAND).sum()}")
    assert dobs.sum() <= (mg[2] < BAND).sum() + flipped.size and np.abs(out["observe"] - ref["observe"]).max() <= 2
    check_grads(sc, out, ref, ("means3D", "means2D", "means2D_abs", "colors", "opacities", "scales", "rotations", "all_map"),
                flipped, None, label)


@pytest.mark.parametrize("kind", ["surfel", "gaussian"])
def test_more_than_2_pow_23_gaussians_keep_their_indices(kind):
    """P >= 2^23: the record word has no room left for the forward's "blended" marks next to the Gaussian index (bits 23-30),
    so the marks are off and the backward falls back to its cull test.  Ids above 2^23 must come through the records intact:
    compared with the reference build (radii, image, gradients) on 8.4 M small splats, the visible ones spread over the
    whole index range."""
    from oracle import refcuda
    var = "surfel" if kind == "surfel" else "gaussian"
    if not refcuda.available(var):
        pytest.skip(f"oracle/_ref/libref_{var}.so did not travel")
    P, W, H = (1 << 23) + 4097, 400, 304
    sc = synth.make_scene(P, W, H, seed=77, sigma_px=0.8, scale_dims=2 if kind == "surfel" else 3)
    gc, go = synth.make_upstream_grads(W, H, seed=78)
    if kind == "surfel":
        out, ref = hz.run_product_surfel(sc, gc, go), hz.run_refcuda_surfel(sc, gc, go)
    else:
        out, ref = hz.run_product_gauss(sc, gc), hz.run_refcuda_gauss(sc, gc)
    assert np.array_equal(out["radii"], ref["radii"])
    vis = np.nonzero(out["radii"] > 0)[0]
    assert vis.max() >= (1 << 23) and vis.size > P // 2
    assert hz.rel_linf(out["color"], ref["color"], 1e-4) <= 1e-4
    for k in out["grads"]:
        a, b = out["grads"][k], ref["grads"][k]
        assert hz.rel_linf(a, b, 1e-4) <= 1e-3, k
        hi = np.arange(a.shape[0]) >= (1 << 23)            # the rows only a 24-bit index reaches
        assert np.abs(b[hi]).max() > 0 and hz.rel_linf(a[hi], b[hi], 1e-3) <= 1e-3, k
