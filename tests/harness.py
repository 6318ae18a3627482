"""Shared test/bench plumbing: run one synthetic scene through
  (a) the product path  -- gs-sr_b200/diff_surfel_rasterization (libgsr_b200.so, GPU),
  (b) the CPU oracle    -- oracle/liborc.so,
  (c) the reference CUDA -- oracle/_ref/libref_surfel.so (GPU),
and compare.  Only tests/, bench.py and __graft_entry__.smoke() import this.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gs-sr_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GRAD_KEYS = ("means3D", "means2D", "shs", "colors", "opacities", "scales", "rotations", "transMat")


def to_torch(scene, device="cuda"):
    import torch
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
    cam = scene.cam
    return dict(means3D=t(scene.means3D), scales=t(scene.scales), rotations=t(scene.rotations),
                opacities=t(scene.opacities), colors=t(scene.colors), shs=t(scene.shs),
                bg=t(cam.bg), view=t(cam.viewmatrix), proj=t(cam.projmatrix), campos=t(cam.campos))


def _audit(out_tensor, out):
    """Decision margins of the forward that produced `out_tensor` (gsr_b200.audit); must run before the backward."""
    from gsr_b200.audit import decision_margins
    margins, info, mism = decision_margins(out_tensor)
    out["margins"], out["audit_info"], out["audit_mismatches"] = margins.cpu().numpy(), info.cpu().numpy(), mism


def run_product_surfel(scene, g_color=None, g_others=None, device="cuda", transMat_precomp=None,
                       scale_modifier=1.0, tt=None, audit=False, strided_scales=False):
    """Forward (+ backward when upstream grads are given) through the drop-in Python API.
    strided_scales: hand `scales` over as the stride-3 view scaling[:, :2] GS-SR uses (scaffold_2dgs_scene.py:17)."""
    import torch
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    cam = scene.cam
    tt = tt or to_torch(scene, device)
    leaves = {}
    for k in ("means3D", "scales", "rotations", "opacities", "colors", "shs"):
        if tt[k] is not None:
            leaves[k] = tt[k].clone().requires_grad_(True)
    if transMat_precomp is not None:
        leaves["transMat"] = torch.from_numpy(transMat_precomp).to(device).requires_grad_(True)
        leaves.pop("scales", None); leaves.pop("rotations", None)
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    rs = GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=tt["bg"],
        scale_modifier=scale_modifier, viewmatrix=tt["view"], projmatrix=tt["proj"], sh_degree=scene.sh_degree,
        campos=tt["campos"], prefiltered=False, debug=False)
    rast = GaussianRasterizer(rs)
    scales_arg = leaves.get("scales")
    if strided_scales and scales_arg is not None:
        s3 = torch.cat([tt["scales"], torch.ones_like(tt["scales"][:, :1])], dim=1).requires_grad_(True)
        leaves["scales"] = s3
        scales_arg = s3[:, :2]
    color, radii, others = rast(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                                shs=leaves.get("shs"), colors_precomp=leaves.get("colors"),
                                scales=scales_arg, rotations=leaves.get("rotations"),
                                cov3D_precomp=leaves.get("transMat"))
    out = dict(color=color.detach().cpu().numpy(), others=others.detach().cpu().numpy(),
               radii=radii.cpu().numpy())
    if audit:
        _audit(color, out)
    if g_color is not None:
        gc = torch.from_numpy(g_color).to(device)
        go = torch.from_numpy(g_others).to(device)
        torch.autograd.backward([color, others], [gc, go])
        grads = {"means2D": means2D.grad}
        for k, v in leaves.items():
            grads[k] = v.grad
        out["grads"] = {k: (None if v is None else v.detach().cpu().numpy()) for k, v in grads.items()}
    return out


def run_oracle_surfel(scene, g_color=None, g_others=None, double=False, transMat_precomp=None,
                      scale_modifier=1.0, tile_stride=1):
    from oracle.oracle import SurfelOracle
    o = SurfelOracle(double=double)
    pre = transMat_precomp is not None
    f = o.forward(scene.cam, scene.means3D, scene.opacities, None if pre else scene.scales,
                  None if pre else scene.rotations, colors=scene.colors, shs=scene.shs,
                  sh_degree=scene.sh_degree, transMat_precomp=transMat_precomp,
                  scale_modifier=scale_modifier, tile_stride=tile_stride)
    out = dict(color=f["color"], others=f["others"], radii=f["radii"], num_rendered=f["num_rendered"],
               k_eval=f["k_eval"], oracle=o)
    if g_color is not None:
        out["grads"] = o.backward(g_color, g_others, tile_stride=tile_stride)
    return out


def run_refcuda_surfel(scene, g_color=None, g_others=None, device="cuda", transMat_precomp=None,
                       scale_modifier=1.0, tt=None):
    import torch
    from oracle.refcuda import RefSurfel
    cam = scene.cam
    tt = tt or to_torch(scene, device)
    r = RefSurfel()
    pre = None if transMat_precomp is None else torch.from_numpy(transMat_precomp).to(device)
    color, radii, others, R = r.forward(
        tt["bg"], tt["view"], tt["proj"], tt["campos"], cam.W, cam.H, cam.tanfovx, cam.tanfovy, tt["means3D"],
        tt["opacities"], None if pre is not None else tt["scales"], None if pre is not None else tt["rotations"],
        colors=tt["colors"], shs=tt["shs"], sh_degree=scene.sh_degree, transMat_precomp=pre,
        scale_modifier=scale_modifier)
    torch.cuda.synchronize()
    out = dict(color=color.cpu().numpy(), others=others.cpu().numpy(), radii=radii.cpu().numpy(),
               num_rendered=R, ref=r)
    if g_color is not None:
        g = r.backward(torch.from_numpy(g_color).to(device), torch.from_numpy(g_others).to(device))
        torch.cuda.synchronize()
        out["grads"] = {k: v.cpu().numpy() for k, v in g.items()}
    return out


# ---------------------------------------------------------------------------------------------
# comparison metrics
# ---------------------------------------------------------------------------------------------
def rel_linf(a, b, outlier_frac=0.0):
    """max|a-b| / max|b|, optionally after dropping the `outlier_frac` largest deviations.

    Two correct float32 rasterizers can disagree on whether a splat with alpha within one ulp of
    1/255 (or T within one ulp of 1e-4 / 0.5) is blended; each such flip moves ONE pixel by up to
    ~4e-3.  The reference's own CUDA build shows the same flips against a double evaluation, so
    image comparisons report the raw value and the value with a 1e-5 fraction of pixels excluded."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    if a.size == 0:
        return 0.0
    d = np.abs(a - b)
    scale = max(np.abs(b).max(), 1e-30)
    if outlier_frac > 0:
        k = int(np.ceil(outlier_frac * d.size))
        if 0 < k < d.size:
            d = np.partition(d, d.size - k - 1)[: d.size - k]
    return float(d.max() / scale)


def compare_forward(a, b, names=("a", "b"), outlier_frac=1e-5, verbose=True):
    """Returns dict of rel-Linf per output (raw, robust)."""
    res = {}
    res["color"] = (rel_linf(a["color"], b["color"]), rel_linf(a["color"], b["color"], outlier_frac))
    for ch, nm in ((0, "depth"), (1, "alpha"), (2, "normal_x"), (3, "normal_y"), (4, "normal_z"),
                   (5, "median_depth"), (6, "distortion")):
        res[nm] = (rel_linf(a["others"][ch], b["others"][ch]), rel_linf(a["others"][ch], b["others"][ch], outlier_frac))
    res["radii_mismatch"] = int((a["radii"] != b["radii"]).sum())
    idx_mis = float((a["others"][7] != b["others"][7]).mean())
    res["surf_idx_mismatch_frac"] = idx_mis
    if verbose:
        print(f"forward {names[0]} vs {names[1]}:")
        for k, v in res.items():
            print(f"   {k:24s} {v}")
    return res


def compare_grads(a, b, names=("a", "b"), verbose=True):
    """rel-Linf and rel-L2 per gradient tensor (relative to the max / norm of b)."""
    res = {}
    for k in GRAD_KEYS:
        if k not in a or k not in b or a[k] is None or b[k] is None or np.size(b[k]) == 0:
            continue
        x, y = np.asarray(a[k], np.float64), np.asarray(b[k], np.float64)
        if x.shape != y.shape:
            y = y.reshape(x.shape)
        linf = float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-30))
        l2 = float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-30))
        res[k] = (linf, l2)
    if verbose:
        print(f"grads {names[0]} vs {names[1]}:  (rel Linf, rel L2)")
        for k, v in res.items():
            print(f"   {k:12s} {v[0]:.3e} {v[1]:.3e}")
    return res


# ---------------------------------------------------------------------------------------------
# 3DGS (diff_gaussian_rasterization) and PGSR plane (diff_plane_rasterization) variants
# ---------------------------------------------------------------------------------------------
GAUSS_GRAD_KEYS = ("means3D", "means2D", "means2D_abs", "shs", "colors", "opacities", "scales", "rotations",
                   "cov3D", "all_map")


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def run_product_gauss(scene, g_color=None, plane=False, all_map=None, g_all_map=None, g_plane_depth=None,
                      cov3D_precomp=None, scale_modifier=1.0, render_geo=True, device="cuda", tt=None, audit=False):
    """Forward (+ backward) through the drop-in diff_gaussian_rasterization / diff_plane_rasterization API."""
    import torch
    if plane:
        from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    else:
        from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    cam = scene.cam
    tt = tt or to_torch(scene, device)
    leaves = {}
    for k in ("means3D", "scales", "rotations", "opacities", "colors", "shs"):
        if tt[k] is not None:
            leaves[k] = tt[k].clone().requires_grad_(True)
    if cov3D_precomp is not None:
        leaves["cov3D"] = torch.from_numpy(cov3D_precomp).to(device).requires_grad_(True)
        leaves.pop("scales", None); leaves.pop("rotations", None)
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    kw = dict(image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=tt["bg"],
              scale_modifier=scale_modifier, viewmatrix=tt["view"], projmatrix=tt["proj"],
              sh_degree=scene.sh_degree, campos=tt["campos"], prefiltered=False, debug=False)
    if plane:
        kw["render_geo"] = render_geo
    rast = GaussianRasterizer(GaussianRasterizationSettings(**kw))
    common = dict(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"], shs=leaves.get("shs"),
                  colors_precomp=leaves.get("colors"), scales=leaves.get("scales"),
                  rotations=leaves.get("rotations"), cov3D_precomp=leaves.get("cov3D"))
    if plane:
        means2D_abs = torch.zeros_like(leaves["means3D"], requires_grad=True)
        if all_map is not None:
            leaves["all_map"] = torch.from_numpy(all_map).to(device).requires_grad_(True)
        color, radii, observe, out_all_map, plane_depth = rast(means2D_abs=means2D_abs,
                                                               all_map=leaves.get("all_map"), **common)
        out = dict(color=_np(color), radii=_np(radii), observe=_np(observe), out_all_map=_np(out_all_map),
                   plane_depth=_np(plane_depth))
    else:
        color, radii = rast(**common)
        out = dict(color=_np(color), radii=_np(radii))
    if audit:
        _audit(color, out)
    if g_color is not None:
        outs, gs = [color], [torch.from_numpy(g_color).to(device)]
        if plane and render_geo and g_all_map is not None:
            outs += [out_all_map, plane_depth]
            gs += [torch.from_numpy(g_all_map).to(device), torch.from_numpy(g_plane_depth).to(device)]
        torch.autograd.backward(outs, gs)
        grads = {"means2D": means2D.grad}
        if plane:
            grads["means2D_abs"] = means2D_abs.grad
        for k, v in leaves.items():
            grads[k] = v.grad
        out["grads"] = {k: _np(v) for k, v in grads.items()}
    return out


def run_oracle_gauss(scene, g_color=None, plane=False, all_map=None, g_all_map=None, g_plane_depth=None,
                     cov3D_precomp=None, scale_modifier=1.0, render_geo=True, double=False):
    from oracle.oracle import GaussOracle
    o = GaussOracle(plane=plane, double=double)
    pre = cov3D_precomp is not None
    out = o.forward(scene.cam, scene.means3D, scene.opacities, None if pre else scene.scales,
                    None if pre else scene.rotations, colors=scene.colors, shs=scene.shs,
                    sh_degree=scene.sh_degree, cov3D_precomp=cov3D_precomp, all_map=all_map,
                    scale_modifier=scale_modifier, render_geo=render_geo)
    out["oracle"] = o
    if g_color is not None:
        out["grads"] = o.backward(g_color, g_all_map, g_plane_depth)
    return out


def run_refcuda_gauss(scene, g_color=None, plane=False, all_map=None, g_all_map=None, g_plane_depth=None,
                      cov3D_precomp=None, scale_modifier=1.0, render_geo=True, device="cuda", tt=None):
    import torch
    from oracle.refcuda import RefGauss
    cam = scene.cam
    tt = tt or to_torch(scene, device)
    r = RefGauss(plane=plane)
    pre = None if cov3D_precomp is None else torch.from_numpy(cov3D_precomp).to(device)
    am = None if all_map is None else torch.from_numpy(all_map).to(device)
    f = r.forward(tt["bg"], tt["view"], tt["proj"], tt["campos"], cam.W, cam.H, cam.tanfovx, cam.tanfovy,
                  tt["means3D"], tt["opacities"], None if pre is not None else tt["scales"],
                  None if pre is not None else tt["rotations"], colors=tt["colors"], shs=tt["shs"],
                  sh_degree=scene.sh_degree, cov3D_precomp=pre, all_map=am, scale_modifier=scale_modifier,
                  render_geo=render_geo)
    torch.cuda.synchronize()
    out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in f.items()}
    out["ref"] = r
    if g_color is not None:
        gam = None if g_all_map is None else torch.from_numpy(g_all_map).to(device)
        gpd = None if g_plane_depth is None else torch.from_numpy(g_plane_depth).to(device)
        g = r.backward(torch.from_numpy(g_color).to(device), gam, gpd)
        torch.cuda.synchronize()
        out["grads"] = {k: v.cpu().numpy() for k, v in g.items()}
    return out


def compare_grads_keys(a, b, keys, names=("a", "b"), verbose=True, outlier_frac=0.0):
    """rel-Linf / rel-L2 per gradient tensor; with outlier_frac the rows (Gaussians) with the largest
    deviation are excluded first (same rule as tests/test_oracle_golden.py)."""
    res = {}
    for k in keys:
        if k not in a or k not in b or a[k] is None or b[k] is None or np.size(b[k]) == 0:
            continue
        x, y = np.asarray(a[k], np.float64), np.asarray(b[k], np.float64)
        y = y.reshape(x.shape)
        d = np.abs(x - y).reshape(x.shape[0], -1).max(axis=1)
        keep = np.ones(x.shape[0], bool)
        if outlier_frac > 0:
            kdrop = int(np.ceil(outlier_frac * x.shape[0]))
            if 0 < kdrop < x.shape[0]:
                keep[np.argsort(d)[-kdrop:]] = False
        linf = float(d[keep].max() / max(np.abs(y).max(), 1e-30))
        l2 = float(np.linalg.norm((x - y)[keep]) / max(np.linalg.norm(y), 1e-30))
        res[k] = (linf, l2)
    if verbose:
        print(f"grads {names[0]} vs {names[1]}:  (rel Linf, rel L2)")
        for k, v in res.items():
            print(f"   {k:12s} {v[0]:.3e} {v[1]:.3e}")
    return res
