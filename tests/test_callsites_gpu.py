"""GPU: one tiny call per reference call site (tests/golden/reference_callsites.json) with exactly the keywords the
reference passes, unpacked into as many values as the reference unpacks -- the drop-ins return what GS-SR's scene classes
expect at every place they call the extensions."""
import importlib
import json
import os

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import harness as hz  # noqa: E402
import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SITES = json.load(open(os.path.join(HERE, "golden", "reference_callsites.json")))
CALLS = [c for c in SITES["calls"] if c["kind"] != "settings"]


@pytest.mark.parametrize("call", CALLS, ids=lambda c: f"{c['file']}:{c['line']}:{c['kind']}")
def test_call_site_runs_with_the_reference_keywords(call):
    mod = importlib.import_module(call["module"])
    if call["kind"] == "distCUDA2":
        d = mod.distCUDA2(torch.from_numpy(synth.make_points(500, seed=1)).cuda())
        assert d.shape == (500,) and d.dtype == torch.float32 and bool((d > 0).all())
        return
    P, W, H = 3000, 96, 80
    surfel = call["module"] == "diff_surfel_rasterization"
    sc = synth.make_scene(P, W, H, seed=3, scale_dims=2 if surfel else 3)
    tt = hz.to_torch(sc)
    settings_site = next(s for s in SITES["calls"] if s["kind"] == "settings" and s["file"] == call["file"])
    values = dict(image_height=H, image_width=W, tanfovx=sc.cam.tanfovx, tanfovy=sc.cam.tanfovy, bg=tt["bg"], scale_modifier=1.0,
                  viewmatrix=tt["view"], projmatrix=tt["proj"], sh_degree=0, campos=tt["campos"], prefiltered=False, debug=False,
                  render_geo=True)
    rs = mod.GaussianRasterizationSettings(**{k: values[k] for k in settings_site["keywords"]})
    rast = mod.GaussianRasterizer(raster_settings=rs)
    args = dict(means3D=tt["means3D"], means2D=torch.zeros_like(tt["means3D"], requires_grad=True),
                means2D_abs=torch.zeros_like(tt["means3D"], requires_grad=True), shs=None, colors_precomp=tt["colors"],
                opacities=tt["opacities"], scales=tt["scales"], rotations=tt["rotations"], cov3D_precomp=None,
                all_map=torch.from_numpy(synth.make_all_map(sc)).cuda() if "all_map" in call["keywords"] else None)
    kw = {k: args[k] for k in call["keywords"]}
    if call["kind"] == "visible_filter":
        radii = rast.visible_filter(**kw)
        assert radii.shape == (P,) and int((radii > 0).sum()) > 0
        return
    out = rast(**kw)
    assert isinstance(out, tuple) and len(out) == call["unpacked_into"]
    assert out[0].shape == (3, H, W) and out[1].shape == (P,) and torch.isfinite(out[0]).all()
