"""CPU: every call GS-SR's own Python makes into the five extension modules (tests/golden/reference_callsites.json, read off the
reference sources by tests/golden/make_golden_callsites.py) binds against the drop-in packages: the imported names exist, the
settings are constructible from exactly the keywords the reference passes, the rasterizer / visible_filter / distCUDA2
signatures accept the reference's arguments."""
import importlib
import inspect
import json
import os

import pytest

torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
SITES = json.load(open(os.path.join(HERE, "golden", "reference_callsites.json")))


def _id(c):
    return f"{c['file']}:{c['line']}:{c['kind']}"


def test_fixture_covers_the_callers_survey_names():
    files = {c["file"] for c in SITES["calls"]}
    for f in ("gssr/scene/twodgs_scene.py", "gssr/scene/pgsr_scene.py", "gssr/scene/vanilla_scene.py", "gssr/scene/scaffold_scene.py",
              "gssr/scene/octree_scene.py", "gssr/gaussian/vanilla_gaussian.py", "gssr/gaussian/scaffold_gaussian.py",
              "gssr/gaussian/octree_gaussian.py", "gssr/utils/vastgaussian_utils.py"):
        assert f in files, f
    assert {c["module"] for c in SITES["calls"]} == {"diff_surfel_rasterization", "diff_gaussian_rasterization",
                                                     "diff_plane_rasterization", "scaffold_filter", "simple_knn._C"}


@pytest.mark.parametrize("imp", SITES["imports"], ids=lambda i: f"{i['file']}:{i['line']}:{i['name']}")
def test_imported_names_exist(imp):
    assert hasattr(importlib.import_module(imp["module"]), imp["name"])


@pytest.mark.parametrize("call", SITES["calls"], ids=_id)
def test_call_site_binds(call):
    mod = importlib.import_module(call["module"])
    kw = {k: None for k in call["keywords"]}
    if call["kind"] == "settings":
        fields = mod.GaussianRasterizationSettings._fields
        assert set(call["keywords"]) == set(fields) and call["positional"] == 0
        mod.GaussianRasterizationSettings(**kw)
    elif call["kind"] == "forward":
        inspect.signature(mod.GaussianRasterizer.forward).bind(None, **kw)
    elif call["kind"] == "visible_filter":
        inspect.signature(mod.GaussianRasterizer.visible_filter).bind(None, **kw)
    else:
        assert call["kind"] == "distCUDA2" and call["positional"] == 1 and not call["keywords"]
        inspect.signature(mod.distCUDA2).bind(None)
