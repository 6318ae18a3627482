#!/usr/bin/env python3
"""Writes tests/golden/reference_callsites.json: every place GS-SR's own Python (gssr/scene, gssr/gaussian, gssr/utils,
the entry scripts) imports from or calls into the five extension modules this repo replaces, read off the reference
sources with `ast` -- which names are imported, which keyword arguments every GaussianRasterizationSettings(...) /
rasterizer(...) / rasterizer.visible_filter(...) / distCUDA2(...) call passes, and how many values the call's result is
unpacked into.  tests/test_callsites_*.py bind these calls against the drop-in packages (CPU: signatures; GPU: a tiny call per
site with exactly these keywords and this unpacking), so "GS-SR's scene classes run unchanged on the drop-ins" is checked
call site by call site although the reference itself cannot travel to the GPU box.

    python tests/golden/make_golden_callsites.py        (needs /root/reference)
"""
import ast
import json
import os

REF = "/root/reference"
MODULES = ("diff_gaussian_rasterization", "diff_surfel_rasterization", "diff_plane_rasterization", "scaffold_filter", "simple_knn")
HERE = os.path.dirname(os.path.abspath(__file__))


def files():
    for root in (os.path.join(REF, "gssr"),):
        for d, _, names in os.walk(root):
            for n in sorted(names):
                if n.endswith(".py"):
                    yield os.path.join(d, n)
    for n in sorted(os.listdir(REF)):
        if n.endswith(".py"):
            yield os.path.join(REF, n)


def call_name(node):
    f = node.func
    if isinstance(f, ast.Name):
        return f.id
    if isinstance(f, ast.Attribute):
        return f.attr
    return None


def module_name(path):
    rel = os.path.relpath(path, REF)[:-3].replace(os.sep, ".")
    return rel[:-9] if rel.endswith(".__init__") else rel


def main():
    out = {"imports": [], "calls": []}
    trees = {path: ast.parse(open(path).read()) for path in files()}
    # local name -> (extension module, original name) per file; names re-exported through gssr modules
    # (octree_scene takes GaussianRasterizer from gssr.scene.scaffold_scene) are followed to the extension
    aliases = {path: {} for path in trees}
    by_module = {module_name(path): path for path in trees}
    for path, tree in trees.items():
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] in MODULES:
                for a in node.names:
                    aliases[path][a.asname or a.name] = (node.module, a.name)
                    out["imports"].append({"file": os.path.relpath(path, REF), "line": node.lineno, "module": node.module, "name": a.name})
    for _ in range(4):
        for path, tree in trees.items():
            for node in ast.walk(tree):
                if isinstance(node, ast.ImportFrom) and node.module in by_module:
                    src = aliases[by_module[node.module]]
                    for a in node.names:
                        if a.name in src and (a.asname or a.name) not in aliases[path]:
                            aliases[path][a.asname or a.name] = src[a.name]
                            out["imports"].append({"file": os.path.relpath(path, REF), "line": node.lineno, "module": src[a.name][0],
                                                   "name": a.name, "via": node.module})
    for path, tree in trees.items():
        rel = os.path.relpath(path, REF)
        alias = aliases[path]
        if not alias:
            continue
        # which extension module this file's rasterizer objects come from
        raster_mod = sorted({m for m, n in alias.values() if n == "GaussianRasterizer"})
        parents = {}
        for node in ast.walk(tree):
            for ch in ast.iter_child_nodes(node):
                parents[ch] = node
        for node in ast.walk(tree):
            if not isinstance(node, ast.Call):
                continue
            name = call_name(node)
            kind = None
            if name in alias and alias[name][1] == "GaussianRasterizationSettings":
                kind, module = "settings", alias[name][0]
            elif name in alias and alias[name][1] == "distCUDA2":
                kind, module = "distCUDA2", alias[name][0]
            elif name == "visible_filter":
                kind, module = "visible_filter", (raster_mod[0] if raster_mod else None)
            elif name == "rasterizer" and raster_mod:
                kind, module = "forward", raster_mod[0]
            if kind is None:
                continue
            arity = None
            p = parents.get(node)
            if isinstance(p, ast.Assign) and len(p.targets) == 1 and isinstance(p.targets[0], ast.Tuple):
                arity = len(p.targets[0].elts)
            out["calls"].append({"file": rel, "line": node.lineno, "kind": kind, "module": module,
                                 "keywords": [k.arg for k in node.keywords], "positional": len(node.args), "unpacked_into": arity})
    with open(os.path.join(HERE, "reference_callsites.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(len(out["imports"]), "imports,", len(out["calls"]), "calls")
    for c in out["calls"]:
        print(c["file"], c["line"], c["kind"], c["module"], c["keywords"], c["positional"], c["unpacked_into"])


if __name__ == "__main__":
    main()
