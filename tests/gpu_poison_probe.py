"""Developer helper: poisoned inputs (NaN / Inf / zero / huge values in a few Gaussians) through the three rasterizers,
ours vs the reference build: no crash, same visible set, same pixels where the reference itself is finite.
python tests/gpu_poison_probe.py   (run under compute-sanitizer for the memory-safety half)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs-sr_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import harness as hz
import synth
from oracle import refcuda


from poison_synth import KINDS, poison  # noqa: E402


def summarize(name, kind, ours, ref):
    ro, rr = ours["radii"], ref["radii"]
    vis_same = np.array_equal(ro > 0, rr > 0)
    fin_r = np.isfinite(ref["color"])
    fin_o = np.isfinite(ours["color"])
    both = fin_r & fin_o
    err = np.abs(ours["color"][both] - ref["color"][both]).max() if both.any() else 0.0
    nbad = int((np.abs(ours["color"] - ref["color"])[both] > 1e-3).sum())
    print(f"{name:8s} {kind:14s} visible same={vis_same} (ours {int((ro>0).sum())} ref {int((rr>0).sum())}) radii eq={np.array_equal(ro, rr)} "
          f"finite px ours {fin_o.mean():.4f} ref {fin_r.mean():.4f} nan-mask eq={np.array_equal(fin_r, fin_o)} max err {err:.2e} px>1e-3: {nbad}", flush=True)
    for k in ours.get("grads", {}):
        go, gr = ours["grads"][k], ref["grads"][k]
        fo, fr = np.isfinite(go), np.isfinite(gr)
        b = fo & fr
        scale = np.abs(gr[b]).max() if b.any() else 1.0
        e = np.abs(go[b] - gr[b]).max() / max(scale, 1e-30) if b.any() else 0.0
        print(f"          grad {k:10s} finite ours {fo.mean():.5f} ref {fr.mean():.5f} mask eq={np.array_equal(fo, fr)} rel err {e:.2e}", flush=True)


kinds = list(KINDS)
only = os.environ.get("KINDS")
if only:
    kinds = only.split(",")
W, H, P = 320, 240, 20000
for kind in kinds:
    rng = np.random.default_rng(5)
    sc = synth.make_scene(P, W, H, seed=11)
    poison(sc, kind, rng, 2)
    gc, go = synth.make_upstream_grads(W, H, seed=12)
    ours = hz.run_product_surfel(sc, gc, go)
    torch.cuda.synchronize()
    ref = hz.run_refcuda_surfel(sc, gc, go) if refcuda.available("surfel") else None
    if ref is not None:
        summarize("surfel", kind, ours, ref)
    for plane in (False, True):
        rng = np.random.default_rng(5)
        sc = synth.make_scene(P, W, H, seed=11, scale_dims=3)
        poison(sc, kind, rng, 3)
        kw = {}
        if plane:
            gc2, go2 = synth.make_upstream_grads(W, H, seed=6, n_others=6, zero_from=6)
            kw = dict(all_map=synth.make_all_map(sc), g_all_map=np.ascontiguousarray(go2[:5]), g_plane_depth=np.ascontiguousarray(go2[5:6]))
        ours = hz.run_product_gauss(sc, gc, plane=plane, **kw)
        torch.cuda.synchronize()
        var = "plane" if plane else "gaussian"
        if refcuda.available(var):
            ref = hz.run_refcuda_gauss(sc, gc, plane=plane, **kw)
            summarize(var, kind, ours, ref)
print("done")
