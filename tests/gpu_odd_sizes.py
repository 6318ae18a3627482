import os, sys
sys.path.insert(0, "/root/repo/gs-sr_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import harness as hz, synth
for (w, h) in ((5, 200), (4097, 3), (3, 4097), (200, 5), (2000, 16)):
    sc = synth.make_scene(3000, w, h, seed=4, sigma_px=2.0)
    gc, go = synth.make_upstream_grads(w, h, seed=5)
    out = hz.run_product_surfel(sc, gc, go)
    ref = hz.run_refcuda_surfel(sc, gc, go)
    orc = hz.run_oracle_surfel(sc, gc, go)
    e = np.abs(out["color"] - ref["color"]).max(); eo = np.abs(orc["color"] - ref["color"]).max()
    print(f"{w}x{h}: ours-vs-ref color {e:.2e} radii eq {np.array_equal(out['radii'], ref['radii'])} | oracle-vs-ref color {eo:.2e} radii eq {np.array_equal(orc['radii'], ref['radii'])}", flush=True)
    for k in out["grads"]:
        print("   ", k, hz.rel_linf(out["grads"][k], ref["grads"][k]))
    bad = np.nonzero(out["radii"] != ref["radii"])[0][:5]
    print("    radii diffs", bad, out["radii"][bad], ref["radii"][bad])
