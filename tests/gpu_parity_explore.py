"""Developer helper (not a pytest): product vs the reference CUDA build at the headline sizes; dumps the deviating
pixels / Gaussians to gpurun_out/ for offline analysis of what the parity tests may attribute to threshold flips."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz, synth
import torch

P, W, H = 2_000_000, 1600, 1060
if len(sys.argv) > 3:
    P, W, H = map(int, sys.argv[1:4])
tag = sys.argv[4] if len(sys.argv) > 4 else "cfgB"
sc = synth.make_scene(P, W, H, seed=0)
gc, go = synth.make_upstream_grads(W, H, seed=1)
tt = hz.to_torch(sc)
out = hz.run_product_surfel(sc, gc, go, tt=tt)
ref = hz.run_refcuda_surfel(sc, gc, go, tt=tt)
N = W * H
chan = [("color0", out["color"][0], ref["color"][0]), ("color1", out["color"][1], ref["color"][1]),
        ("color2", out["color"][2], ref["color"][2])] + [(f"o{c}", out["others"][c], ref["others"][c]) for c in range(11)]
dev = np.zeros((len(chan), N), np.float32)
scale = np.zeros(len(chan))
for i, (nm, a, b) in enumerate(chan):
    scale[i] = max(np.abs(b).max(), 1e-30)
    dev[i] = (np.abs(a.astype(np.float64) - b) / scale[i]).ravel()
print("radii mismatches:", int((out["radii"] != ref["radii"]).sum()), "of", P)
for i, (nm, _, _) in enumerate(chan):
    d = dev[i]
    print(f"{nm:7s} scale={scale[i]:.3e} max={d.max():.2e} frac>1e-4={np.mean(d > 1e-4):.2e} frac>1e-5={np.mean(d > 1e-5):.2e} "
          f"p99.9={np.quantile(d, 0.999):.2e} p99.99={np.quantile(d, 0.9999):.2e}")
idx_mis = (out["others"][7] != ref["others"][7]).ravel()
print("surf idx mismatch frac:", idx_mis.mean())
worst = dev[[0, 1, 2, 3, 4, 5, 6, 7, 9]].max(axis=0)          # colour, depth, alpha, normal, distortion (not median-selected)
sel = np.where((worst > 2e-5) | idx_mis)[0]
print("pixels dumped:", sel.size, " of which idx mismatch:", int(idx_mis[sel].sum()),
      " dev>1e-4:", int((worst > 1e-4).sum()), " dev>1e-4 & idx mismatch:", int(((worst > 1e-4) & idx_mis).sum()))
os.makedirs("gpurun_out", exist_ok=True)
dump = dict(pix=sel.astype(np.int32), dev=dev[:, sel], idx_ours=out["others"][7].ravel()[sel], idx_ref=ref["others"][7].ravel()[sel],
            alpha_ref=ref["others"][1].ravel()[sel], scale=scale)
# gradients
print("grads ours vs ref:")
for k in ("means3D", "means2D", "colors", "opacities", "scales", "rotations"):
    a = out["grads"][k].astype(np.float64); b = ref["grads"][k].astype(np.float64).reshape(a.shape)
    mx = max(np.abs(b).max(), 1e-30)
    d = np.abs(a - b).reshape(P, -1).max(1) / mx
    l2 = np.linalg.norm(a - b) / np.linalg.norm(b)
    print(f"  {k:10s} max|ref|={mx:.3e} relLinf={d.max():.2e} relL2={l2:.2e} frac>1e-3={np.mean(d > 1e-3):.2e} frac>1e-4={np.mean(d > 1e-4):.2e} "
          f"p99.9={np.quantile(d, 0.999):.2e} p99.99={np.quantile(d, 0.9999):.2e}")
    gs = np.where(d > 1e-4)[0][:200000]
    dump[f"g_{k}_idx"] = gs.astype(np.int32); dump[f"g_{k}_dev"] = d[gs].astype(np.float32)
    dump[f"g_{k}_ours"] = a[gs].astype(np.float32); dump[f"g_{k}_ref"] = b[gs].astype(np.float32)
np.savez_compressed(f"gpurun_out/parity_explore_{tag}.npz", **dump)
print("saved")
