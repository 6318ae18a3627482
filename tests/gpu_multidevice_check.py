"""Developer check (needs 2 GPUs, not a pytest): tensors on cuda:1 while the current device is cuda:0 -- every drop-in
must run on the tensors' device and stream and give the same results as on cuda:0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import harness as hz, synth
assert torch.cuda.device_count() >= 2
sc = synth.make_scene(20000, 320, 240, seed=21, sigma_px=3.0, rotate_camera=True, bg=(0.1, 0.2, 0.3))
gc, go = synth.make_upstream_grads(320, 240, seed=22)
torch.cuda.set_device(0)
a = hz.run_product_surfel(sc, gc, go, device="cuda:0")
b = hz.run_product_surfel(sc, gc, go, device="cuda:1")
assert np.array_equal(a["color"], b["color"]) and np.array_equal(a["radii"], b["radii"])
for k in a["grads"]:
    if a["grads"][k] is not None:
        d = np.abs(a["grads"][k] - b["grads"][k]).max() / max(np.abs(a["grads"][k]).max(), 1e-30)
        assert d < 1e-5, (k, d)
sc3 = synth.make_scene(20000, 320, 240, seed=23, sigma_px=3.0, scale_dims=3)
g3, _ = synth.make_upstream_grads(320, 240, seed=24)
a = hz.run_product_gauss(sc3, g_color=g3, device="cuda:0"); b = hz.run_product_gauss(sc3, g_color=g3, device="cuda:1")
assert np.array_equal(a["color"], b["color"])
from simple_knn._C import distCUDA2
pts = torch.from_numpy(synth.make_points(50000, seed=5))
assert torch.equal(distCUDA2(pts.to("cuda:0")).cpu(), distCUDA2(pts.to("cuda:1")).cpu())
from gsr_b200.ssim import ssim
x, y = torch.rand(3, 100, 120), torch.rand(3, 100, 120)
assert abs(float(ssim(x.to("cuda:1"), y.to("cuda:1"))) - float(ssim(x.to("cuda:0"), y.to("cuda:0")))) < 1e-7
s1 = torch.cuda.Stream(device="cuda:1")
with torch.cuda.stream(s1):
    c = hz.run_product_surfel(sc, gc, go, device="cuda:1")
print("multi-device ok: cuda:1 tensors with current device cuda:0 (default and side stream) match cuda:0")
