"""Developer helper (not a pytest): TSDF fusion throughput on a 256^3 chunk x 32 views @ 1600x1060 --
gsr_tsdf_fuse vs the reference's per-view torch rule (restated with the same torch ops, on the GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs-sr_b200"))
import numpy as np, torch
from tsdf_synth import build_tsdf_case
from gsr_b200.tsdf import TSDFFusion

c = build_tsdf_case("bench")
Ng = 256
ax = torch.linspace(-1.2, 1.2, Ng, device="cuda")
xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
pts = torch.stack([xx.ravel(), yy.ravel(), zz.ravel()], -1).contiguous()
projs = [torch.from_numpy(m).cuda() for m in c["projs"]]
depths = [torch.from_numpy(d).cuda() for d in c["depthmaps"]]
rgbs = [torch.from_numpy(r).cuda() for r in c["rgbmaps"]]
center = torch.from_numpy(c["center"]).cuda(); radius = c["radius"]; vs = c["voxel_size"]
f = TSDFFusion(projs, depths, rgbs, center=c["center"], radius=radius)


def torch_rule(samples):
    from tsdf_synth import torch_rule_unbounded
    return torch_rule_unbounded(samples, projs, depths, rgbs, center, radius, vs)


def timed(fn, n):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out

for vpl in (32, 16, 8, 4, 2):
    f.views_per_launch = vpl
    t, _ = timed(lambda: f.compute_unbounded_tsdf(pts, True, vs), 5)
    t2, _ = timed(lambda: f.compute_unbounded_tsdf(pts, True, vs, return_rgb=True), 3)
    print(f"views_per_launch={vpl}: {t:.2f} ms, with rgb {t2:.2f} ms")
f.views_per_launch = None
print("default groups:", f._launch_groups(1), f._launch_groups(4))
t_ours, a = timed(lambda: f.compute_unbounded_tsdf(pts, True, vs), 5)
t_rgb, _ = timed(lambda: f.compute_unbounded_tsdf(pts, True, vs, return_rgb=True), 3)
t_ref, b = timed(lambda: torch_rule(pts), 2)
n, V = pts.shape[0], len(projs)
d = (a - b).abs()
print(f"samples={n} views={V}: gsr_tsdf_fuse {t_ours:.2f} ms ({n*V/t_ours/1e6:.1f} G sample-views/s), with rgb {t_rgb:.2f} ms; "
      f"torch rule {t_ref:.1f} ms -> x{t_ref/t_ours:.1f}; max|diff|={d.max().item():.3g} frac>2e-5={(d>2e-5).float().mean().item():.2e}")
print(f"HBM algorithmic bytes {n*16/1e6:.0f} MB -> {n*16/t_ours/1e6:.1f} GB/s")
