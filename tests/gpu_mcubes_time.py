"""Developer helper: marching-cubes kernel times on a 512^3 lattice (GPU box).  python tests/gpu_mcubes_time.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs-sr_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from gsr_b200.mesh import extract_triangle_mesh

n = int(os.environ.get("N", 512))
ax = torch.arange(n, device="cuda", dtype=torch.float32)
z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
c = (n - 1) / 2
f = torch.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2) - 0.35 * n
f = f + 3.0 * torch.sin(x * 0.21) * torch.sin(y * 0.17) * torch.sin(z * 0.13)      # a bumpy sphere
w = torch.full_like(f, 2.0)
rgb = torch.rand((n, n, n, 3), device="cuda")
del x, y, z
for name, kw in (("tsdf only", {}), ("weights", dict(weight=w, min_weight=1.0)), ("weights+rgb", dict(weight=w, min_weight=1.0, rgb=rgb))):
    for _ in range(2):
        m = extract_triangle_mesh(f, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        m = extract_triangle_mesh(f, **kw)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / K:.3f} ms per extraction, V={m.vertices.shape[0]} F={m.triangles.shape[0]}")
# CPU oracle on a 128^3 crop for scale
from oracle import mcubes_oracle as mc
fc = f[192:320, 192:320, 20:148].cpu().numpy()
t = time.time(); v, fa, _ = mc.extract(fc); dt = time.time() - t
print(f"numpy oracle 128^3 crop: {dt*1e3:.1f} ms, F={len(fa)}")
# cluster filter on the 512^3 mesh
from gsr_b200.mesh import post_process_mesh
m = extract_triangle_mesh(f)
for _ in range(2):
    p = post_process_mesh(m, cluster_to_keep=50)
torch.cuda.synchronize()
t = time.time()
for _ in range(5):
    p = post_process_mesh(m, cluster_to_keep=50)
torch.cuda.synchronize()
print(f"post_process_mesh: {(time.time() - t) / 5 * 1e3:.2f} ms wall, {m.triangles.shape[0]} -> {p.triangles.shape[0]} triangles")
