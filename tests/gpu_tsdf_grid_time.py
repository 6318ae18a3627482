"""Developer helper: BoundedTSDFVolume.integrate at mesh-extraction scale (512^3 voxels x 32 views @ 1600x1060, depth only and
depth + RGB), then the mesh of that volume.  python tests/gpu_tsdf_grid_time.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs-sr_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from gsr_b200.tsdf import BoundedTSDFVolume
from gsr_b200.mesh import post_process_mesh
from tsdf_synth import build_tsdf_case

c = build_tsdf_case("bench")
projs = [torch.from_numpy(m).cuda() for m in c["projs"]]
depths = [torch.from_numpy(d).cuda() for d in c["depthmaps"]]
rgbs = [torch.rand(3, d.shape[-2], d.shape[-1], device="cuda") for d in depths]
n = int(os.environ.get("N", 512))
for with_rgb in (False, True):
    def run():
        vol = BoundedTSDFVolume((-1.2, -1.2, -1.2), 2.4 / (n - 1), (n, n, n), 5 * 2.4 / (n - 1), 10.0, with_rgb=with_rgb)
        vol.integrate(projs, depths, rgbs if with_rgb else None)
        return vol
    vol = run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); vol = run(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print(f"integrate {n}^3 x {len(projs)} views rgb={with_rgb}: {t:.2f} ms = {n**3*len(projs)/t/1e6:.1f} G voxel-views/s", flush=True)
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    vol.extract_triangle_mesh(); torch.cuda.synchronize()
    m0.record(); mesh = vol.extract_triangle_mesh(); m1.record(); torch.cuda.synchronize()
    print(f"   mesh: {m0.elapsed_time(m1):.2f} ms, V={mesh.vertices.shape[0]} F={mesh.triangles.shape[0]}", flush=True)
    post = post_process_mesh(mesh, cluster_to_keep=50)
    print(f"   post: {post.triangles.shape[0]} triangles kept", flush=True)
