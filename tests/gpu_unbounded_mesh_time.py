"""Developer helper: TSDFFusion.extract_mesh_unbounded at the reference's chunk size (resolution 512 = one 512^3 lattice, and
1024 = 1023^3 samples meshed in one piece) on the synthetic bench views.  python tests/gpu_unbounded_mesh_time.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs-sr_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from gsr_b200.tsdf import TSDFFusion
from tsdf_synth import build_tsdf_case

c = build_tsdf_case("bench")
f = TSDFFusion([torch.from_numpy(m) for m in c["projs"]], [torch.from_numpy(d) for d in c["depthmaps"]], None,
               center=c["center"], radius=c["radius"])
xyz = torch.from_numpy(c["center"]).cuda() + 0.4 * c["radius"] * torch.randn(100000, 3, device="cuda")
for res in [int(v) for v in os.environ.get("RES", "512,1024").split(",")]:
    torch.cuda.synchronize(); t = time.time()
    mesh = f.extract_mesh_unbounded(resolution=res, gaussians_xyz=xyz)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f"resolution {res}: {dt*1e3:.0f} ms, V={mesh.vertices.shape[0]} F={mesh.triangles.shape[0]}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
