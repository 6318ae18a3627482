"""Developer helper (not a pytest): eager vs CUDA-graph iterations/s of the 2DGS harness iteration (config 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from train_harness import measure_iters_per_s, measure_graph_iters_per_s
P, W, H = 100_000, 800, 800
v, _ = measure_iters_per_s("ours", P, W, H, iters=40, warmup=5, fused_ssim=True, fused_post=True)
print(f"eager, fused ops:      {v:8.1f} it/s")
g, loss, ov = measure_graph_iters_per_s(P, W, H, iters=200, warmup=5)
print(f"one CUDA graph/iter:   {g:8.1f} it/s  (loss {loss:.5f}, capture_overflow {ov})")
try:
    r, _ = measure_iters_per_s("reference", P, W, H, iters=40, warmup=5)
    print(f"reference kernels:     {r:8.1f} it/s")
except Exception as e:
    print("reference arm unavailable:", e)


def raster_only(P, W, H, sh):
    """fwd+bwd of the surfel rasterizer alone: eager calls vs replays of one captured graph."""
    import gpu_profile as gp
    sc, tt, gct, got, rast, leaves, m2d = gp.setup(P, W, H, sh)

    def step():
        m = torch.zeros_like(leaves["means3D"], requires_grad=True)
        color, radii, others = rast(means3D=leaves["means3D"], means2D=m, opacities=leaves["opacities"],
                                    colors_precomp=leaves.get("colors"), shs=leaves.get("shs"), scales=leaves["scales"],
                                    rotations=leaves["rotations"])
        torch.autograd.backward([color, others], [gct, got])

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            for v in leaves.values():
                v.grad = None
            step()
        e = timed(lambda: ([setattr(v, "grad", None) for v in leaves.values()], step()), 30)
    torch.cuda.current_stream().wait_stream(side)
    for v in leaves.values():
        v.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    r = timed(g.replay, 30)
    print(f"surfel fwd+bwd P={P} {W}x{H} sh={sh}: eager {e:.3f} ms, graph replay {r:.3f} ms")


raster_only(100_000, 800, 800, True)
raster_only(2_000_000, 1600, 1060, False)
