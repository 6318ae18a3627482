"""Developer helper (not a pytest): distCUDA2 time on point sets that are not uniform: planar, collinear, heavy duplicates,
one far outlier (degenerate bounding box for the Morton codes)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
import torch
from simple_knn._C import distCUDA2
from oracle import refcuda
N = 1_000_000
rng = np.random.default_rng(3)


def ev(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


cases = {}
cases["uniform cube"] = rng.random((N, 3), dtype=np.float32)
p = rng.random((N, 3), dtype=np.float32); p[:, 2] = 0.5; cases["planar (z constant)"] = p
p = rng.random((N, 3), dtype=np.float32); p[:, 1:] = 0.25; cases["collinear"] = p
p = rng.random((N // 4, 3), dtype=np.float32); cases["every point 4 times"] = np.repeat(p, 4, axis=0)
p = rng.random((N, 3), dtype=np.float32); p[0] = 1e4; cases["one far outlier"] = p
p = (rng.normal(0, 1, (N, 3)) * np.array([1, 1, 1e-3])).astype(np.float32); cases["gaussian slab"] = p
for name, pts in cases.items():
    t = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    a, out = ev(lambda: distCUDA2(t))
    msg = f"{name:22s} ours {a:8.2f} ms"
    if refcuda.available("knn"):
        b, ref = ev(lambda: refcuda.ref_dist2_knn3(t), 1)
        same = bool(torch.equal(out, ref))
        msg += f"  reference {b:8.2f} ms  x{b/a:.1f}  bit-identical={same}"
    print(msg, flush=True)
