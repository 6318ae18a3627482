"""Developer helper (not a pytest): A/B kernel variants selected by gsr_set_option("dbg", v) on cfg-B:
library cudaEvent times per kernel + gradient agreement with the default variant.  dbg bit 0 = per-tile bitonic
network instead of the bucketed sort; other library builds can be compared with GSR_B200_LIB=/path/to/lib.so.

    python tests/gpu_variants.py 0 1
"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_profile as gp
import torch, gsr_b200
P, W, H = 2_000_000, 1600, 1060
variants = [int(v) for v in sys.argv[1:]] or [0, 1]
sc, tt, gct, got, rast, leaves, m2d = gp.setup(P, W, H)
L = gsr_b200.lib()


def run(v, n=5):
    L.gsr_set_option(b"dbg", v)
    for _ in range(2):
        gp.product_step(rast, leaves, m2d, gct, got)
    acc = np.zeros(16)
    for _ in range(n):
        for t in leaves.values():
            t.grad = None
        L.gsr_profile_enable(1)
        gp.product_step(rast, leaves, m2d, gct, got)
        buf = (ctypes.c_float * 16)(); L.gsr_profile_read(buf); acc += np.array(list(buf))
    L.gsr_profile_enable(0)
    torch.cuda.synchronize()
    return {k: t.grad.clone() for k, t in leaves.items()}, acc / n


base = None
for v in variants:
    g, t = run(v)
    msg = f"dbg={v}: render_fwd={t[5]*1e3:.0f}us render_bwd={t[6]*1e3:.0f}us build={t[4]*1e3:.0f}us"
    if base is None:
        base = g
    else:
        msg += " | grad rel-max diff " + " ".join(f"{k}={float((g[k]-base[k]).abs().max()/base[k].abs().max()):.1e}" for k in g)
    print(msg)
L.gsr_set_option(b"dbg", 0)
