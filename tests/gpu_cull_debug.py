"""Developer helper: find pixels where the forward with / without the contribution boxes differ at cfg-B and dump the
tile's record stream for offline analysis of the cull test (gpurun_out/cull_debug.npz)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as hz, synth
import torch, gsr_b200
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer

P, W, H = 2_000_000, 1600, 1060
sc = synth.make_scene(P, W, H, seed=0)
tt = hz.to_torch(sc)
L = gsr_b200.lib()
rs = GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"], tt["proj"], 0, tt["campos"], False, False)


def fwd():
    m3 = tt["means3D"].clone().requires_grad_(True)
    return GaussianRasterizer(rs)(means3D=m3, means2D=torch.zeros_like(m3), opacities=tt["opacities"], colors_precomp=tt["colors"],
                                  scales=tt["scales"], rotations=tt["rotations"])


L.gsr_set_option(b"no_cull", 1)
c0, r0, o0 = fwd()
L.gsr_set_option(b"no_cull", 0)
c1, r1, o1 = fwd()
diff = ((c0 != c1).any(0) | (o0[:7] != o1[:7]).any(0))
pix = diff.flatten().nonzero().flatten().cpu().numpy()
print("differing pixels:", pix.size, [(int(p % W), int(p // W)) for p in pix[:10]])
fn = c1.grad_fn
geom, binning, image = fn.saved_tensors[-3:]
R = fn.num_rendered
N = W * H
gx, gy = (W + 15) // 16, (H + 15) // 16
ntiles = gx * gy


def al(x, a=128):
    return (x + a - 1) // a * a


ib = image.cpu().numpy(); base = (-image.data_ptr()) % 256
off = 0
off = al(off); off_finalT = off; off += 3 * N * 4
off = al(off); off_ncontrib = off; off += 2 * N * 4
off = al(off); off_tc = off; off += ntiles * 32 * 4
off = al(off); off_to = off; off += (ntiles + 1) * 4
tile_offset = ib[base + off_to: base + off_to + (ntiles + 1) * 4].view(np.uint32)
final_T = ib[base + off_finalT: base + off_finalT + N * 4].view(np.float32)
n_contrib = ib[base + off_ncontrib: base + off_ncontrib + N * 4].view(np.uint32)
bbase = (-binning.data_ptr()) % 256
n = max(R, 1)
koff = al(0); poff = al(koff + n * 8); pstride = (n + 7) & ~7
dump = dict(pix=pix, W=W, H=H)
for k, p in enumerate(pix[:8]):
    x, y = int(p % W), int(p // W)
    t = (y // 16) * gx + (x // 16)
    b0, b1 = int(tile_offset[t]), int(tile_offset[t + 1])
    planes = []
    for pl in range(6):
        st = bbase + poff + (pl * pstride + b0) * 16
        planes.append(binning[st: st + (b1 - b0) * 16].cpu().numpy().view(np.float32).reshape(-1, 4))
    dump[f"planes{k}"] = np.stack(planes)
    dump[f"meta{k}"] = np.array([x, y, t, b0, b1, final_T[p], n_contrib[p]], np.float64)
    print("pixel", x, y, "tile", t, "entries", b1 - b0, "T", final_T[p], "last", n_contrib[p], "color cull", c1[:, y, x].tolist(), "nocull", c0[:, y, x].tolist())
np.savez_compressed("gpurun_out/cull_debug.npz", **dump)
