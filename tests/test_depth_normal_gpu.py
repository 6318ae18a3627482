"""GPU (pytest -m gpu): the fused PGSR depth->normal kernels (csrc/depth_normal.cu through gsr_b200.depth_normal and the C
ABI) against the reference-derived golden vectors and the float64 oracle.

Tolerances: north_star's forward 1e-4 / gradient 1e-3, relative to the tensor's max.  Normals are unit vectors, so 1e-4
is absolute there.  The reference evaluates x z / fx - cx z / fx with float32 cancellation, so its own output sits up to
~1e-4 from a float64 evaluation (tests/test_depth_normal_oracle_golden.py); the kernel forms the point in the same order
(the 3-term row-vector x matrix product) to stay inside the tolerance."""
import os

import numpy as np
import pytest
import torch

from depth_normal_synth import DN_CASES, build_dn_case
from oracle import depth_normal_oracle as dno

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run(c, weighted):
    from gsr_b200.depth_normal import render_normal_weighted
    d = torch.from_numpy(c["depth"]).cuda().requires_grad_(True)
    w = torch.from_numpy(c["weight"]).cuda() if weighted else None
    n = render_normal_weighted(d, torch.from_numpy(c["K"]), w)
    n.backward(torch.from_numpy(c["g"]).cuda())
    return n.detach().cpu().numpy(), d.grad.cpu().numpy()


@pytest.mark.parametrize("name", DN_CASES)
@pytest.mark.parametrize("weighted", [False, True])
def test_kernel_matches_reference_vectors(name, weighted):
    c = build_dn_case(name)
    g = np.load(os.path.join(GOLD, f"depth_normal_{name}.npz"))
    tag = "_w" if weighted else ""
    n, gd = run(c, weighted)
    assert np.abs(n - g["normal" + tag]).max() <= 1e-4, np.abs(n - g["normal" + tag]).max()
    assert np.abs(gd - g["grad" + tag]).max() <= 1e-3 * np.abs(g["grad" + tag]).max(), np.abs(gd - g["grad" + tag]).max() / np.abs(g["grad" + tag]).max()
    assert np.all(n[:, 0, :] == 0) and np.all(n[:, -1, :] == 0) and np.all(n[:, :, 0] == 0) and np.all(n[:, :, -1] == 0)


def test_reference_signature_and_full_resolution_against_float64():
    """normal_from_depth_image(depth, K, extrinsic) -> (H, W, 3) like the reference; at 1600x900 the kernel must be as
    close to a float64 evaluation as the reference's own float32 chain is (both are bounded by the same cancellation)."""
    from gsr_b200.depth_normal import normal_from_depth_image
    H, W = 900, 1600
    rng = np.random.default_rng(5)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    depth = (5.0 + 0.002 * xs - 0.003 * ys + 0.2 * np.sin(xs / 40.0) * np.cos(ys / 30.0)).astype(np.float32)
    K = np.array([[1.2 * W, 0, W / 2.0], [0, 1.2 * W, H / 2.0], [0, 0, 1]], np.float32)
    g = (rng.normal(size=(3, H, W)) / (W * H)).astype(np.float32)
    d = torch.from_numpy(depth).cuda().requires_grad_(True)
    n = normal_from_depth_image(d, torch.from_numpy(K), torch.eye(4))
    assert n.shape == (H, W, 3)
    n.permute(2, 0, 1).backward(torch.from_numpy(g).cuda())
    n64, g64 = dno.value_and_grad(depth, K, g, None, dtype=torch.float64)
    n32, g32 = dno.value_and_grad(depth, K, g, None, dtype=torch.float32)            # the reference's arithmetic on the CPU
    err_k = np.abs(n.detach().permute(2, 0, 1).cpu().numpy() - n64).max()
    err_r = np.abs(n32 - n64).max()
    print(f"normal error vs float64: kernel {err_k:.2e}, reference chain {err_r:.2e}")
    assert err_k <= max(2 * err_r, 1e-4)
    ge_k = np.abs(d.grad.cpu().numpy() - g64).max() / np.abs(g64).max()
    ge_r = np.abs(g32 - g64).max() / np.abs(g64).max()
    print(f"grad error vs float64: kernel {ge_k:.2e}, reference chain {ge_r:.2e}")
    assert ge_k <= max(2 * ge_r, 1e-3)


def test_rejects_offsets_and_cpu_tensors():
    from gsr_b200.depth_normal import normal_from_depth_image
    K = torch.eye(3)
    with pytest.raises(NotImplementedError):
        normal_from_depth_image(torch.ones(8, 8).cuda(), K, None, offset=torch.zeros(8, 8, 8).cuda())
    with pytest.raises(RuntimeError):
        normal_from_depth_image(torch.ones(8, 8), K)
