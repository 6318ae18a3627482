"""GPU: fused SSIM (gsr_b200.ssim -> gsr_ssim_forward/backward) against the reference's golden vectors, against the
float64 oracle at the bench resolution, plus identities (ssim(x, x) = 1 with zero gradient; scaling of the upstream)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from ssim_synth import SSIM_CASES, build_ssim_case  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VALUE_TOL = 1e-5     # absolute, on an SSIM in [-1, 1]  (north_star forward tolerance 1e-4)
GRAD_TOL = 1e-4      # rel-Linf of the gradient image   (north_star gradient tolerance 1e-3)


def run(a, b, upstream=1.0):
    from gsr_b200.ssim import ssim
    x = torch.from_numpy(a).cuda().requires_grad_(True)
    v = ssim(x, torch.from_numpy(b).cuda())
    (v * upstream).backward()
    return float(v.detach()), x.grad.cpu().numpy()


@pytest.mark.parametrize("name", list(SSIM_CASES))
def test_matches_reference_golden(name):
    a, b = build_ssim_case(name)
    gold = np.load(os.path.join(GOLD, f"ssim_{name}.npz"))
    v, g = run(a, b)
    assert abs(v - float(gold["value"])) <= VALUE_TOL
    assert np.abs(g - gold["grad"]).max() <= GRAD_TOL * np.abs(gold["grad"]).max()


def test_full_resolution_vs_float64_oracle():
    from oracle import ssim_oracle
    rng = np.random.default_rng(9)
    H, W = 1060, 1600
    yy, xx = np.meshgrid(np.arange(H) / H, np.arange(W) / W, indexing="ij")
    b = np.stack([0.5 + 0.4 * np.sin(6 * xx + 2 * yy + c) for c in range(3)]).astype(np.float32)
    a = np.clip(b + 0.1 * rng.normal(size=b.shape), 0, 1).astype(np.float32)
    v, g = run(a, b, upstream=-0.2)
    ov, og = ssim_oracle.ssim_value_and_grad(a, b, dtype=torch.float64)
    assert abs(v - ov) <= VALUE_TOL
    assert np.abs(g - (-0.2) * og).max() <= GRAD_TOL * np.abs(0.2 * og).max()


def test_identities_and_errors():
    from gsr_b200.ssim import ssim
    a, _ = build_ssim_case("small")
    v, g = run(a, a.copy())
    assert abs(v - 1.0) <= 1e-6 and np.abs(g).max() <= 1e-6
    x = torch.rand((2, 3, 40, 50), device="cuda", requires_grad=True)
    y = torch.rand((2, 3, 40, 50), device="cuda")
    v4 = ssim(x, y)
    v3 = torch.stack([ssim(x[i], y[i]) for i in range(2)]).mean()
    assert abs(float(v4) - float(v3)) <= 1e-6
    with pytest.raises(NotImplementedError):
        ssim(x, y, window_size=7)
    with pytest.raises(RuntimeError):
        ssim(x.cpu(), y.cpu())
    with pytest.raises(RuntimeError):
        ssim(x, y.requires_grad_(True)).backward()
