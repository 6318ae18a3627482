"""GPU: one GS-SR-style 2DGS training iteration (tests/train_harness.py) driven through the drop-in rasterizer
vs the same iteration driven through the unmodified reference kernels (oracle/_ref): losses, the means2D
densification statistic and the parameters after Adam agree."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _ref_available():
    from oracle import refcuda
    return refcuda.available("surfel")


def test_training_iteration_matches_reference_kernels():
    if not _ref_available():
        pytest.skip("oracle/_ref/libref_surfel.so did not travel")
    from train_harness import MiniTwoDGSTrainer
    kw = dict(P=20000, W=256, H=192, seed=5, lambda_dist=100.0)
    a, b = MiniTwoDGSTrainer(impl="ours", **kw), MiniTwoDGSTrainer(impl="reference", **kw)
    la, da = a.step()
    lb, db = b.step()
    for k in da:
        assert abs(da[k] - db[k]) <= 1e-5 * max(abs(db[k]), 1e-3), (k, da[k], db[k])
    # densification statistic = norm of the returned means2D gradient (vanilla_gaussian.py:428-430)
    ga, gb = a.xyz_gradient_accum.double(), b.xyz_gradient_accum.double()
    assert float((ga - gb).norm() / gb.norm()) <= 1e-3
    assert torch.equal(a.denom, b.denom) and torch.equal(a.max_radii2D, b.max_radii2D)
    # two more steps: Adam amplifies sign flips of ~0 gradients, so compare losses only, loosely
    for _ in range(2):
        la, _ = a.step()
        lb, _ = b.step()
    assert abs(la - lb) <= 2e-3 * abs(lb), (la, lb)


def test_training_loss_decreases_with_dropin():
    from train_harness import MiniTwoDGSTrainer
    tr = MiniTwoDGSTrainer(P=20000, W=256, H=192, seed=6, impl="ours", lambda_dist=0.0)
    first = tr.step()[0]
    for _ in range(25):
        last = tr.step()[0]
    assert np.isfinite(last) and last < first
    assert float(tr.denom.max()) == 26.0


def test_scaffold_2dgs_iteration_matches_reference_kernels():
    """config-3 flow: visible_filter on a (N,6)[:, :3] slice, opacity-masked neural Gaussians, colors_precomp and the
    stride-3 scaling[:, :2] view, anchor statistics."""
    if not _ref_available():
        pytest.skip("oracle/_ref/libref_surfel.so did not travel")
    from train_harness import MiniScaffold2DGSTrainer
    kw = dict(n_anchor=8000, k=5, W=256, H=192, seed=9, lambda_dist=100.0)
    a, b = MiniScaffold2DGSTrainer(impl="ours", **kw), MiniScaffold2DGSTrainer(impl="reference", **kw)
    la, da = a.step()
    lb, db = b.step()
    assert a.last == b.last and a.last["rendered"] > 1000, (a.last, b.last)
    for k in da:
        assert abs(da[k] - db[k]) <= 1e-5 * max(abs(db[k]), 1e-3), (k, da[k], db[k])
    ga, gb = a.offset_gradient_accum.double(), b.offset_gradient_accum.double()
    assert float((ga - gb).norm() / gb.norm()) <= 1e-3
    assert torch.equal(a.offset_denom, b.offset_denom) and torch.equal(a.anchor_demon, b.anchor_demon)
    assert torch.allclose(a.opacity_accum, b.opacity_accum)


def test_pgsr_iteration_matches_reference_kernels():
    """config-4 flow: two views per iteration through the plane rasterizer, autograd-built all_map, means2D_abs, out_observe."""
    from oracle import refcuda
    if not refcuda.available("plane"):
        pytest.skip("oracle/_ref/libref_plane.so did not travel")
    from train_harness import MiniPGSRTrainer
    kw = dict(P=20000, W=256, H=144, seed=13)
    a, b = MiniPGSRTrainer(impl="ours", **kw), MiniPGSRTrainer(impl="reference", **kw)
    la, da = a.step()
    lb, db = b.step()
    assert a.last == b.last and a.last["observed"] > 1000, (a.last, b.last)
    for k in da:
        assert abs(da[k] - db[k]) <= 1e-5 * max(abs(db[k]), 1e-3), (k, da[k], db[k])
    for x, y in ((a.xyz_gradient_accum, b.xyz_gradient_accum), (a.xyz_gradient_accum_abs, b.xyz_gradient_accum_abs)):
        assert float((x.double() - y.double()).norm() / y.double().norm()) <= 1e-3
    assert torch.equal(a.denom, b.denom) and torch.equal(a.max_radii2D, b.max_radii2D)
    for _ in range(2):
        la, _ = a.step()
        lb, _ = b.step()
    assert abs(la - lb) <= 2e-3 * abs(lb), (la, lb)


def test_fused_depth_normal_in_the_pgsr_iteration():
    """Swapping the torch normal_from_depth_image chain x alpha for gsr_b200.depth_normal leaves the PGSR iteration's losses
    and densification statistics unchanged (the normal loss is the only consumer, pgsr_scene.py:105-113)."""
    from train_harness import MiniPGSRTrainer
    kw = dict(P=20000, W=256, H=144, seed=13, impl="ours")
    a, b = MiniPGSRTrainer(**kw), MiniPGSRTrainer(**kw)
    a.fused_post = True
    la, da = a.step()
    lb, db = b.step()
    for k in da:
        assert abs(da[k] - db[k]) <= 2e-5 * max(abs(db[k]), 1e-3), (k, da[k], db[k])
    for x, y in ((a.xyz_gradient_accum, b.xyz_gradient_accum), (a.xyz_gradient_accum_abs, b.xyz_gradient_accum_abs)):
        assert float((x.double() - y.double()).norm() / y.double().norm()) <= 1e-3
    for _ in range(2):
        la, _ = a.step()
        lb, _ = b.step()
    assert abs(la - lb) <= 2e-3 * abs(lb), (la, lb)


def test_fused_ssim_in_the_training_iteration():
    """Swapping VanillaScene.ssim for gsr_b200.ssim leaves the iteration's losses and statistics unchanged."""
    from train_harness import MiniTwoDGSTrainer
    kw = dict(P=20000, W=256, H=192, seed=5, lambda_dist=100.0, impl="ours")
    a, b = MiniTwoDGSTrainer(**kw), MiniTwoDGSTrainer(**kw)
    a.fused_ssim = True
    la, da = a.step()
    lb, db = b.step()
    for k in da:
        assert abs(da[k] - db[k]) <= 1e-5 * max(abs(db[k]), 1e-3), (k, da[k], db[k])
    ga, gb = a.xyz_gradient_accum.double(), b.xyz_gradient_accum.double()
    assert float((ga - gb).norm() / gb.norm()) <= 1e-4
