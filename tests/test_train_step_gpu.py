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


def test_octree_2dgs_iteration_matches_reference_kernels():
    """config-5 per-tile flow (Octree2DGSScene): set_anchor_mask -> visible_filter on the masked anchors only -> neural
    Gaussians with the level input -> surfel rasterizer; both arms must keep the same anchors and agree on losses and
    anchor statistics."""
    if not _ref_available():
        pytest.skip("oracle/_ref/libref_surfel.so did not travel")
    from train_harness import MiniOctree2DGSTrainer
    kw = dict(n_anchor=8000, k=5, W=256, H=192, seed=21, lambda_dist=100.0)
    a, b = MiniOctree2DGSTrainer(impl="ours", **kw), MiniOctree2DGSTrainer(impl="reference", **kw)
    la, da = a.step()
    lb, db = b.step()
    assert torch.equal(a.anchor_mask, b.anchor_mask) and 0.2 < float(a.anchor_mask.float().mean()) < 0.98   # LOD mask is active
    assert a.last == b.last and a.last["rendered"] > 1000, (a.last, b.last)
    for k in da:
        assert abs(da[k] - db[k]) <= 1e-5 * max(abs(db[k]), 1e-3), (k, da[k], db[k])
    ga, gb = a.offset_gradient_accum.double(), b.offset_gradient_accum.double()
    assert float((ga - gb).norm() / gb.norm()) <= 1e-3
    assert torch.equal(a.offset_denom, b.offset_denom) and torch.equal(a.anchor_demon, b.anchor_demon)


def test_octree_pgsr_iteration_matches_reference_kernels():
    """config-4 flow (OctreePGSRScene): per view set_anchor_mask + octree prefilter + neural Gaussians through the plane
    rasterizer with the autograd-built all_map; two views per iteration."""
    from oracle import refcuda
    if not (refcuda.available("plane") and refcuda.available("filter")):
        pytest.skip("oracle/_ref/libref_plane.so / libref_filter.so did not travel")
    from train_harness import MiniOctreePGSRTrainer
    kw = dict(n_anchor=6000, k=5, W=256, H=144, seed=23)
    a, b = MiniOctreePGSRTrainer(impl="ours", **kw), MiniOctreePGSRTrainer(impl="reference", **kw)
    la, da = a.step()
    lb, db = b.step()
    assert a.last == b.last and a.last["gaussians"] > 1000, (a.last, b.last)
    for k in da:
        assert abs(da[k] - db[k]) <= 2e-5 * max(abs(db[k]), 1e-3), (k, da[k], db[k])
    ga, gb = a.offset_gradient_accum.double(), b.offset_gradient_accum.double()
    assert float((ga - gb).norm() / gb.norm()) <= 1e-3
    assert torch.equal(a.offset_denom, b.offset_denom)


def test_anchor_growing_selects_the_same_anchors():
    """ScaffoldGaussian.adjust_anchor / anchor_growing (scaffold_gaussian.py:557-651) is driven by the norm of the
    rasterizer's means2D gradient accumulated per offset.  After the same iterations through the drop-in and through the
    reference kernels, the voxels that would receive new anchors and their max-pooled features must coincide (a gradient
    within float rounding of a level's threshold may move single voxels)."""
    if not _ref_available():
        pytest.skip("oracle/_ref/libref_surfel.so did not travel")
    from train_harness import MiniScaffold2DGSTrainer, anchor_growing_candidates
    kw = dict(n_anchor=8000, k=5, W=256, H=192, seed=31, lambda_dist=0.0)
    a, b = MiniScaffold2DGSTrainer(impl="ours", **kw), MiniScaffold2DGSTrainer(impl="reference", **kw)
    a.voxel_size = b.voxel_size = 0.002
    for _ in range(3):
        a.step(); b.step()
    # thresholds scaled to this synthetic scene's gradient magnitudes so that every level selects a few hundred offsets
    thr = float((a.offset_gradient_accum / a.offset_denom.clamp(min=1)).flatten().quantile(0.9))
    ca = anchor_growing_candidates(a, grad_threshold=thr, check_interval=2, success_threshold=0.8, seed=5)
    cb = anchor_growing_candidates(b, grad_threshold=thr, check_interval=2, success_threshold=0.8, seed=5)
    total = 0
    for (xa, fa), (xb, fb) in zip(ca, cb):
        sa = {tuple(v) for v in torch.round(xa / 1e-4).long().tolist()}
        sb = {tuple(v) for v in torch.round(xb / 1e-4).long().tolist()}
        total += len(sb)
        assert len(sa ^ sb) <= max(2, 0.01 * len(sb)), (len(sa), len(sb), len(sa ^ sb))
        if xa.shape == xb.shape and torch.equal(xa, xb):
            assert torch.allclose(fa, fb, rtol=1e-4, atol=1e-6)
    assert total > 100
