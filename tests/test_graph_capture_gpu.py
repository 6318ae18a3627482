"""GPU tests (pytest -m gpu) of CUDA-graph capture (include/gsr_b200.h "CUDA-graph capture", gsr_b200.graphs):
a forward + backward of each rasterizer recorded into a torch.cuda.CUDAGraph replays to the eager results, a replay whose
num_rendered outgrows the captured capacity is reported (and stays memory safe), a capture without history fails
loudly, and a whole captured 2DGS training iteration (tests/train_harness.py) follows the eager iteration."""
import numpy as np
import pytest
import torch

import harness as hz
import synth

pytestmark = pytest.mark.gpu


def _surfel_setup(P=60_000, W=480, H=360, seed=3):
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    sc = synth.make_scene(P, W, H, seed=seed)
    gc, go = synth.make_upstream_grads(W, H)
    tt = hz.to_torch(sc)
    gct, got = torch.from_numpy(gc).cuda(), torch.from_numpy(go).cuda()
    rs = GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"], tt["proj"], 0,
                                       tt["campos"], False, False)
    rast = GaussianRasterizer(rs)
    leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "colors")}

    def step():
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        color, radii, others = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                    colors_precomp=leaves["colors"], scales=leaves["scales"], rotations=leaves["rotations"])
        torch.autograd.backward([color, others], [gct, got])
        return color, radii, others, m2d
    return leaves, step


def _capture(step, leaves, warmup=2):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            for v in leaves.values():
                v.grad = None
            step()
    torch.cuda.current_stream().wait_stream(side)
    for v in leaves.values():
        v.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    return g, out


def _eager(step, leaves, extra_grad_of=()):
    """One eager step; detached copies of its outputs / gradients.  Nothing of its autograd graph survives (an
    AccumulateGrad node kept alive from an eager step on the default stream would invalidate a later capture)."""
    for v in leaves.values():
        v.grad = None
    out = step()
    torch.cuda.synchronize()
    res = [o.detach().clone() for i, o in enumerate(out) if i not in extra_grad_of]
    extra = [out[i].grad.clone() for i in extra_grad_of]
    grads = {k: v.grad.clone() for k, v in leaves.items()}
    del out
    for v in leaves.values():
        v.grad = None
    return res, grads, extra


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_surfel_forward_backward_replays_to_the_eager_result():
    from gsr_b200.graphs import capture_overflow
    leaves, step = _surfel_setup()
    (color_e, radii_e, others_e), _, _ = _eager(step, leaves, extra_grad_of=(3,))
    capture_overflow(reset=True)
    g, (color, radii, others, m2d) = _capture(step, leaves)
    # new inputs in the static tensors: the replay must follow them
    with torch.no_grad():
        leaves["opacities"].mul_(0.9)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    assert capture_overflow() == 0
    rep = {k: v.grad.clone() for k, v in leaves.items()}
    rep_color, rep_radii, rep_others, rep_m2d = color.clone(), radii.clone(), others.clone(), m2d.grad.clone()
    (color_e2, radii_e2, others_e2), grads_e2, (m2d_e2,) = _eager(step, leaves, extra_grad_of=(3,))
    assert torch.equal(rep_color, color_e2) and torch.equal(rep_radii, radii_e2) and torch.equal(rep_others, others_e2)
    assert not torch.equal(color_e2, color_e)
    for k in leaves:                  # float atomics: order-dependent rounding only
        assert _rel(rep[k], grads_e2[k]) < 2e-5, k
    assert _rel(rep_m2d, m2d_e2) < 2e-5


@pytest.mark.parametrize("plane", [False, True])
def test_ewa_forward_backward_replays_to_the_eager_result(plane):
    from gsr_b200.graphs import capture_overflow
    P, W, H = 50_000, 400, 300
    sc = synth.make_scene(P, W, H, seed=5, scale_dims=3)
    gc, go = synth.make_upstream_grads(W, H, seed=6, n_others=6, zero_from=6)
    tt = hz.to_torch(sc)
    gct = torch.from_numpy(gc).cuda()
    gam, gpd = torch.from_numpy(np.ascontiguousarray(go[:5])).cuda(), torch.from_numpy(np.ascontiguousarray(go[5:6])).cuda()
    kw = dict(image_height=H, image_width=W, tanfovx=sc.cam.tanfovx, tanfovy=sc.cam.tanfovy, bg=tt["bg"], scale_modifier=1.0,
              viewmatrix=tt["view"], projmatrix=tt["proj"], sh_degree=0, campos=tt["campos"], prefiltered=False, debug=False)
    leaves = {k: tt[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "colors")}
    if plane:
        from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        kw["render_geo"] = True
        leaves["all_map"] = torch.from_numpy(synth.make_all_map(sc)).cuda().requires_grad_(True)
    else:
        from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    rast = GaussianRasterizer(GaussianRasterizationSettings(**kw))

    def step():
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        common = dict(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], colors_precomp=leaves["colors"],
                      scales=leaves["scales"], rotations=leaves["rotations"])
        if plane:
            m2a = torch.zeros_like(leaves["means3D"], requires_grad=True)
            c, r, ob, oam, pd = rast(means2D_abs=m2a, all_map=leaves["all_map"], **common)
            torch.autograd.backward([c, oam, pd], [gct, gam, gpd])
            return c, r, ob, oam, pd
        c, r = rast(**common)
        torch.autograd.backward([c], [gct])
        return c, r

    capture_overflow(reset=True)
    g, out = _capture(step, leaves)
    g.replay()
    torch.cuda.synchronize()
    assert capture_overflow() == 0
    rep_out = [o.detach().clone() for o in out]
    rep = {k: v.grad.clone() for k, v in leaves.items()}
    eager_out, grads_e, _ = _eager(step, leaves)
    assert len(rep_out) == len(eager_out)
    for a, b in zip(rep_out, eager_out):
        assert torch.equal(a, b)
    for k in leaves:
        assert _rel(rep[k], grads_e[k]) < 2e-5, k


def test_replay_that_outgrows_the_captured_capacity_is_reported():
    import gsr_b200
    from diff_surfel_rasterization import last_num_rendered
    from gsr_b200.graphs import capture_overflow
    L = gsr_b200.lib()
    leaves, step = _surfel_setup(P=40_000, W=320, H=240, seed=9)
    _eager(step, leaves)
    R = last_num_rendered()
    assert R > 4096
    try:
        L.gsr_set_option(b"force_capacity", R // 2)          # the graph is recorded with half the needed capacity
        capture_overflow(reset=True)
        g, out = _capture(step, leaves, warmup=1)
        g.replay()
        torch.cuda.synchronize()
        assert capture_overflow(reset=False) == R            # sticky ...
        assert capture_overflow(reset=True) == R
        assert capture_overflow() == 0                       # ... until reset
        assert torch.isfinite(out[0]).all()                  # truncated lists, but a well-formed frame
    finally:
        L.gsr_set_option(b"force_capacity", 0)
    (color, *_), _, _ = _eager(step, leaves)                # the eager path is unaffected
    assert torch.isfinite(color).all()


def test_capture_without_history_fails_loudly():
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    W, H = 208, 176                                          # a resolution no other test of this process uses
    sc = synth.make_scene(2_000, W, H, seed=1)
    tt = hz.to_torch(sc)
    rast = GaussianRasterizer(GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, tt["bg"], 1.0, tt["view"],
                                                            tt["proj"], 0, tt["campos"], False, False))
    args = dict(means3D=tt["means3D"], means2D=torch.zeros_like(tt["means3D"]), opacities=tt["opacities"],
                colors_precomp=tt["colors"], scales=tt["scales"], rotations=tt["rotations"])
    g = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="history"):
        with torch.cuda.graph(g):
            rast(**args)
    torch.cuda.synchronize()
    color, radii, _ = rast(**args)                           # eager call afterwards: fine, and it creates the history
    torch.cuda.synchronize()
    assert torch.isfinite(color).all() and int((radii > 0).sum()) > 0


def test_captured_training_iteration_follows_the_eager_iteration():
    """tests/train_harness.py: K iterations replayed from one graph against K eager iterations from the same
    initial state (same losses up to float-atomic noise amplified by K Adam steps)."""
    from gsr_b200.graphs import capture_overflow
    from train_harness import MiniTwoDGSTrainer
    K = 4
    kw = dict(P=20_000, W=320, H=240, seed=2)
    eager = MiniTwoDGSTrainer(impl="ours", **kw)
    eager.fused_ssim = eager.fused_post = True
    graphed = MiniTwoDGSTrainer(impl="ours", capturable=True, **kw)
    graphed.fused_ssim = graphed.fused_post = True
    warm = 2
    le = [eager.step()[0] for _ in range(warm + K)]
    capture_overflow(reset=True)
    graph, loss = graphed.capture(warmup=warm)               # the warm-up iterations train; recording the graph does not
    lg = []
    for _ in range(K):
        graph.replay()
        lg.append(float(loss))
    assert capture_overflow() == 0
    for a, b in zip(le[warm:], lg):
        assert abs(a - b) <= 2e-3 * abs(a), (le, lg)
    assert lg[-1] < le[0]                                    # and it does train
    assert torch.allclose(eager.xyz, graphed.xyz, rtol=0, atol=2e-3)
    assert int((eager.denom != graphed.denom).sum()) <= 2    # visibility counters (a borderline Gaussian may flip)
