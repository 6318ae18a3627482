"""GPU stress test of the decoupled record ring of the backward kernels (surfel_render_bwd.cu / ewa_render.cu).

The warps of a CTA release a ring stage through a shared-memory arrival counter and the last arriver refills it with
cp.async.bulk: no CTA barrier in the walk.  compute-sanitizer's racecheck does not model that handshake
(fence + atomic counter + fence.proxy.async), so it is checked behaviourally: gsr_set_option("dbg", 2) makes the same
kernels execute a __syncthreads() per batch -- with it the refill is trivially ordered after every warp's reads -- and the
decoupled run must reproduce that result over many launches on workloads built to provoke the race window:
  * lists of 3000-12000 entries per tile (100-380 batches of 32), walked hundreds of entries deep;
  * warps of one CTA finishing at very different list positions (the opaque centre of the heavy tile terminates early,
    its corners late), so fast warps run the full ring depth ahead of slow ones;
  * thousands of small tiles next to the heavy one, so CTAs of all kinds share an SM.
Gradients are float atomic sums, so equality is up to the summation order (<= 1e-5 of the tensor's max; a stage
overwritten while still being read yields garbage records, i.e. errors of order 1 or NaN)."""
import numpy as np
import pytest

import harness as hz
import synth

pytestmark = pytest.mark.gpu
REPS = 300


def heavy_scene(P_heavy, P_wide, W, H, seed, scale_dims=2):
    sc = synth.make_scene(P_heavy + P_wide, W, H, seed=seed, sigma_px=1.5, scale_dims=scale_dims)
    rng = np.random.default_rng(seed + 1)
    f = 1.2 * W
    z = sc.means3D[:P_heavy, 2].astype(np.float64)
    u, v = rng.uniform(20.0, 28.0, P_heavy), rng.uniform(20.0, 28.0, P_heavy)      # the middle of one heavy tile (1, 1)
    sc.means3D[:P_heavy, 0] = ((u - W / 2) * z / f).astype(np.float32)
    sc.means3D[:P_heavy, 1] = ((v - H / 2) * z / f).astype(np.float32)
    sc.opacities[:P_heavy] = rng.uniform(0.01, 0.05, (P_heavy, 1)).astype(np.float32)
    return sc


def _close(a, b, keys):
    for k in keys:
        x, y = a[k].astype(np.float64), b[k].astype(np.float64)
        assert np.isfinite(x).all(), k
        err = np.abs(x - y).max() / max(np.abs(y).max(), 1e-30)
        assert err <= 1e-5, (k, err)


@pytest.mark.parametrize("P_heavy", [3000, 12000])
def test_surfel_backward_ring_equals_barrier_ring(P_heavy):
    import gsr_b200
    W, H = 192, 128
    sc = heavy_scene(P_heavy, 20000, W, H, seed=101)
    gc, go = synth.make_upstream_grads(W, H, seed=103)
    tt = hz.to_torch(sc)
    L = gsr_b200.lib()
    L.gsr_set_option(b"dbg", 2)
    try:
        ref = hz.run_product_surfel(sc, gc, go, tt=tt)
    finally:
        L.gsr_set_option(b"dbg", 0)
    keys = ("means3D", "means2D", "colors", "opacities", "scales", "rotations")
    assert ref["others"][1][22:27, 22:27].min() > 0.9                  # the tile centre saturates early, its rim walks on
    for _ in range(REPS):
        out = hz.run_product_surfel(sc, gc, go, tt=tt)
        assert np.array_equal(out["color"], ref["color"])
        _close(out["grads"], ref["grads"], keys)


@pytest.mark.parametrize("plane", [False, True])
def test_ewa_backward_ring_equals_barrier_ring(plane):
    import gsr_b200
    W, H = 192, 128
    sc = heavy_scene(8000, 20000, W, H, seed=111, scale_dims=3)
    gc, go = synth.make_upstream_grads(W, H, seed=113, n_others=6, zero_from=6)
    kw = dict(g_color=gc, plane=plane)
    if plane:
        kw.update(all_map=synth.make_all_map(sc), render_geo=True, g_all_map=np.ascontiguousarray(go[:5]),
                  g_plane_depth=np.ascontiguousarray(go[5:6]))
    tt = hz.to_torch(sc)
    L = gsr_b200.lib()
    L.gsr_set_option(b"dbg", 2)
    try:
        ref = hz.run_product_gauss(sc, tt=tt, **kw)
    finally:
        L.gsr_set_option(b"dbg", 0)
    keys = ("means3D", "means2D", "colors", "opacities", "scales", "rotations") + (("all_map", "means2D_abs") if plane else ())
    for _ in range(REPS // 2):
        out = hz.run_product_gauss(sc, tt=tt, **kw)
        assert np.array_equal(out["color"], ref["color"])
        _close(out["grads"], ref["grads"], keys)
