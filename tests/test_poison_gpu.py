"""GPU: poisoned inputs (tests/poison_synth.py: NaN / Inf / zero / huge / negative values in a few Gaussians of a 20 k scene)
through the three rasterizers against the unmodified reference build on the same inputs: same visible set and radii, the
same pixels finite and equal, the same gradients finite and equal -- garbage in, the REFERENCE's garbage out -- and, first of
all, no fault: before this test a NaN radius in the 3DGS path left an uninitialised key in a tile list (an illegal read in
ewa_build_records), and a NaN opacity was culled where the reference's fminf() blends it with alpha 0.99.

One documented divergence: surfels with |T| >= 1e10 (world scales ~1e7 and up) are invisible here while the
reference still renders finite ones up to ~1e16 -- the record form (adjugate rows, det T) would overflow; the case checks
that nothing non-finite leaks instead."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import harness as hz  # noqa: E402
import synth  # noqa: E402
from poison_synth import KINDS, poison  # noqa: E402

W, H, P = 320, 240, 20000


def _compare(ours, ref, grad_tol, skip_mask_of=()):
    assert np.array_equal(ours["radii"], ref["radii"])
    fo, fr = np.isfinite(ours["color"]), np.isfinite(ref["color"])
    assert np.array_equal(fo, fr)
    assert np.abs(ours["color"][fo] - ref["color"][fo]).max() <= 1e-4
    for k, go in ours["grads"].items():
        gr = ref["grads"][k]
        mo, mr = np.isfinite(go), np.isfinite(gr)
        if k not in skip_mask_of:
            assert np.array_equal(mo, mr), k
        both = mo & mr
        scale = max(float(np.abs(gr[both]).max()), 1e-30) if both.any() else 1.0
        assert float(np.abs(go[both] - gr[both]).max()) / scale <= grad_tol, k


@pytest.mark.parametrize("kind", KINDS)
def test_surfel_rasterizer_on_poisoned_inputs(kind):
    from oracle import refcuda
    sc = synth.make_scene(P, W, H, seed=11)
    poison(sc, kind, np.random.default_rng(5), 2)
    gc, go = synth.make_upstream_grads(W, H, seed=12)
    ours = hz.run_product_surfel(sc, gc, go)
    torch.cuda.synchronize()
    if kind == "huge_scale":                                   # beyond the documented domain limit: contained, not equal
        assert np.isfinite(ours["color"]).all() and np.isfinite(ours["others"]).all()
        assert all(np.isfinite(g).all() for g in ours["grads"].values())
        return
    if not refcuda.available("surfel"):
        pytest.skip("oracle/_ref/libref_surfel.so did not travel")
    _compare(ours, hz.run_refcuda_surfel(sc, gc, go), 1e-3)


@pytest.mark.parametrize("plane", [False, True])
@pytest.mark.parametrize("kind", KINDS)
def test_ewa_rasterizers_on_poisoned_inputs(kind, plane):
    from oracle import refcuda
    sc = synth.make_scene(P, W, H, seed=11, scale_dims=3)
    poison(sc, kind, np.random.default_rng(5), 3)
    gc, _ = synth.make_upstream_grads(W, H, seed=12)
    kw = {}
    if plane:
        _, go2 = synth.make_upstream_grads(W, H, seed=6, n_others=6, zero_from=6)
        with np.errstate(all="ignore"):
            am = synth.make_all_map(sc)
        kw = dict(all_map=am, g_all_map=np.ascontiguousarray(go2[:5]), g_plane_depth=np.ascontiguousarray(go2[5:6]))
    ours = hz.run_product_gauss(sc, gc, plane=plane, **kw)
    torch.cuda.synchronize()                                   # no fault, whatever the inputs
    var = "plane" if plane else "gaussian"
    if not refcuda.available(var):
        pytest.skip(f"oracle/_ref/libref_{var}.so did not travel")
    # huge scales put 0.99-alpha layers over the whole frame (the reference's fminf again): gradients of everything behind
    # them are differences of nearly equal numbers -- 2e-3 there.  A zero quaternion makes the caller's all_map NaN: which
    # other Gaussians' all_map gradients that NaN reaches depends on the accumulation order, the rest must agree.
    _compare(ours, hz.run_refcuda_gauss(sc, gc, plane=plane, **kw), 2e-3 if kind == "huge_scale" else 1e-3,
             skip_mask_of=("all_map",) if kind == "zero_quat" else ())
