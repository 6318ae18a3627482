"""CPU: the numpy TSDF-fusion oracle (oracle/tsdf_oracle.py) against the golden vectors produced by the
reference's own nested functions (tests/golden/make_golden_tsdf.py)."""
import os

import numpy as np
import pytest

from oracle import tsdf_oracle
from tsdf_synth import TSDF_CASES, build_tsdf_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# float32 evaluation-order differences (BLAS matmul vs explicit sums) move sdf by ~1e-6 relative; a sample
# within that distance of a mask threshold (|ndc| = 1, sdf = -trunc) may flip one view in or out.
FLIP_FRACTION = 2e-4
TOL = 2e-5          # CPU oracle vs reference (same float32 op sequence up to the matmul's summation order)
TOL_GPU = 1e-4      # north_star forward tolerance (1e-4 of the [-1, 1] tsdf / [0, 1] colour range): the kernel
                    # contracts the projection into FMAs, which moves the bilinear sample position by ~1e-5 px


def check_against_golden(tsdf, rgb, gold, tol=TOL):
    d = np.abs(tsdf - gold["tsdf"])
    bad = d > tol
    assert bad.mean() <= FLIP_FRACTION, (bad.sum(), d.max())
    if rgb is not None:
        dr = np.abs(rgb - gold["rgb"]).max(-1)
        assert (dr > tol).mean() <= FLIP_FRACTION, ((dr > tol).sum(), dr.max())


@pytest.mark.parametrize("name", TSDF_CASES)
def test_oracle_matches_reference_golden(name):
    c = build_tsdf_case(name)
    gold = np.load(os.path.join(GOLD, f"tsdf_{name}.npz"))
    tsdf, rgb = tsdf_oracle.compute_unbounded_tsdf(c["samples"], c["contracted"], c["center"], c["radius"], c["voxel_size"],
                                                   c["projs"], c["depthmaps"], c["rgbmaps"], return_rgb=True)
    assert (gold["tsdf"] != 1).mean() > 0.03          # the case really fuses something
    check_against_golden(tsdf, rgb, gold)


def test_grid_sample_border_matches_torch():
    import torch
    rng = np.random.default_rng(0)
    img = rng.normal(size=(7, 9)).astype(np.float32)
    g = rng.uniform(-1.2, 1.2, size=(500, 2)).astype(np.float32)
    g[:4] = [[-1, -1], [1, 1], [1, -1], [0, 0]]
    ref = torch.nn.functional.grid_sample(torch.from_numpy(img)[None, None], torch.from_numpy(g)[None, None],
                                          mode="bilinear", padding_mode="border", align_corners=True).reshape(-1).numpy()
    mine = tsdf_oracle.grid_sample_border(img, g[:, 0], g[:, 1])
    assert np.abs(mine - ref).max() < 1e-6


def test_no_views_and_empty():
    c = build_tsdf_case("contracted", n=10)
    t = tsdf_oracle.compute_unbounded_tsdf(c["samples"], True, c["center"], c["radius"], c["voxel_size"], [], [], [])
    assert np.all(t == 1)
