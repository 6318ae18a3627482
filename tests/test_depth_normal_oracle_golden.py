"""CPU: the depth->normal oracle (oracle/depth_normal_oracle.py) against tests/golden/depth_normal_*.npz, which hold the
outputs of the reference's own normal_from_depth_image (tests/golden/make_golden_depth_normal.py)."""
import os

import numpy as np
import pytest

from depth_normal_synth import DN_CASES, build_dn_case
from oracle import depth_normal_oracle as dno

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", DN_CASES)
@pytest.mark.parametrize("weighted", [False, True])
def test_oracle_matches_reference_vectors(name, weighted):
    c = build_dn_case(name)
    g = np.load(os.path.join(GOLD, f"depth_normal_{name}.npz"))
    tag = "_w" if weighted else ""
    n, gd = dno.value_and_grad(c["depth"], c["K"], c["g"], c["weight"] if weighted else None)
    assert np.abs(n - g["normal" + tag]).max() <= 1e-6
    assert np.abs(gd - g["grad" + tag]).max() <= 1e-5 * max(np.abs(g["grad" + tag]).max(), 1e-30)
    assert np.all(n[:, 0, :] == 0) and np.all(n[:, :, -1] == 0)            # one-pixel zero border


def test_float32_error_of_the_reference_itself():
    """The reference forms x z / fx - cx z / fx in float32: its own distance from a float64 evaluation sets what parity
    between two float32 implementations can mean for this op (tolerances of tests/test_depth_normal_gpu.py)."""
    c = build_dn_case("tilted")
    g = np.load(os.path.join(GOLD, "depth_normal_tilted.npz"))
    import torch
    n64, _ = dno.value_and_grad(c["depth"], c["K"], c["g"], None, dtype=torch.float64)
    assert 1e-7 < np.abs(g["normal"] - n64).max() < 5e-4
