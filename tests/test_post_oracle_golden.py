"""CPU: oracle/surfel_post_oracle.py against the golden vectors built from the reference's own depth_to_normal."""
import os

import numpy as np
import pytest
import torch

from oracle import surfel_post_oracle as po
from post_synth import POST_CASES, build_post_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle_run(c, dtype=torch.float32):
    allmap = torch.from_numpy(c["allmap"]).to(dtype).requires_grad_(True)
    rn, sd, sn = po.postprocess(allmap, torch.from_numpy(c["wvt"]).to(dtype), torch.from_numpy(c["full_proj"]).to(dtype), c["depth_ratio"])
    g = {k: torch.from_numpy(v).to(dtype) for k, v in c["g"].items()}
    torch.autograd.backward([rn, sd, sn], [g["normal"], g["depth"], g["surf_normal"]])
    return rn.detach().numpy(), sd.detach().numpy(), sn.detach().numpy(), allmap.grad.numpy()


@pytest.mark.parametrize("name", list(POST_CASES))
def test_oracle_matches_reference_golden(name):
    c = build_post_case(name)
    gold = np.load(os.path.join(GOLD, f"post_{name}.npz"))
    rn, sd, sn, grad = oracle_run(c)
    assert np.abs(rn - gold["normal"]).max() <= 1e-6
    assert np.abs(sd - gold["depth"]).max() <= 1e-6
    assert np.abs(sn - gold["surf_normal"]).max() <= 1e-6
    assert np.array_equal(np.isnan(grad), np.isnan(gold["grad"]))          # the reference's 0/0 pixels
    m = ~np.isnan(gold["grad"])
    assert np.abs(grad[m] - gold["grad"][m]).max() <= 1e-5 * np.abs(gold["grad"][m]).max()
