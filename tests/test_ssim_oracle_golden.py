"""CPU: oracle/ssim_oracle.py against the golden vectors produced by the reference's own VanillaScene.ssim."""
import os

import numpy as np
import pytest

from oracle import ssim_oracle
from ssim_synth import SSIM_CASES, build_ssim_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", list(SSIM_CASES))
def test_oracle_matches_reference_golden(name):
    a, b = build_ssim_case(name)
    gold = np.load(os.path.join(GOLD, f"ssim_{name}.npz"))
    v, g = ssim_oracle.ssim_value_and_grad(a, b)
    assert abs(v - float(gold["value"])) <= 1e-6
    assert np.abs(g - gold["grad"]).max() <= 1e-6 * max(1.0, np.abs(gold["grad"]).max() * 1e3)
    assert np.abs(g - gold["grad"]).max() <= 2e-5 * np.abs(gold["grad"]).max()
