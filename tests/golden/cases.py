"""Seeded case definitions shared by make_golden.py (reference CUDA on a B200) and the
tests that replay them through the CPU oracle and the product path."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402

CASES = ("colors", "sh", "ragged", "precompT", "scalemod", "q3", "allgrads", "dense")


def q3_width():
    """An image width for which the reference backward's int(focal*tan*2) truncates to W-1
    (SURVEY quirk Q3), found by replaying the float32 expression."""
    for W in range(64, 200):
        cam = synth.make_camera(W, 48)
        fx = np.float32(W) / (np.float32(2.0) * np.float32(cam.tanfovx))
        if int(np.float32(np.float32(fx * np.float32(cam.tanfovx)) * np.float32(2))) != W:
            return W
    raise RuntimeError("no Q3 width found")


def build_case(name):
    """-> (scene, dL_dcolor, dL_dothers, kwargs for run_*_surfel)"""
    kw = {}
    if name == "colors":
        sc = synth.make_scene(1500, 96, 64, seed=101, sigma_px=3.0, rotate_camera=True, bg=(0.1, 0.2, 0.3))
    elif name == "sh":
        sc = synth.make_scene(800, 96, 64, seed=102, sigma_px=3.0, sh=True, rotate_camera=True, bg=(0.3, 0.0, 0.5))
    elif name == "ragged":
        sc = synth.make_scene(1200, 90, 50, seed=103, sigma_px=2.5, rotate_camera=True, behind_fraction=0.2)
    elif name == "precompT":
        sc = synth.make_scene(1000, 96, 64, seed=104, sigma_px=3.0, bg=(0.2, 0.2, 0.2))
        from oracle.oracle import SurfelOracle
        o = SurfelOracle()
        o.forward(sc.cam, sc.means3D, sc.opacities, sc.scales, sc.rotations, colors=sc.colors)
        kw["transMat_precomp"] = o.geom()["transMat"].copy()
    elif name == "scalemod":
        sc = synth.make_scene(1000, 96, 64, seed=105, sigma_px=2.5, rotate_camera=True)
        kw["scale_modifier"] = 1.3
    elif name == "q3":
        W = q3_width()
        sc = synth.make_scene(1000, W, 48, seed=106, sigma_px=3.0, rotate_camera=True, bg=(0.5, 0.5, 0.5))
    elif name == "allgrads":
        sc = synth.make_scene(1000, 96, 64, seed=107, sigma_px=3.0, rotate_camera=True, bg=(0.1, 0.1, 0.1))
    elif name == "dense":
        # enough splats per pixel that most pixels terminate on the T < 1e-4 rule
        sc = synth.make_scene(6000, 64, 48, seed=108, sigma_px=4.0, opacity_sigma=1.0, rotate_camera=True)
    else:
        raise KeyError(name)
    W, H = sc.cam.W, sc.cam.H
    zero_from = 11 if name == "allgrads" else 7
    gc, go = synth.make_upstream_grads(W, H, seed=1000 + len(name), zero_from=zero_from)
    return sc, gc, go, kw
