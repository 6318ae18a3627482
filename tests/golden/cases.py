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


# ---- 3DGS / plane / filter / knn cases ------------------------------------------------------------
GAUSS_CASES = ("g_colors", "g_sh", "g_cov", "g_scalemod", "g_ragged", "g_dense", "p_geo", "p_nogeo", "p_sh")
FILTER_CASES = ("f_basic", "f_ragged")
KNN_CASES = ("k_uniform", "k_clustered", "k_tiny")


def build_gauss_case(name):
    """-> (scene, kwargs for run_*_gauss incl. upstream grads)"""
    kw = {}
    plane = name.startswith("p_")
    if name == "g_colors":
        sc = synth.make_scene(1500, 96, 64, seed=201, sigma_px=3.0, rotate_camera=True, bg=(0.1, 0.2, 0.3), scale_dims=3)
    elif name == "g_sh":
        sc = synth.make_scene(800, 96, 64, seed=202, sigma_px=3.0, sh=True, rotate_camera=True, bg=(0.3, 0.0, 0.5), scale_dims=3)
    elif name == "g_cov":
        sc = synth.make_scene(1000, 96, 64, seed=203, sigma_px=3.0, bg=(0.2, 0.2, 0.2), scale_dims=3)
        from oracle.oracle import GaussOracle
        o = GaussOracle()
        o.forward(sc.cam, sc.means3D, sc.opacities, sc.scales, sc.rotations, colors=sc.colors)
        kw["cov3D_precomp"] = o.geom()["cov3D"].copy()
    elif name == "g_scalemod":
        sc = synth.make_scene(1000, 96, 64, seed=204, sigma_px=2.5, rotate_camera=True, scale_dims=3)
        kw["scale_modifier"] = 1.3
    elif name == "g_ragged":
        # wide field: part of the scene beyond the 1.3 tanfov clamp, points behind the camera, ragged image
        sc = synth.make_scene(1200, 90, 50, seed=205, sigma_px=2.5, rotate_camera=True, behind_fraction=0.2, scale_dims=3)
        sc.means3D[:, :2] *= 1.5
    elif name == "g_dense":
        sc = synth.make_scene(6000, 64, 48, seed=206, sigma_px=4.0, opacity_sigma=1.0, rotate_camera=True, scale_dims=3)
    elif name == "p_geo":
        sc = synth.make_scene(1500, 96, 64, seed=207, sigma_px=3.0, rotate_camera=True, bg=(0.1, 0.2, 0.3), scale_dims=3)
    elif name == "p_nogeo":
        sc = synth.make_scene(1000, 96, 64, seed=208, sigma_px=3.0, rotate_camera=True, scale_dims=3)
        kw["render_geo"] = False
    elif name == "p_sh":
        sc = synth.make_scene(800, 90, 50, seed=209, sigma_px=3.0, sh=True, rotate_camera=True, behind_fraction=0.1, scale_dims=3)
    else:
        raise KeyError(name)
    W, H = sc.cam.W, sc.cam.H
    gc, go = synth.make_upstream_grads(W, H, seed=2000 + len(name), n_others=6, zero_from=6)
    kw["g_color"] = gc
    kw["plane"] = plane
    if plane:
        kw["all_map"] = synth.make_all_map(sc)
        if kw.get("render_geo", True):
            kw["g_all_map"] = np.ascontiguousarray(go[:5])
            kw["g_plane_depth"] = np.ascontiguousarray(go[5:6])
    return sc, kw


def build_filter_case(name):
    if name == "f_basic":
        return synth.make_scene(4000, 160, 96, seed=301, sigma_px=3.0, rotate_camera=True, scale_dims=3), {}
    if name == "f_ragged":
        sc = synth.make_scene(4000, 90, 50, seed=302, sigma_px=0.5, rotate_camera=True, behind_fraction=0.3, scale_dims=3)
        sc.means3D[:, :2] *= 2.0   # many anchors off-screen: rect area 0
        return sc, {"scale_modifier": 0.7}
    raise KeyError(name)


def build_knn_case(name):
    if name == "k_uniform":
        return synth.make_points(5000, seed=401)
    if name == "k_clustered":
        return synth.make_points(7000, seed=402, clustered=True)
    if name == "k_tiny":
        return synth.make_points(5, seed=403)
    raise KeyError(name)
