"""CPU tests: the C-ABI library loads without a GPU and exports every symbol declared in
include/*.h; the drop-in Python modules mirror the reference API surface and its argument
validation (no compute call is made here)."""
import ctypes
import glob
import os
import re

import pytest
import torch

import harness as hz

INCLUDE = os.path.join(hz.ROOT, "include")


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(INCLUDE, "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(gsr_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_fn")))


def test_library_loads_and_exports_every_declared_symbol():
    import gsr_b200
    assert os.path.exists(gsr_b200.LIB_PATH), "libgsr_b200.so must be built in-tree (python -c 'import __graft_entry__ as g; g.build()')"
    L = gsr_b200.lib()
    fns = declared_functions()
    assert len(fns) >= 8
    for name in fns:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert L.gsr_abi_version() == 1
    assert L.gsr_build_arch() == b"sm_100a"


def test_missing_library_fails_loudly(monkeypatch):
    import gsr_b200._lib as lib_mod
    monkeypatch.setattr(lib_mod, "_LIB", None)
    monkeypatch.setattr(lib_mod, "LIB_PATH", "/nonexistent/libgsr_b200.so")
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        lib_mod.lib()


def test_settings_fields_match_reference_order():
    from diff_surfel_rasterization import GaussianRasterizationSettings
    # S/diff_surfel_rasterization/__init__.py:158-170
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
        "projmatrix", "sh_degree", "campos", "prefiltered", "debug")


def _settings():
    from diff_surfel_rasterization import GaussianRasterizationSettings
    z = torch.zeros(3)
    return GaussianRasterizationSettings(8, 8, 1.0, 1.0, z, 1.0, torch.eye(4), torch.eye(4), 0, z, False, False)


def test_argument_validation_like_reference():
    from diff_surfel_rasterization import GaussianRasterizer
    r = GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    # S/diff_surfel_rasterization/__init__.py:192-196
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), colors_precomp=m,
          scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), colors_precomp=m)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), colors_precomp=m, scales=torch.zeros(4, 2),
          rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 9))


def test_cpu_tensors_are_rejected_not_silently_computed():
    from diff_surfel_rasterization import GaussianRasterizer
    r = GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        r(means3D=m, means2D=m, opacities=torch.zeros(4, 1), colors_precomp=m, scales=torch.zeros(4, 2),
          rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="num_points, 3"):
        r(means3D=torch.zeros(4, 2), means2D=m, opacities=torch.zeros(4, 1), colors_precomp=m,
          scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))


def test_invalid_arguments_return_error_codes_without_gpu():
    import gsr_b200
    L = gsr_b200.lib()
    rc = L.gsr_mark_visible(5, None, None, None, None, None)
    assert rc == -1 and b"invalid" in L.gsr_last_error()
    assert L.gsr_set_option(b"nope", 1) == -1
    assert L.gsr_set_option(b"no_cull", 0) == 0
    buf = (ctypes.c_float * 16)()
    assert L.gsr_profile_read(buf) == -1  # profiling never enabled


def test_synthetic_scene_is_deterministic():
    import numpy as np
    import synth
    a, b = synth.make_scene(100, 64, 48, seed=7, sh=True), synth.make_scene(100, 64, 48, seed=7, sh=True)
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    assert np.allclose(np.linalg.norm(a.rotations, axis=1), 1, atol=1e-6)
    # full_proj = view @ proj and proj[3,2] == 1 (clip w = view z), cameras/__init__.py:85-88
    assert abs(a.cam.projmatrix[2, 3] - 1.0) < 1e-6


def test_ewa_modules_mirror_reference_api():
    """diff_gaussian_rasterization / diff_plane_rasterization / scaffold_filter / simple_knn._C surface."""
    import diff_gaussian_rasterization as dgr
    import diff_plane_rasterization as dpr
    import scaffold_filter as sf
    from simple_knn._C import distCUDA2
    base = ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
            "sh_degree", "campos", "prefiltered")
    assert dgr.GaussianRasterizationSettings._fields == base + ("debug",)          # G/...__init__.py:157-169
    assert dpr.GaussianRasterizationSettings._fields == base + ("render_geo", "debug")   # L/...__init__.py:173-186
    assert sf.GaussianRasterizationSettings._fields == base + ("debug",)           # F/...__init__.py:160-172
    z = torch.zeros(3)
    m = torch.zeros(4, 3)
    rg = dgr.GaussianRasterizer(dgr.GaussianRasterizationSettings(8, 8, 1.0, 1.0, z, 1.0, torch.eye(4), torch.eye(4), 0, z, False, False))
    rp = dpr.GaussianRasterizer(dpr.GaussianRasterizationSettings(8, 8, 1.0, 1.0, z, 1.0, torch.eye(4), torch.eye(4), 0, z, False, True, False))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        rg(means3D=m, means2D=m, opacities=torch.zeros(4, 1), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed"):
        rp(means3D=m, means2D=m, means2D_abs=m, opacities=torch.zeros(4, 1), colors_precomp=m)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        rg(means3D=m, means2D=m, opacities=torch.zeros(4, 1), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        rp(means3D=m, means2D=m, means2D_abs=m, opacities=torch.zeros(4, 1), colors_precomp=m, scales=m,
           rotations=torch.zeros(4, 4), all_map=torch.zeros(4, 5))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        sf.GaussianRasterizer(sf.GaussianRasterizationSettings(8, 8, 1.0, 1.0, z, 1.0, torch.eye(4), torch.eye(4), 0, z, False, False)).visible_filter(m, m, torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        distCUDA2(m)


def test_ewa_entry_points_reject_bad_arguments_without_gpu():
    import gsr_b200
    L = gsr_b200.lib()
    assert L.gsr_visible_filter(5, 8, 8, None, None, 1.0, None, None, None, None, 1.0, 1.0, 0, None, 0, None) == -1
    assert L.gsr_dist2_knn3(5, None, None, None, None) == -1
    assert L.gsr_dist2_knn3(0, None, None, None, None) == 0
    assert L.gsr_dist2_knn3_workspace(1000) > 1000 * 16
    assert L.gsr_gaussian_backward(0, *([0] * 3), None, 0, 0, *([None] * 4), 1.0, *([None] * 5), 1.0, 1.0, *([None] * 14), 0, None) == 0


def test_tsdf_fuse_argument_validation_without_gpu():
    import gsr_b200
    L = gsr_b200.lib()
    assert L.gsr_tsdf_fuse(-1, None, 0, None, 1.0, 0.01, 0, None, 1, None, None, None, None) == -1
    assert L.gsr_tsdf_fuse(4, None, 0, None, 1.0, 0.01, 0, None, 1, None, None, None, None) == -1   # samples required
    assert L.gsr_tsdf_fuse(0, None, 0, None, 1.0, 0.01, 0, None, 1, None, None, None, None) == 0    # nothing to do
    assert b"gsr_tsdf_fuse" in L.gsr_last_error()


def test_mesh_entry_points_argument_validation_without_gpu():
    import ctypes as C
    import gsr_b200
    L = gsr_b200.lib()
    nv, nt = C.c_longlong(0), C.c_longlong(0)
    assert L.gsr_mc_workspace_bytes(0, 4, 4) == 0
    ws = L.gsr_mc_workspace_bytes(512, 512, 512)
    assert 3.25 * 512 ** 3 < ws < 3.3 * 512 ** 3                    # 1 + 2 bytes + 2 bits per voxel + counters
    assert L.gsr_mc_count(0, 4, 4, None, None, 0.0, 0.0, None, C.byref(nv), C.byref(nt), None) == -1
    assert L.gsr_mc_count(4, 4, 4, None, None, 0.0, 0.0, None, C.byref(nv), C.byref(nt), None) == -1      # tsdf required
    assert L.gsr_mc_count(2048, 2048, 2048, None, None, 0.0, 0.0, None, C.byref(nv), C.byref(nt), None) != 0
    assert b"slabs" in L.gsr_last_error()
    assert L.gsr_mc_emit(4, 4, 4, None, None, 0.0, None, 1.0, None, None, None, None, None) == -1
    assert L.gsr_mesh_clusters(-1, 0, None, None, None, None, None, None, None) == -1
    assert L.gsr_mesh_clusters(0, 0, None, None, None, None, None, None, None) == 0                          # nothing to do
    assert L.gsr_mesh_clusters(3, 1, None, None, None, None, None, None, None) == -1
    assert L.gsr_mesh_keep_clusters(0, None, None, 50, None, None) == 0
    assert L.gsr_mesh_keep_clusters(5, None, None, 50, None, None) == -1
    assert L.gsr_mesh_filter_workspace_bytes(1000, 2000) >= 1000 * 5 + 2000
    assert L.gsr_mesh_filter_count(10, 10, None, None, None, C.byref(nv), C.byref(nt), None) == -1
    assert L.gsr_mesh_filter_emit(10, 10, None, None, None, None, None, None, None, None) == -1


def test_header_is_plain_c99():
    """The drop-in boundary is a C ABI: include/gsr_b200.h must compile as C (no C++, no torch / CUDA types) and a C
    program that links only against the shared library must resolve every declared symbol."""
    import subprocess
    import tempfile
    import gsr_b200
    src = '#include "gsr_b200.h"\n#include <stdio.h>\nint main(void) {\n  printf("%d %s\\n", gsr_abi_version(), gsr_build_arch());\n'
    src += "  typedef void (*fn)(void);\n  fn p[] = {" + ", ".join(f"(fn){n}" for n in declared_functions()) + "};\n  return p[0] == 0;\n}\n"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        libdir = os.path.dirname(gsr_b200.LIB_PATH)
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", INCLUDE, c, "-o", exe, "-L", libdir,
                        "-l:libgsr_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["1", "sm_100a"]
