"""ctypes front-end for the CPU oracle (oracle/liborc.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product path (gs-sr_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def _lib(double=False):
    name = "liborc64.so" if double else "liborc.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.orc_surfel_forward.restype = C.c_void_p
        lib.orc_surfel_num_rendered.restype = C.c_int64
        lib.orc_surfel_k_eval.restype = C.c_int64
        lib.orc_surfel_k_blend.restype = C.c_int64
        for fn in ("orc_surfel_num_rendered", "orc_surfel_k_eval", "orc_surfel_k_blend", "orc_surfel_free"):
            getattr(lib, fn).argtypes = [C.c_void_p]
        _LIBS[name] = lib
    return _LIBS[name]


def _fp(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_float_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class SurfelOracle:
    """Forward/backward of the 2DGS surfel rasterizer on the CPU (oracle/surfel_oracle.c)."""

    def __init__(self, double=False):
        self.lib = _lib(double)
        self.h = None

    def close(self):
        if self.h:
            self.lib.orc_surfel_free(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, cam, means3D, opacities, scales=None, rotations=None, colors=None,
                shs=None, sh_degree=0, transMat_precomp=None, scale_modifier=1.0,
                prefiltered=False, tile_stride=1):
        self.close()
        P = means3D.shape[0]
        W, H = cam.W, cam.H
        self._in = dict(means3D=_f32(means3D), shs=_f32(shs), colors=_f32(colors),
                        opac=_f32(opacities), scales=_f32(scales), rot=_f32(rotations),
                        Tpre=_f32(transMat_precomp))
        i = self._in
        M = 0 if shs is None else shs.shape[1]
        color = np.zeros((3, H, W), np.float32)
        others = np.zeros((11, H, W), np.float32)
        radii = np.zeros((P,), np.int32)
        err = C.c_int(0)
        self.P, self.M, self.W, self.H = P, M, W, H
        h = self.lib.orc_surfel_forward(
            C.c_int(P), C.c_int(sh_degree), C.c_int(M), _fp(_f32(cam.bg)), C.c_int(W), C.c_int(H),
            _fp(i["means3D"]), _fp(i["shs"]), _fp(i["colors"]), _fp(i["opac"]), _fp(i["scales"]),
            C.c_float(scale_modifier), _fp(i["rot"]), _fp(i["Tpre"]), _fp(_f32(cam.viewmatrix)),
            _fp(_f32(cam.projmatrix)), _fp(_f32(cam.campos)), C.c_float(cam.tanfovx),
            C.c_float(cam.tanfovy), C.c_int(int(prefiltered)), C.c_int(tile_stride), _fp(color),
            _fp(others), radii.ctypes.data_as(c_int_p), C.byref(err))
        self.h = h
        if err.value == 1:
            raise RuntimeError("prefiltered trap: a point was culled although prefiltered is set")
        return dict(color=color, others=others, radii=radii,
                    num_rendered=int(self.lib.orc_surfel_num_rendered(C.c_void_p(h))),
                    k_eval=int(self.lib.orc_surfel_k_eval(C.c_void_p(h))),
                    k_blend=int(self.lib.orc_surfel_k_blend(C.c_void_p(h))))

    def geom(self):
        P = self.P
        d = dict(depths=np.zeros(P, np.float32), xy=np.zeros((P, 2), np.float32),
                 transMat=np.zeros((P, 9), np.float32), normal_opacity=np.zeros((P, 4), np.float32),
                 rgb=np.zeros((P, 3), np.float32), tiles_touched=np.zeros(P, np.int32))
        self.lib.orc_surfel_get_geom(C.c_void_p(self.h), _fp(d["depths"]), _fp(d["xy"]), _fp(d["transMat"]),
                                     _fp(d["normal_opacity"]), _fp(d["rgb"]),
                                     d["tiles_touched"].ctypes.data_as(c_int_p))
        return d

    def binning(self):
        R = int(self.lib.orc_surfel_num_rendered(C.c_void_p(self.h)))
        ntiles = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        pl = np.zeros(max(R, 1), np.uint32)
        rg = np.zeros((ntiles, 2), np.uint32)
        self.lib.orc_surfel_get_binning(C.c_void_p(self.h), pl.ctypes.data_as(C.c_void_p),
                                        rg.ctypes.data_as(C.c_void_p))
        return pl[:R], rg

    def image_state(self):
        N = self.W * self.H
        ft = np.zeros((3, N), np.float32)
        nc = np.zeros((2, N), np.uint32)
        self.lib.orc_surfel_get_image_state(C.c_void_p(self.h), _fp(ft), nc.ctypes.data_as(C.c_void_p))
        return ft, nc

    def backward(self, dL_dcolor, dL_dothers, tile_stride=1):
        P, M = self.P, self.M
        i = self._in
        g = dict(means2D=np.zeros((P, 3), np.float32), colors=np.zeros((P, 3), np.float32),
                 opacities=np.zeros((P, 1), np.float32), means3D=np.zeros((P, 3), np.float32),
                 transMat=np.zeros((P, 9), np.float32), shs=np.zeros((P, M, 3), np.float32),
                 scales=np.zeros((P, 2), np.float32), rotations=np.zeros((P, 4), np.float32),
                 normal=np.zeros((P, 3), np.float32), means2D_raw=np.zeros((P, 2), np.float32))
        dpix, doth = _f32(dL_dcolor), _f32(dL_dothers)
        self.lib.orc_surfel_backward(
            C.c_void_p(self.h), _fp(i["means3D"]), _fp(i["shs"]), _fp(i["colors"]), _fp(i["scales"]),
            _fp(i["rot"]), _fp(i["Tpre"]), _fp(dpix), _fp(doth), C.c_int(tile_stride),
            _fp(g["means2D"]), _fp(g["colors"]), _fp(g["opacities"]), _fp(g["means3D"]),
            _fp(g["transMat"]), _fp(g["shs"]), _fp(g["scales"]), _fp(g["rotations"]),
            _fp(g["normal"]), _fp(g["means2D_raw"]))
        return g


def mark_visible(means3D, viewmatrix, double=False):
    P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    _lib(double).orc_mark_visible(C.c_int(P), _fp(_f32(means3D)), _fp(_f32(viewmatrix)),
                                  out.ctypes.data_as(C.c_void_p))
    return out.astype(bool)


def _glib(double=False):
    lib = _lib(double)
    if not getattr(lib, "_gauss_ready", False):
        lib.orc_gauss_forward.restype = C.c_void_p
        lib.orc_gauss_num_rendered.restype = C.c_int64
        lib.orc_gauss_num_rendered.argtypes = [C.c_void_p]
        lib.orc_gauss_free.argtypes = [C.c_void_p]
        lib._gauss_ready = True
    return lib


class GaussOracle:
    """Forward/backward of the 3DGS (plane=False) / PGSR plane (plane=True) rasterizer on the CPU
    (oracle/gauss_oracle.c)."""

    def __init__(self, plane=False, double=False):
        self.lib = _glib(double)
        self.plane = plane
        self.h = None

    def close(self):
        if self.h:
            self.lib.orc_gauss_free(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, cam, means3D, opacities, scales=None, rotations=None, colors=None, shs=None,
                sh_degree=0, cov3D_precomp=None, all_map=None, scale_modifier=1.0, prefiltered=False,
                render_geo=True):
        self.close()
        P = means3D.shape[0]
        W, H = cam.W, cam.H
        self._in = dict(means3D=_f32(means3D), shs=_f32(shs), colors=_f32(colors), opac=_f32(opacities),
                        scales=_f32(scales), rot=_f32(rotations), cov=_f32(cov3D_precomp), all_map=_f32(all_map))
        i = self._in
        M = 0 if shs is None else shs.shape[1]
        self.P, self.M, self.W, self.H = P, M, W, H
        color = np.zeros((3, H, W), np.float32)
        radii = np.zeros((P,), np.int32)
        observe = np.zeros((P,), np.int32)
        out_all_map = np.zeros((5, H, W), np.float32)
        plane_depth = np.zeros((1, H, W), np.float32)
        err = C.c_int(0)
        geo = bool(self.plane and render_geo)
        h = self.lib.orc_gauss_forward(
            C.c_int(P), C.c_int(sh_degree), C.c_int(M), _fp(_f32(cam.bg)), C.c_int(W), C.c_int(H),
            _fp(i["means3D"]), _fp(i["shs"]), _fp(i["colors"]), _fp(i["opac"]), _fp(i["scales"]),
            C.c_float(scale_modifier), _fp(i["rot"]), _fp(i["cov"]), _fp(i["all_map"]),
            _fp(_f32(cam.viewmatrix)), _fp(_f32(cam.projmatrix)), _fp(_f32(cam.campos)), C.c_float(cam.tanfovx),
            C.c_float(cam.tanfovy), C.c_int(int(prefiltered)), C.c_int(int(self.plane)), C.c_int(int(geo)),
            _fp(color), radii.ctypes.data_as(c_int_p), observe.ctypes.data_as(c_int_p), _fp(out_all_map),
            _fp(plane_depth), C.byref(err))
        self.h = h
        if err.value == 1:
            raise RuntimeError("prefiltered trap: a point was culled although prefiltered is set")
        out = dict(color=color, radii=radii, num_rendered=int(self.lib.orc_gauss_num_rendered(C.c_void_p(h))))
        if self.plane:
            out.update(observe=observe, out_all_map=out_all_map, plane_depth=plane_depth)
        return out

    def geom(self):
        P = self.P
        d = dict(depths=np.zeros(P, np.float32), xy=np.zeros((P, 2), np.float32), cov3D=np.zeros((P, 6), np.float32),
                 conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32))
        self.lib.orc_gauss_get_geom(C.c_void_p(self.h), _fp(d["depths"]), _fp(d["xy"]), _fp(d["cov3D"]),
                                    _fp(d["conic_opacity"]), _fp(d["rgb"]))
        return d

    def backward(self, dL_dcolor, dL_dall_map=None, dL_dplane_depth=None):
        P, M, W, H = self.P, self.M, self.W, self.H
        i = self._in
        g = dict(means2D=np.zeros((P, 3), np.float32), means2D_abs=np.zeros((P, 3), np.float32),
                 conic=np.zeros((P, 2, 2), np.float32), opacities=np.zeros((P, 1), np.float32),
                 colors=np.zeros((P, 3), np.float32), means3D=np.zeros((P, 3), np.float32),
                 cov3D=np.zeros((P, 6), np.float32), shs=np.zeros((P, M, 3), np.float32),
                 scales=np.zeros((P, 3), np.float32), rotations=np.zeros((P, 4), np.float32),
                 all_map=np.zeros((P, 5), np.float32))
        dam = _f32(dL_dall_map) if dL_dall_map is not None else np.zeros((5, H, W), np.float32)
        dpd = _f32(dL_dplane_depth) if dL_dplane_depth is not None else np.zeros((1, H, W), np.float32)
        self.lib.orc_gauss_backward(
            C.c_void_p(self.h), _fp(i["means3D"]), _fp(i["shs"]), _fp(i["colors"]), _fp(i["all_map"]),
            _fp(i["scales"]), _fp(i["rot"]), _fp(i["cov"]), _fp(_f32(dL_dcolor)), _fp(dam), _fp(dpd),
            _fp(g["means2D"]), _fp(g["means2D_abs"]), _fp(g["conic"]), _fp(g["opacities"]), _fp(g["colors"]),
            _fp(g["means3D"]), _fp(g["cov3D"]), _fp(g["shs"]), _fp(g["scales"]), _fp(g["rotations"]),
            _fp(g["all_map"]))
        if not self.plane:
            g.pop("means2D_abs"); g.pop("all_map")
        return g


def visible_filter(cam, means3D, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0,
                   prefiltered=False, double=False):
    """scaffold_filter.visible_filter on the CPU (oracle/gauss_oracle.c: orc_visible_filter) -> radii."""
    P = means3D.shape[0]
    radii = np.zeros((P,), np.int32)
    trap = _glib(double).orc_visible_filter(
        C.c_int(P), C.c_int(cam.W), C.c_int(cam.H), _fp(_f32(means3D)), _fp(_f32(scales)), C.c_float(scale_modifier),
        _fp(_f32(rotations)), _fp(_f32(cov3D_precomp)), _fp(_f32(cam.viewmatrix)), _fp(_f32(cam.projmatrix)),
        C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), C.c_int(int(prefiltered)), radii.ctypes.data_as(c_int_p))
    if trap:
        raise RuntimeError("prefiltered trap: a point was culled although prefiltered is set")
    return radii


def dist2_knn3(points):
    """simple_knn.distCUDA2 on the CPU (exhaustive; oracle/gauss_oracle.c: orc_dist2_knn3)."""
    pts = _f32(points)
    out = np.zeros((pts.shape[0],), np.float32)
    _lib(False).orc_dist2_knn3(C.c_int(pts.shape[0]), _fp(pts), _fp(out))
    return out
