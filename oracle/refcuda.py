"""ctypes front-end for the UNMODIFIED reference CUDA rasterizers (oracle/_ref/libref_*.so,
built by oracle/build_ref.sh from /root/reference/submodules/*) -- TEST INFRASTRUCTURE.

Used by tests/ (GPU parity against the reference itself), tests/golden/make_golden.py
(captures the vectors that pin the CPU oracle) and bench.py --impl reference.
Needs a GPU; never imported by the product path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def available(variant="surfel"):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_{variant}.so"))


def _lib(variant):
    if variant not in _LIBS:
        path = os.path.join(_HERE, "_ref", f"libref_{variant}.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run oracle/build_ref.sh where /root/reference exists")
        L = C.CDLL(path)
        L.ref_create.restype = C.c_void_p
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_num_rendered.argtypes = [C.c_void_p]
        _LIBS[variant] = L
    return _LIBS[variant]


def _p(t):
    if t is None or t.numel() == 0:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


class RefSurfel:
    """Reference diff-surfel-rasterization forward/backward on raw device pointers.
    Runs on the legacy default stream like the reference (= torch's default current stream, so it is ordered after the
    torch ops that produced its inputs; no synchronisation is added here); callers synchronise before reading results
    on the host."""

    def __init__(self):
        self.L = _lib("surfel")
        self.h = self.L.ref_create()

    def __del__(self):
        try:
            self.L.ref_free(C.c_void_p(self.h))
        except Exception:
            pass

    def forward(self, bg, view, proj, campos, W, H, tanfovx, tanfovy, means3D, opacities, scales=None,
                rotations=None, colors=None, shs=None, sh_degree=0, transMat_precomp=None,
                scale_modifier=1.0, prefiltered=False, debug=False):
        dev = means3D.device
        P = means3D.shape[0]
        M = 0 if shs is None or shs.numel() == 0 else shs.shape[1]
        color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        others = torch.zeros((11, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        self._saved = dict(P=P, M=M, D=sh_degree, W=W, H=H, bg=bg, view=view, proj=proj, campos=campos,
                           tanfovx=tanfovx, tanfovy=tanfovy, means3D=means3D, shs=shs, colors=colors,
                           scales=scales, rotations=rotations, Tpre=transMat_precomp,
                           scale_modifier=scale_modifier, radii=radii, debug=debug)
        rc = self.L.ref_forward(
            C.c_void_p(self.h), C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(bg), C.c_int(W), C.c_int(H),
            _p(means3D), _p(shs), _p(colors), _p(opacities), _p(scales), C.c_float(scale_modifier),
            _p(rotations), _p(transMat_precomp), _p(view), _p(proj), _p(campos), C.c_float(tanfovx),
            C.c_float(tanfovy), C.c_int(int(prefiltered)), _p(color), _p(others), _p(radii), C.c_int(int(debug)))
        if rc != 0:
            raise RuntimeError(f"ref_forward failed rc={rc}")
        return color, radii, others, self.L.ref_num_rendered(C.c_void_p(self.h))

    def backward(self, dL_dcolor, dL_dothers):
        s = self._saved
        P, M = s["P"], s["M"]
        dev = s["means3D"].device
        z = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=dev)  # noqa: E731
        g = dict(means2D=z(P, 3), normal=z(P, 3), opacities=z(P, 1), colors=z(P, 3), means3D=z(P, 3),
                 transMat=z(P, 9), shs=z(P, M, 3), scales=z(P, 2), rotations=z(P, 4))
        rc = self.L.ref_backward(
            C.c_void_p(self.h), C.c_int(P), C.c_int(s["D"]), C.c_int(M), _p(s["bg"]), C.c_int(s["W"]),
            C.c_int(s["H"]), _p(s["means3D"]), _p(s["shs"]), _p(s["colors"]), _p(s["scales"]),
            C.c_float(s["scale_modifier"]), _p(s["rotations"]), _p(s["Tpre"]), _p(s["view"]), _p(s["proj"]),
            _p(s["campos"]), C.c_float(s["tanfovx"]), C.c_float(s["tanfovy"]), _p(s["radii"]),
            _p(dL_dcolor.contiguous()), _p(dL_dothers.contiguous()), _p(g["means2D"]), _p(g["normal"]),
            _p(g["opacities"]), _p(g["colors"]), _p(g["means3D"]), _p(g["transMat"]), _p(g["shs"]),
            _p(g["scales"]), _p(g["rotations"]), C.c_int(int(s["debug"])))
        if rc != 0:
            raise RuntimeError(f"ref_backward failed rc={rc}")
        return g


class RefGauss:
    """Reference diff-gaussian-rasterization (plane=False) / diff-plane-rasterization (plane=True)
    forward/backward on raw device pointers (legacy default stream; callers synchronise)."""

    def __init__(self, plane=False):
        self.plane = plane
        self.L = _lib("plane" if plane else "gaussian")
        self.h = self.L.ref_create()

    def __del__(self):
        try:
            self.L.ref_free(C.c_void_p(self.h))
        except Exception:
            pass

    def forward(self, bg, view, proj, campos, W, H, tanfovx, tanfovy, means3D, opacities, scales=None,
                rotations=None, colors=None, shs=None, sh_degree=0, cov3D_precomp=None, all_map=None,
                scale_modifier=1.0, prefiltered=False, render_geo=True, debug=False):
        dev = means3D.device
        P = means3D.shape[0]
        M = 0 if shs is None or shs.numel() == 0 else shs.shape[1]
        color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        self._saved = dict(P=P, M=M, D=sh_degree, W=W, H=H, bg=bg, view=view, proj=proj, campos=campos,
                           tanfovx=tanfovx, tanfovy=tanfovy, means3D=means3D, shs=shs, colors=colors,
                           scales=scales, rotations=rotations, cov3D=cov3D_precomp, all_map=all_map,
                           scale_modifier=scale_modifier, radii=radii, debug=debug, render_geo=render_geo)
        head = [C.c_void_p(self.h), C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(bg), C.c_int(W), C.c_int(H),
                _p(means3D), _p(shs), _p(colors), _p(opacities), _p(scales), C.c_float(scale_modifier),
                _p(rotations), _p(cov3D_precomp)]
        cam = [_p(view), _p(proj), _p(campos), C.c_float(tanfovx), C.c_float(tanfovy), C.c_int(int(prefiltered))]
        out = dict(color=color, radii=radii)
        if self.plane:
            observe = torch.zeros((P,), dtype=torch.int32, device=dev)
            out_all_map = torch.zeros((5, H, W), dtype=torch.float32, device=dev)
            plane_depth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
            rc = self.L.ref_forward(*head, _p(all_map), *cam, _p(color), _p(radii), _p(observe), _p(out_all_map),
                                    _p(plane_depth), C.c_int(int(render_geo)), C.c_int(int(debug)))
            out.update(observe=observe, out_all_map=out_all_map, plane_depth=plane_depth)
            self._saved["out_all_map"] = out_all_map
        else:
            rc = self.L.ref_forward(*head, *cam, _p(color), _p(radii), C.c_int(int(debug)))
        if rc != 0:
            raise RuntimeError(f"ref_forward failed rc={rc}")
        out["num_rendered"] = self.L.ref_num_rendered(C.c_void_p(self.h))
        return out

    def backward(self, dL_dcolor, dL_dall_map=None, dL_dplane_depth=None):
        s = self._saved
        P, M = s["P"], s["M"]
        dev = s["means3D"].device
        z = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=dev)  # noqa: E731
        g = dict(means2D=z(P, 3), conic=z(P, 2, 2), opacities=z(P, 1), colors=z(P, 3), means3D=z(P, 3),
                 cov3D=z(P, 6), shs=z(P, M, 3), scales=z(P, 3), rotations=z(P, 4))
        mid = [_p(s["scales"]), C.c_float(s["scale_modifier"]), _p(s["rotations"]), _p(s["cov3D"]), _p(s["view"]),
               _p(s["proj"]), _p(s["campos"]), C.c_float(s["tanfovx"]), C.c_float(s["tanfovy"]), _p(s["radii"]),
               _p(dL_dcolor.contiguous())]
        tail = [_p(g["conic"]), _p(g["opacities"]), _p(g["colors"]), _p(g["means3D"]), _p(g["cov3D"]), _p(g["shs"]),
                _p(g["scales"]), _p(g["rotations"])]
        if self.plane:
            g["means2D_abs"] = z(P, 3)
            g["all_map"] = z(P, 5)
            H, W = s["H"], s["W"]
            dam = dL_dall_map.contiguous() if dL_dall_map is not None else z(5, H, W)
            dpd = dL_dplane_depth.contiguous() if dL_dplane_depth is not None else z(1, H, W)
            rc = self.L.ref_backward(
                C.c_void_p(self.h), C.c_int(P), C.c_int(s["D"]), C.c_int(M), _p(s["bg"]), _p(s["out_all_map"]),
                C.c_int(W), C.c_int(H), _p(s["means3D"]), _p(s["shs"]), _p(s["colors"]), _p(s["all_map"]), *mid,
                _p(dam), _p(dpd), _p(g["means2D"]), _p(g["means2D_abs"]), *tail, _p(g["all_map"]),
                C.c_int(int(s["render_geo"])), C.c_int(int(s["debug"])))
        else:
            rc = self.L.ref_backward(
                C.c_void_p(self.h), C.c_int(P), C.c_int(s["D"]), C.c_int(M), _p(s["bg"]), C.c_int(s["W"]),
                C.c_int(s["H"]), _p(s["means3D"]), _p(s["shs"]), _p(s["colors"]), *mid, _p(g["means2D"]), *tail,
                C.c_int(int(s["debug"])))
        if rc != 0:
            raise RuntimeError(f"ref_backward failed rc={rc}")
        return g


def ref_visible_filter(means3D, scales, rotations, view, proj, W, H, tanfovx, tanfovy, scale_modifier=1.0,
                       cov3D_precomp=None, prefiltered=False):
    """Reference scaffold_filter visible_filter (F/rasterize_points.cu:220-284) -> radii (P,) int32."""
    path = os.path.join(_HERE, "_ref", "libref_filter.so")
    L = _LIBS.setdefault("filter", C.CDLL(path))
    P = means3D.shape[0]
    radii = torch.zeros((P,), dtype=torch.int32, device=means3D.device)
    # no synchronisation: the reference kernels launch on the legacy default stream, which is torch's default stream
    rc = L.ref_visible_filter(C.c_int(P), C.c_int(W), C.c_int(H), _p(means3D), _p(scales), C.c_float(scale_modifier),
                              _p(rotations), _p(cov3D_precomp), _p(view), _p(proj), C.c_float(tanfovx),
                              C.c_float(tanfovy), C.c_int(int(prefiltered)), _p(radii), C.c_int(0))
    if rc != 0:
        raise RuntimeError(f"ref_visible_filter failed rc={rc}")
    return radii


def ref_dist2_knn3(points):
    """Reference simple_knn distCUDA2 (K/spatial.cu:14-25) -> (P,) float32."""
    path = os.path.join(_HERE, "_ref", "libref_knn.so")
    L = _LIBS.setdefault("knn", C.CDLL(path))
    P = points.shape[0]
    out = torch.zeros((P,), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    rc = L.ref_dist2_knn3(C.c_int(P), _p(points.contiguous()), _p(out))
    if rc != 0:
        raise RuntimeError(f"ref_dist2_knn3 failed rc={rc}")
    return out
