"""ctypes loader for libgsr_b200.so (C ABI in include/gsr_b200.h)."""
from __future__ import annotations

import ctypes as C
import itertools
import os

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# GSR_B200_LIB: developer hook to A/B another build of the same library (never a fallback: it must exist)
LIB_PATH = os.environ.get("GSR_B200_LIB") or os.path.join(_PKG_ROOT, "libgsr_b200.so")

BUFFER_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)
_fp = C.c_void_p  # device pointers travel as integers

_LIB = None

_ERRORS = {-1: "invalid argument", -2: "CUDA error", -3: "buffer callback failed",
           -4: "prefiltered violation", -5: "num_rendered overflow"}


def lib():
    """Load libgsr_b200.so; fail loudly when it is absent (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libgsr_b200.so not found at {LIB_PATH}: build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C gs-sr_b200/csrc`. "
            "gsr_b200 has no CPU/PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    L.gsr_last_error.restype = C.c_char_p
    L.gsr_build_arch.restype = C.c_char_p
    L.gsr_abi_version.restype = C.c_int
    L.gsr_surfel_forward.restype = C.c_int
    L.gsr_surfel_forward.argtypes = (
        [BUFFER_FN, BUFFER_FN, BUFFER_FN, C.c_void_p, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_int]
        + [_fp] * 5 + [C.c_float] + [_fp] * 5 + [C.c_float, C.c_float, C.c_int] + [_fp] * 3 + [C.c_int, C.c_void_p])
    L.gsr_surfel_backward.restype = C.c_int
    L.gsr_surfel_backward.argtypes = (
        [C.c_int] * 4 + [_fp, C.c_int, C.c_int] + [_fp] * 4 + [C.c_float] + [_fp] * 5 + [C.c_float, C.c_float]
        + [_fp] * 15 + [C.c_int, C.c_void_p])
    L.gsr_mark_visible.restype = C.c_int
    L.gsr_mark_visible.argtypes = [C.c_int, _fp, _fp, _fp, _fp, C.c_void_p]
    head = [BUFFER_FN, BUFFER_FN, BUFFER_FN, C.c_void_p, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_int]
    L.gsr_gaussian_forward.restype = C.c_int
    L.gsr_gaussian_forward.argtypes = (head + [_fp] * 5 + [C.c_float] + [_fp] * 5 + [C.c_float, C.c_float, C.c_int]
                                       + [_fp] * 2 + [C.c_int, C.c_void_p])
    L.gsr_gaussian_backward.restype = C.c_int
    L.gsr_gaussian_backward.argtypes = (
        [C.c_int] * 4 + [_fp, C.c_int, C.c_int] + [_fp] * 4 + [C.c_float] + [_fp] * 5 + [C.c_float, C.c_float]
        + [_fp] * 14 + [C.c_int, C.c_void_p])
    L.gsr_plane_forward.restype = C.c_int
    L.gsr_plane_forward.argtypes = (head + [_fp] * 5 + [C.c_float] + [_fp] * 6 + [C.c_float, C.c_float, C.c_int]
                                    + [_fp] * 5 + [C.c_int, C.c_int, C.c_void_p])
    L.gsr_plane_backward.restype = C.c_int
    L.gsr_plane_backward.argtypes = (
        [C.c_int] * 4 + [_fp, _fp, C.c_int, C.c_int] + [_fp] * 5 + [C.c_float] + [_fp] * 5 + [C.c_float, C.c_float]
        + [_fp] * 18 + [C.c_int, C.c_int, C.c_void_p])
    L.gsr_visible_filter.restype = C.c_int
    L.gsr_visible_filter.argtypes = ([C.c_int] * 3 + [_fp, _fp, C.c_float] + [_fp] * 4 + [C.c_float, C.c_float, C.c_int, _fp,
                                                                                    C.c_int, C.c_void_p])
    L.gsr_dist2_knn3_workspace.restype = C.c_size_t
    L.gsr_dist2_knn3_workspace.argtypes = [C.c_int]
    L.gsr_dist2_knn3.restype = C.c_int
    L.gsr_dist2_knn3.argtypes = [C.c_int, _fp, _fp, _fp, C.c_void_p]
    L.gsr_tsdf_fuse.restype = C.c_int
    L.gsr_tsdf_fuse.argtypes = [C.c_longlong, _fp, C.c_int, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int, _fp,
                                C.c_int, _fp, _fp, _fp, C.c_void_p]
    L.gsr_ssim_tile_count.restype = C.c_size_t
    L.gsr_ssim_tile_count.argtypes = [C.c_int] * 3
    L.gsr_ssim_forward.restype = C.c_int
    L.gsr_ssim_forward.argtypes = [C.c_int] * 3 + [_fp, _fp, C.POINTER(C.c_float)] + [_fp] * 4 + [C.c_void_p]
    L.gsr_ssim_backward.restype = C.c_int
    L.gsr_ssim_backward.argtypes = [C.c_int] * 3 + [_fp, _fp, C.POINTER(C.c_float)] + [_fp] * 5 + [C.c_void_p]
    L.gsr_surfel_post_forward.restype = C.c_int
    L.gsr_surfel_post_forward.argtypes = [C.c_int, C.c_int, _fp, _fp, C.c_float, _fp, _fp, _fp, C.c_void_p]
    L.gsr_surfel_post_backward.restype = C.c_int
    L.gsr_surfel_post_backward.argtypes = [C.c_int, C.c_int, _fp, _fp, C.c_float] + [_fp] * 6 + [C.c_void_p]
    if hasattr(L, "gsr_tsdf_integrate_grid"):
        L.gsr_tsdf_integrate_grid.restype = C.c_int
        L.gsr_tsdf_integrate_grid.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_int, _fp,
                                                              C.c_int, _fp, _fp, _fp, C.c_void_p]
    if hasattr(L, "gsr_mc_count"):
        L.gsr_mc_workspace_bytes.restype = C.c_size_t
        L.gsr_mc_workspace_bytes.argtypes = [C.c_int] * 3
        L.gsr_mc_count.restype = C.c_int
        L.gsr_mc_count.argtypes = [C.c_int] * 3 + [_fp, _fp, C.c_float, C.c_float, _fp, C.POINTER(C.c_longlong),
                                                   C.POINTER(C.c_longlong), C.c_void_p]
        L.gsr_mc_emit.restype = C.c_int
        L.gsr_mc_emit.argtypes = [C.c_int] * 3 + [_fp, _fp, C.c_float, C.POINTER(C.c_float), C.c_float, _fp, _fp, _fp, _fp,
                                                  C.c_void_p]
    if hasattr(L, "gsr_mesh_clusters"):
        _ll = C.c_longlong
        L.gsr_mesh_clusters.restype = C.c_int
        L.gsr_mesh_clusters.argtypes = [_ll, _ll] + [_fp] * 6 + [C.c_void_p]
        L.gsr_mesh_cluster_sizes.restype = C.c_int
        L.gsr_mesh_cluster_sizes.argtypes = [_ll, _fp, _fp, _ll, C.POINTER(_ll), C.c_void_p]
        L.gsr_mesh_keep_clusters.restype = C.c_int
        L.gsr_mesh_keep_clusters.argtypes = [_ll, _fp, _fp, C.c_uint, _fp, C.c_void_p]
        L.gsr_mesh_filter_workspace_bytes.restype = C.c_size_t
        L.gsr_mesh_filter_workspace_bytes.argtypes = [_ll, _ll]
        L.gsr_mesh_filter_count.restype = C.c_int
        L.gsr_mesh_filter_count.argtypes = [_ll, _ll, _fp, _fp, _fp, C.POINTER(_ll), C.POINTER(_ll), C.c_void_p]
        L.gsr_mesh_filter_emit.restype = C.c_int
        L.gsr_mesh_filter_emit.argtypes = [_ll, _ll] + [_fp] * 7 + [C.c_void_p]
    if hasattr(L, "gsr_depth_normal_forward"):
        L.gsr_depth_normal_forward.restype = C.c_int
        L.gsr_depth_normal_forward.argtypes = [C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_void_p]
        L.gsr_depth_normal_backward.restype = C.c_int
        L.gsr_depth_normal_backward.argtypes = [C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_void_p]
    if hasattr(L, "gsr_surfel_audit"):      # absent only from older builds loaded through the GSR_B200_LIB developer hook
        L.gsr_last_num_rendered.restype = C.c_int
        L.gsr_last_num_rendered.argtypes = []
        L.gsr_surfel_audit.restype = C.c_int
        L.gsr_surfel_audit.argtypes = [C.c_int] * 4 + [_fp] * 5 + [C.c_void_p]
        L.gsr_ewa_audit.restype = C.c_int
        L.gsr_ewa_audit.argtypes = [C.c_int] * 5 + [_fp] * 5 + [C.c_void_p]
    if hasattr(L, "gsr_capture_overflow"):
        L.gsr_capture_overflow.restype = C.c_uint
        L.gsr_capture_overflow.argtypes = [C.c_int]
    _LIB = L
    return L


def check(rc, what):
    if rc < 0:
        msg = lib().gsr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({_ERRORS.get(rc, rc)}): {msg}")
    return rc


def ptr(t):
    """Device pointer of a tensor, None (-> NULL) for empty / absent tensors, as the
    reference glue turns empty tensors into nullptr."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


# The three resize callbacks of the reference protocol (S/rasterize_points.cu:31-37) are created
# ONCE per process (building a ctypes thunk costs more than the allocation it performs); the
# `user` pointer carries the id of the TorchBuffers instance that serves the call.
_ACTIVE = {}
_NEXT_ID = itertools.count(1)      # next() is atomic under the GIL: two host threads never share an id


def _make_cb(name):
    def cb(user, nbytes):
        try:
            return _ACTIVE[user].alloc(name, int(nbytes))
        except Exception:  # surfaced as GSR_E_ALLOC by the library
            return 0
    return BUFFER_FN(cb)


GEOM_FN, BINNING_FN, IMAGE_FN = _make_cb("geom"), _make_cb("binning"), _make_cb("image")


class TorchBuffers:
    """Scratch owner for one forward call, backed by torch's caching allocator on `device`."""

    geom_fn, binning_fn, image_fn = GEOM_FN, BINNING_FN, IMAGE_FN

    def __init__(self, device):
        import torch
        self._torch = torch
        self.device = device
        self.tensors = {}
        self.user = next(_NEXT_ID)

    def __enter__(self):
        _ACTIVE[self.user] = self
        return self

    def __exit__(self, *exc):
        _ACTIVE.pop(self.user, None)
        return False

    def alloc(self, name, nbytes):
        t = self._torch.empty(nbytes + 256, dtype=self._torch.uint8, device=self.device)
        self.tensors[name] = t
        return t.data_ptr()

    def get(self, name):
        t = self.tensors.get(name)
        if t is None:
            t = self._torch.empty(0, dtype=self._torch.uint8, device=self.device)
        return t
