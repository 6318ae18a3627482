"""Multi-view TSDF / colour fusion on sample points -- host-side mirror of the closure that
``GaussianExtractor.extract_mesh_unbounded`` builds (/root/reference/gssr/utils/mesh_utils.py:181-277)
and hands to ``marching_cubes_with_contraction`` as ``sdf`` (gssr/utils/mcube_utils.py:57-68).

    fusion = TSDFFusion(viewpoint_stack_full_proj, depthmaps, rgbmaps, center, radius)
    sdf_function = lambda x: fusion.compute_unbounded_tsdf(x, inv_contraction=True, voxel_size=vs)
    _, rgbs = fusion.compute_unbounded_tsdf(vertices, inv_contraction=None, voxel_size=vs, return_rgb=True)

Same argument names and meaning as the reference's nested ``compute_unbounded_tsdf`` (:209-246); the only
difference is that ``inv_contraction`` is a flag (anything but None selects the reference's
``unnormalize(uncontract(x))`` with this object's center / radius, :187-193, :248-250) because the
contraction runs inside the CUDA kernel.  Views are fused many per kernel launch (``gsr_tsdf_fuse``,
gs-sr_b200/csrc/tsdf.cu; 17 depth-only / 8 depth+RGB 1600x1060 views per launch); no CPU / PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import struct

import torch

from . import check, lib
from ._torch_util import f32c, on_device, stream_ptr

_VIEW_BYTES = 96


class TSDFFusion:
    # Views fused per kernel launch are bounded so that the maps one launch gathers from stay mostly resident in
    # the 126 MB L2 (all sample blocks of a launch walk the same maps); the running tsdf / weight / rgb state is
    # carried between launches through the C ABI's init=0 continuation (8..20 B/sample per launch).  Budgets are
    # empirical (B200, 256^3 samples, 1600x1060 maps: 16 depth-only views or 8 depth+RGB views per launch are
    # 5 % / 35 % faster than all 32 at once; smaller groups lose to the per-launch re-read of the state).
    MAP_BUDGET_BYTES = {1: 112 << 20, 4: 224 << 20}

    def __init__(self, full_proj_transforms, depthmaps, rgbmaps=None, center=None, radius=1.0, device=None,
                 views_per_launch=None):
        """full_proj_transforms: sequence of (4,4) tensors (``viewpoint_cam.full_proj_transform``);
        depthmaps: sequence of (1,H,W) or (H,W) tensors; rgbmaps: sequence of (3,H,W) tensors or None;
        center (3,), radius: the bounding sphere of ``estimate_bounding_sphere`` (:124-135)."""
        if len(full_proj_transforms) != len(depthmaps) or (rgbmaps is not None and len(rgbmaps) != len(depthmaps)):
            raise ValueError("need one projection, one depth map (and one rgb map) per view")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TSDFFusion needs a CUDA device (gsr_b200 has no CPU path)")
        self.radius = float(radius)
        c = [0.0, 0.0, 0.0] if center is None else [float(v) for v in torch.as_tensor(center).flatten().tolist()]
        self._center = (ctypes.c_float * 3)(*c)
        self._keep = []
        blob = bytearray()
        for i, (m, d) in enumerate(zip(full_proj_transforms, depthmaps)):
            d = d.to(self.device, torch.float32)
            d = d.reshape(d.shape[-2], d.shape[-1]).contiguous()
            H, W = d.shape
            rgb_ptr = 0
            if rgbmaps is not None:
                c3 = rgbmaps[i].to(self.device, torch.float32).contiguous()
                if tuple(c3.shape) != (3, H, W):
                    raise ValueError(f"rgb map {i} has shape {tuple(c3.shape)}, expected (3, {H}, {W})")
                self._keep.append(c3)
                rgb_ptr = c3.data_ptr()
            self._keep.append(d)
            mm = torch.as_tensor(m, dtype=torch.float32).cpu().reshape(16).tolist()
            blob += struct.pack("<16f4i2Q", *mm, W, H, 0, 0, d.data_ptr(), rgb_ptr)
        self.nviews = len(depthmaps)
        self.has_rgb = rgbmaps is not None
        self._map_bytes = [int(d.shape[-2]) * int(d.shape[-1]) * 4 for d in depthmaps]
        self.views_per_launch = views_per_launch
        assert len(blob) == self.nviews * _VIEW_BYTES
        self._views = torch.frombuffer(blob, dtype=torch.uint8).to(self.device) if self.nviews else \
            torch.empty(0, dtype=torch.uint8, device=self.device)

    def _launch_groups(self, planes_per_view):
        """Consecutive view ranges (first, count) whose maps fit the L2 budget (view order is preserved: the
        running mean is evaluated in the reference's order)."""
        if self.nviews == 0:
            return [(0, 0)]
        groups, first, acc = [], 0, 0
        for v in range(self.nviews):
            b = self._map_bytes[v] * planes_per_view
            full = (self.views_per_launch is not None and v - first >= self.views_per_launch) or \
                   (self.views_per_launch is None and v > first and acc + b > self.MAP_BUDGET_BYTES[planes_per_view])
            if full:
                groups.append((first, v - first))
                first, acc = v, 0
            acc += b
        groups.append((first, self.nviews - first))
        return groups

    @torch.no_grad()
    def compute_unbounded_tsdf(self, samples, inv_contraction, voxel_size, return_rgb=False):
        if samples.dim() != 2 or samples.shape[1] != 3:
            raise RuntimeError("samples must have dimensions (num_points, 3)")
        if return_rgb and not self.has_rgb:
            raise RuntimeError("return_rgb=True needs rgbmaps")
        pts = f32c(samples, "samples", self.device)
        n = pts.shape[0]
        tsdfs = torch.empty((n,), dtype=torch.float32, device=self.device)
        rgbs = torch.empty((n, 3), dtype=torch.float32, device=self.device) if return_rgb else None
        if n:
            groups = self._launch_groups(4 if return_rgb else 1)
            weights = torch.empty((n,), dtype=torch.float32, device=self.device) if len(groups) > 1 else None
            with on_device(self.device):
                for gi, (v0, nv) in enumerate(groups):
                    check(lib().gsr_tsdf_fuse(
                        n, pts.data_ptr(), int(inv_contraction is not None), self._center, self.radius,
                        float(voxel_size), nv, (self._views.data_ptr() + _VIEW_BYTES * v0) if nv else None,
                        int(gi == 0), tsdfs.data_ptr(), weights.data_ptr() if weights is not None else None,
                        rgbs.data_ptr() if return_rgb else None, stream_ptr(self.device)), "gsr_tsdf_fuse")
        if return_rgb:
            return tsdfs, rgbs
        return tsdfs
