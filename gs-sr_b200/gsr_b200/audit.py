"""Decision audit of a finished rasterizer forward (verification hook, see include/gsr_b200.h gsr_surfel_audit).

``decision_margins(out)`` takes any differentiable output tensor of a drop-in ``GaussianRasterizer`` call (it reaches the
forward's scratch buffers through ``out.grad_fn``, exactly the tensors the backward will read) and returns

    margins  (5,H,W) surfel / (3,H,W) EWA float32: smallest relative distance of any blend decision at the pixel to its
             threshold (alpha vs 1/255, T(1-alpha) vs 1e-4, T vs 0.5, [depth vs 0.2, rho3d vs rho2d]),
    info     (2,H,W) int32: number of blended splats, Gaussian index of the last contributor (-1 = none),
    mismatches  int: pixels where the audit's replay differs from the forward's stored final_T / last contributor (0).
"""
from __future__ import annotations

import torch

from ._lib import check, lib
from ._torch_util import on_device, stream_ptr


def decision_margins(out):
    fn = out.grad_fn
    if fn is None or not hasattr(fn, "num_rendered"):
        raise RuntimeError("decision_margins needs an output of a gsr_b200 GaussianRasterizer call made with autograd enabled")
    kind = getattr(fn, "_forward_cls", type(fn)).__module__.rsplit(".", 1)[-1]     # the drop-in package that made the node
    if not any(k in kind for k in ("surfel", "gaussian", "plane")):
        raise RuntimeError(f"decision_margins: cannot tell the rasterizer family of {type(fn).__name__} ({kind})")
    P, M, H, W = fn.dims
    geom, binning, image = fn.saved_tensors[-3:]
    dev = binning.device
    surfel = "surfel" in kind
    margins = torch.empty((5 if surfel else 3, H, W), dtype=torch.float32, device=dev)
    info = torch.empty((2, H, W), dtype=torch.int32, device=dev)
    mism = torch.zeros((1,), dtype=torch.int32, device=dev)
    if P == 0:
        margins.fill_(1e30); info[0].zero_(); info[1].fill_(-1)
        return margins, info, 0
    with on_device(dev):
        if surfel:
            check(lib().gsr_surfel_audit(P, int(fn.num_rendered), W, H, binning.data_ptr(), image.data_ptr(), margins.data_ptr(),
                                         info.data_ptr(), mism.data_ptr(), stream_ptr(dev)), "gsr_surfel_audit")
        else:
            geo = int("plane" in kind and bool(getattr(fn.raster_settings, "render_geo", False)))
            check(lib().gsr_ewa_audit(P, int(fn.num_rendered), W, H, geo, binning.data_ptr(), image.data_ptr(),
                                      margins.data_ptr(), info.data_ptr(), mism.data_ptr(), stream_ptr(dev)), "gsr_ewa_audit")
    return margins, info, int(mism.item())
