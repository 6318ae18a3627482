"""CUDA-graph capture of the rasterizers (include/gsr_b200.h, "CUDA-graph capture").

The reference extensions cannot be recorded into a CUDA graph: their forward blocks on a device->host copy of
``num_rendered`` (S/cuda_rasterizer/rasterizer_impl.cu:282).  The drop-in rasterizers can: while torch's current
stream is capturing, a forward records every kernel once, laid out for ``capture_margin`` percent of the
``num_rendered`` history of earlier eager calls of the same (device, rasterizer, resolution), and returns without
waiting.  Typical use (a fixed-size training iteration -- between two densification steps of GS-SR's trainer)::

    from gsr_b200.graphs import capture, capture_overflow
    graph, loss = capture(step, warmup=3, before_capture=lambda: optimizer.zero_grad(set_to_none=True))
    for _ in range(n):                       # step = forward + loss + backward + optimizer.step() (capturable=True)
        graph.replay()
    torch.cuda.synchronize()
    if capture_overflow():                   # a replay's num_rendered outgrew the captured capacity
        ...                                  # frame was truncated: raise the margin / warm up again and re-capture
"""
from __future__ import annotations

from ._lib import check, lib

PREFILTERED_VIOLATION = 0xFFFFFFFF


def set_capture_margin(percent: int) -> None:
    """Capacity of a captured forward's binning buffer in percent (>= 100, default 150) of the decayed maximum of
    ``num_rendered`` seen by eager forwards of the same configuration."""
    check(lib().gsr_set_option(b"capture_margin", int(percent)), "gsr_set_option(capture_margin)")


def capture_overflow(reset: bool = True) -> int:
    """0 when every replay since the last reset fitted the captured capacity; otherwise the ``num_rendered`` that did
    not fit (``PREFILTERED_VIOLATION`` for a `prefiltered` violation).  Call it after the replays have completed, from
    the thread that captured the graph."""
    return int(lib().gsr_capture_overflow(1 if reset else 0))


def capture(step, warmup: int = 3, before_capture=None):
    """Record ``step()`` into a CUDA graph the way the rasterizers need it: ``warmup`` eager calls on a side stream first
    (they give every (rasterizer, resolution) its ``num_rendered`` history and let the allocator / optimizer reach their
    steady state), then one captured call.  ``before_capture()`` runs between the two (typically
    ``optimizer.zero_grad(set_to_none=True)``: gradients allocated during capture live in the graph's pool and are
    overwritten, not accumulated, by every replay).  Returns ``(graph, result of the captured step())``; the result's
    tensors are static and hold the outputs of the latest ``graph.replay()``."""
    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(int(warmup), 1)):
            if before_capture is not None:
                before_capture()
            step()
    torch.cuda.current_stream().wait_stream(side)
    if before_capture is not None:
        before_capture()
    capture_overflow(reset=True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        result = step()
    return graph, result
