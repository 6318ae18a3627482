"""CUDA-graph capture of the rasterizers (include/gsr_b200.h, "CUDA-graph capture").

The reference extensions cannot be recorded into a CUDA graph: their forward blocks on a device->host copy of
``num_rendered`` (S/cuda_rasterizer/rasterizer_impl.cu:282).  The drop-in rasterizers can: while torch's current
stream is capturing, a forward records every kernel once, laid out for ``capture_margin`` percent of the
``num_rendered`` history of earlier eager calls of the same (device, rasterizer, resolution), and returns without
waiting.  Typical use (a fixed-size training iteration -- between two densification steps of GS-SR's trainer)::

    for _ in range(3):                       # eager warm-up: num_rendered history, allocator, optimizer state
        step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()                               # forward + loss + backward + optimizer.step() of capturable optimizers
    for _ in range(n):
        g.replay()
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
