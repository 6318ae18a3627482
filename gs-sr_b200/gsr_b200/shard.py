"""One-tile-per-GPU sharding of VastGaussian tiles (the reference trains them sequentially
in one process: /root/reference/train_split.py:24-38).

Tiles are independent sub-scenes (own Gaussians, cameras, optimizer), so the data path has
NO collective: rank r works on tiles r, r+world, ...  The only exchanges are the end-of-job
gather of per-tile metrics and the max-over-ranks of device timings, both through
torch.distributed (NCCL on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

import os
import re
from typing import Any, Dict, List, Sequence


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), \
        int(os.environ.get("LOCAL_RANK", "0"))


def _tile_key(name: str):
    m = re.search(r"(\d+)$", name)
    return (0, int(m.group(1))) if m else (1, name)


def assign_tiles(tile_names: Sequence[str], world_size: int) -> List[List[str]]:
    """Deterministic round-robin of tiles over ranks.  Names are sorted by their numeric
    suffix first: the reference pairs an UNSORTED os.listdir with index-named configs
    (train_split.py:15-16)."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    ordered = sorted(tile_names, key=_tile_key)
    return [list(ordered[r::world_size]) for r in range(world_size)]


def gather_metrics(local: Dict[str, Any]) -> List[Dict[str, Any]]:
    """Per-rank metric dicts gathered on every rank (rank order). Single process: [local]."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out: List[Any] = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device-side timing) over all ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
