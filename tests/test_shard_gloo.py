"""CPU tests of the N>1 host logic: tile -> rank assignment and the end-of-job exchanges,
run with world_size 2 over gloo (127.0.0.1 rendezvous)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness  # noqa: F401  (sys.path)
from gsr_b200 import shard


def test_assign_tiles_sorted_round_robin():
    names = ["tile_0010", "tile_0002", "tile_0001", "tile_0000", "tile_0003"]
    a = shard.assign_tiles(names, 2)
    assert a == [["tile_0000", "tile_0002", "tile_0010"], ["tile_0001", "tile_0003"]]
    assert shard.assign_tiles(names, 1) == [sorted(names)]
    assert sum(len(x) for x in shard.assign_tiles(names, 8)) == len(names)
    with pytest.raises(ValueError):
        shard.assign_tiles(names, 0)


def test_single_process_fallbacks():
    assert shard.gather_metrics({"a": 1}) == [{"a": 1}]
    assert shard.max_over_ranks(3.5) == 3.5 and shard.sum_over_ranks(2.0) == 2.0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tiles = shard.assign_tiles([f"tile_{i:04d}" for i in range(5)], world)[rank]
        ms = 10.0 + 5.0 * rank
        got = shard.gather_metrics({"rank": rank, "tiles": tiles, "gaussians": 100 * (rank + 1)})
        q.put((rank, shard.max_over_ranks(ms), shard.sum_over_ranks(100.0 * (rank + 1)), got))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    for rank, mx, sm, got in res:
        assert mx == 15.0 and sm == 300.0
        assert [g["rank"] for g in got] == [0, 1]
        assert got[0]["tiles"] == ["tile_0000", "tile_0002", "tile_0004"] and got[1]["tiles"] == ["tile_0001", "tile_0003"]
