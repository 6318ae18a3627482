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


# ---- config 5: per-tile TSDF volumes reduced onto one rank (extract_mesh_split.py:58-119) --------------------------------
def _tsdf_worker(rank, world, port, q):
    """Each rank fuses ITS tile's views (box-filtered cameras) into the same bounded lattice with the CPU oracle, then the
    partial volumes are reduced onto rank 0; the result must equal fusing all views in one volume."""
    import numpy as np
    from gsr_b200.tsdf import cameras_in_box, reduce_partial_volume
    from oracle import tsdf_oracle
    from tsdf_synth import build_tsdf_case
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = build_tsdf_case("world_rgb", n=8)
        eyes = c["eyes"]
        boxes = [[-10.0, 0.0, -10.0, 10.0], [0.0, 10.0, -10.0, 10.0]]            # two tiles split at x = 0 (box.txt rectangles)
        mine = cameras_in_box(eyes, boxes[rank])
        grid = dict(origin=(-0.8, -0.8, -0.8), voxel_size=0.1, dims=(16, 16, 16), sdf_trunc=0.3, depth_trunc=5.0)
        sel = lambda xs: [xs[i] for i in mine]  # noqa: E731
        t, w, rgb = tsdf_oracle.integrate_grid(projs=sel(c["projs"]), depthmaps=sel(c["depthmaps"]), rgbmaps=sel(c["rgbmaps"]), **grid)
        out = reduce_partial_volume(torch.from_numpy(t), torch.from_numpy(w), torch.from_numpy(rgb), dst=0)
        if rank == 0:
            every = sorted(cameras_in_box(eyes, boxes[0]) + cameras_in_box(eyes, boxes[1]))
            pick = lambda xs: [xs[i] for i in every]  # noqa: E731
            ft, fw, frgb = tsdf_oracle.integrate_grid(projs=pick(c["projs"]), depthmaps=pick(c["depthmaps"]), rgbmaps=pick(c["rgbmaps"]), **grid)
            q.put((len(mine), len(every), float(np.abs(out[0].numpy() - ft).max()), float(np.abs(out[1].numpy() - fw).max()),
                   float(np.abs(out[2].numpy() - frgb).max()), float(fw.max())))
        else:
            assert out is None
            q.put((len(mine),))
    finally:
        dist.destroy_process_group()


def test_two_rank_tsdf_volume_reduce():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tsdf_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    full = [r for r in res if len(r) > 1][0]
    other = [r for r in res if len(r) == 1][0]
    n0, n_all, dt, dw, drgb, wmax = full
    assert n0 > 0 and other[0] > 0 and n0 + other[0] == n_all            # both tiles own cameras, every camera is in one tile
    assert dw == 0.0 and wmax > 2.0                                       # weights are small integers: the sum is exact
    assert dt <= 2e-6 and drgb <= 2e-6                                    # a weighted mean re-associated: float rounding only


def test_combine_partial_volumes_matches_sequential_integration():
    import numpy as np
    from gsr_b200.tsdf import combine_partial_volumes
    from oracle import tsdf_oracle
    from tsdf_synth import build_tsdf_case
    c = build_tsdf_case("world_rgb", n=8)
    grid = dict(origin=(-0.8, -0.8, -0.8), voxel_size=0.1, dims=(16, 12, 10), sdf_trunc=0.3, depth_trunc=5.0)
    parts = []
    for sl in (slice(0, 2), slice(2, 3), slice(3, 5)):
        t, w, rgb = tsdf_oracle.integrate_grid(projs=c["projs"][sl], depthmaps=c["depthmaps"][sl], rgbmaps=c["rgbmaps"][sl], **grid)
        parts.append((torch.from_numpy(t), torch.from_numpy(w), torch.from_numpy(rgb)))
    t, w, rgb = combine_partial_volumes(parts)
    ft, fw, frgb = tsdf_oracle.integrate_grid(projs=c["projs"], depthmaps=c["depthmaps"], rgbmaps=c["rgbmaps"], **grid)
    assert t.shape == (10, 12, 16) and np.array_equal(w.numpy(), fw)
    assert np.abs(t.numpy() - ft).max() <= 2e-6 and np.abs(rgb.numpy() - frgb).max() <= 2e-6
    # continuing a volume equals integrating everything at once (the init=0 continuation of the C ABI)
    st = tsdf_oracle.integrate_grid(projs=c["projs"][:2], depthmaps=c["depthmaps"][:2], rgbmaps=c["rgbmaps"][:2], **grid)
    ct, cw, crgb = tsdf_oracle.integrate_grid(projs=c["projs"][2:], depthmaps=c["depthmaps"][2:], rgbmaps=c["rgbmaps"][2:], state=st, **grid)
    assert np.array_equal(ct, ft) and np.array_equal(cw, fw) and np.array_equal(crgb, frgb)
