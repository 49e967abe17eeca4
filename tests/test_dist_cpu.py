"""World-size-2/3 gloo tests (CPU) of the multi-GPU host logic: frame sharding, the strip plan and
the halo-row exchange of ofps_b200/dist.py.  The kernels themselves need a GPU (test_gpu_tiled.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ofps_b200 import dist as odist


def test_round_robin_sharding():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in odist.shard_round_robin(64, r, world))
        assert seen == list(range(64))
        assert odist.shard_round_robin(64, 0, world)[:2] == [0, world][:len(odist.shard_round_robin(64, 0, world))]
    assert odist.shard_round_robin(3, 5, 8) == []


@pytest.mark.parametrize("h,block,search,world", [(2160, 8, 32, 2), (2160, 8, 32, 4), (2160, 8, 32, 8), (4320, 16, 16, 8),
                                                  (1080, 16, 16, 4), (360, 16, 8, 3)])
def test_strip_plan_covers_frame(h, block, search, world):
    plan = odist.strip_plan(h, block, search, world)
    assert sum(s.nby for s in plan) == h // block
    assert plan[0].y0 == 0 and plan[0].halo_top == 0 and plan[-1].halo_bottom == 0
    assert plan[-1].y0 + plan[-1].own_rows == h              # remainder rows stay with the last rank
    for a, b in zip(plan, plan[1:]):
        assert a.y0 + a.rows == b.y0 and a.by0 + a.nby == b.by0
        assert b.halo_top == min(search, b.y0) and a.halo_bottom == min(search, h - b.y0)
    assert max(s.nby for s in plan) - min(s.nby for s in plan) <= 1


def test_strip_plan_rejects_oversized_search():
    with pytest.raises(ValueError):
        odist.strip_plan(64, 8, 32, 8)        # 1 block row per strip, halo of 32 rows spans 4 strips
    with pytest.raises(ValueError):
        odist.strip_plan(64, 16, 4, 8)        # fewer block rows than ranks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _halo_worker(rank, world, port, h, w, block, search, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frame = np.random.default_rng(42).integers(0, 256, (h, w), dtype=np.uint8)   # same on every rank
        plan = odist.strip_plan(h, block, search, world)
        s = plan[rank]
        buf = torch.zeros((s.halo_top + s.own_rows + s.halo_bottom) * w, dtype=torch.uint8)
        buf[s.halo_top * w:(s.halo_top + s.own_rows) * w] = torch.from_numpy(frame[s.y0:s.y0 + s.own_rows].reshape(-1).copy())
        for r in odist.exchange_halos(buf, plan, rank, w):
            r.wait()
        want = frame[s.y0 - s.halo_top:s.y0 + s.own_rows + s.halo_bottom].reshape(-1)
        ok = bool((buf.numpy() == want).all())
        # frame sharding: every rank reports its pair indices, rank 0 checks the union
        mine = torch.zeros(16, dtype=torch.int64)
        idx = odist.shard_round_robin(16, rank, world)
        mine[idx] = 1
        dist.all_reduce(mine)
        ok = ok and bool((mine == 1).all())
        open(os.path.join(out_dir, f"rank{rank}.{'ok' if ok else 'fail'}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,block,search", [(2, 96, 8, 12), (3, 200, 8, 32), (2, 64, 16, 16)])
def test_halo_exchange_gloo(tmp_path, world, h, block, search):
    port = _free_port()
    mp.spawn(_halo_worker, args=(world, port, h, 40, block, search, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"rank{r}.ok" for r in range(world)]
