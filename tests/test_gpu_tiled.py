"""Spatially tiled block matching (BASELINE config 3: 4K 8x8/+-32 across 2/4/8 GPUs).

On a 1-GPU box the strips of every rank are run one after the other on the same device with the halo
rows copied straight from the full frame (the exchange itself is covered by the gloo tests and by the
multi-GPU test below); the concatenated strips must equal the whole-frame result bit-for-bit."""
import os
import socket

import numpy as np
import pytest

from ofps_b200 import capi, synth
from ofps_b200 import dist as odist

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,block,search,world", [(640, 360, 16, 8, 2), (512, 288, 8, 32, 3), (1920, 1080, 16, 16, 4),
                                                    (3840, 2160, 8, 32, 8)])
def test_strips_equal_whole_frame(ctx, w, h, block, search, world):
    prev, cur, _ = synth.make_pair(w, h, search, index=7)
    whole = ctx.block_match(prev, cur, block, search, 0)
    ents, mvs, costs = [], [], []
    for rank in range(world):
        t = odist.TiledBlockMatcher(ctx, w, h, block, search, 0, rank, world)
        t.load(prev, cur, fill_halos=True)
        t.run(exchange=False)
        ctx.sync()
        ents.append(t.entries.cpu().numpy())
        mvs.append(t.mv.cpu().numpy())
        costs.append(t.cost.cpu().numpy())
    assert np.concatenate(ents).tobytes() == whole["entries"].tobytes()
    np.testing.assert_array_equal(np.concatenate(mvs).reshape(whole["mv"].shape), whole["mv"])
    np.testing.assert_array_equal(np.concatenate(costs).reshape(whole["cost"].shape).astype(np.uint32), whole["cost"])


@pytest.mark.parametrize("w,h,block,search,world", [(640, 368, 16, 16, 2), (512, 288, 8, 32, 3), (640, 360, 16, 16, 2),
                                                    (648, 364, 8, 16, 3)])
def test_stream_strips_equal_whole_frames(ctx, w, h, block, search, world):
    frames = synth.make_stream(4, w, h, search)
    whole = ctx.block_match(frames[:-1], frames[1:], block, search, 0, want=("entries",))["entries"]
    parts = []
    for rank in range(world):
        t = odist.TiledStreamMatcher(ctx, w, h, block, search, 0, len(frames), rank, world)
        t.load(frames, fill_halos=True)
        t.run(exchange=False)
        ctx.sync()
        parts.append(t.entries.cpu().numpy())
    got = np.concatenate(parts, axis=1)
    assert got.tobytes() == whole.tobytes()


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        w, h, block, search = 3840, 2160, 8, 32
        prev, cur, _ = synth.make_pair(w, h, search, index=7)
        ctx = capi.Context(rank)
        t = odist.TiledBlockMatcher(ctx, w, h, block, search, 0, rank, world)
        t.load(prev, cur)
        t.run()
        got = t.gather_entries()
        # stream of tiled frames: one halo exchange for all frames, one batched strip launch
        frames = synth.make_stream(3, 1920, 1088, 16)
        ts = odist.TiledStreamMatcher(ctx, 1920, 1088, 16, 16, 0, 3, rank, world)
        ts.load(frames)
        ts.run()
        ctx.sync()
        mine = ts.entries.cpu().numpy()
        ref = ctx.block_match(frames[:-1], frames[1:], 16, 16, 0, want=("entries",))["entries"]
        s = ts.strip
        stream_ok = mine.tobytes() == ref[:, s.by0 * ts.nbx:(s.by0 + s.nby) * ts.nbx].tobytes()
        flags = torch.tensor([int(stream_ok)], device=f"cuda:{rank}")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            whole = ctx.block_match(prev, cur, block, search, 0, want=("entries",))["entries"]
            ok = got.tobytes() == whole.tobytes() and bool(flags.item())
            # detector on the gathered list == detector on the single-GPU list (same order, same bits)
            a = ctx.detect_block_motion(got)
            b = ctx.detect_block_motion(whole)
            ok = ok and a[:3] == b[:3] and a[3].tobytes() == b[3].tobytes()
            open(os.path.join(out_dir, "ok" if ok else "fail"), "w").close()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_tiled_4k_over_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert os.listdir(tmp_path) == ["ok"]
