"""Spatially tiled block matching of ONE large frame pair across N GPUs (BASELINE configs[3] and the
north star's 8K target): strips of whole block rows, NCCL halo-row exchange of the previous frame
overlapped with the interior rows.  Launch with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_tiled.py [8k|4k]

Each timed iteration = halo exchange + strip kernels on every rank; device time by CUDA events on the
kernel stream (which waits on the exchange), max over ranks.  Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from ofps_b200 import capi, synth
from ofps_b200 import dist as odist


def stream_mode(ctx, which, world, rank, lr, iters):
    """A stream of 17 tiled 8K frames (16 pairs) per step: one halo exchange for all frames + one batched launch."""
    w, h, block, search, n_frames = 7680, 4320, 16, 16, 17
    frames = synth.make_stream(n_frames, w, h, search)
    t = odist.TiledStreamMatcher(ctx, w, h, block, search, 0, n_frames, rank, world)
    t.load(frames)
    del frames

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    iters = max(iters // 5, 5)
    for _ in range(3):
        t.run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(t.kernel_stream)
    for _ in range(iters):
        t.run()
    e1.record(t.kernel_stream)
    barrier()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        v = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{lr}")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms = float(v[0])
    if rank == 0:
        pairs = n_frames - 1
        print(json.dumps({"workload": f"stream of {n_frames} {w}x{h} frames ({pairs} pairs/step), {block}x{block}/+-{search} SAD, tiled over "
                                      f"{world} GPU(s), one NCCL halo exchange per step", "n_gpus": world, "ms_per_step": ms,
                          "ms_per_pair": ms / pairs, "mpix_s": w * h * pairs / ms / 1e3, "iters": iters}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "8k"
    w, h, block, search = (7680, 4320, 16, 16) if which.startswith("8k") else (3840, 2160, 8, 32)
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = capi.Context(lr)
    if which.endswith("stream"):
        return stream_mode(ctx, which, world, rank, lr, iters)
    prev, cur, _ = synth.make_pair(w, h, search, index=3)
    t = odist.TiledBlockMatcher(ctx, w, h, block, search, 0, rank, world)
    t.load(prev, cur)
    del prev, cur

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(5):
        t.run()
    barrier()
    ks = t.kernel_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ks)
    for _ in range(iters):
        t.run()
    e1.record(ks)
    barrier()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        v = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{lr}")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms = float(v[0])
    if rank == 0:
        print(json.dumps({"workload": f"{w}x{h} pair, {block}x{block}/+-{search} SAD, tiled over {world} GPU(s) with NCCL halo rows",
                          "n_gpus": world, "ms_per_pair": ms, "mpix_s": w * h / ms / 1e3, "iters": iters,
                          "halo_rows_each_way": search, "halo_bytes_each_way": search * w}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
