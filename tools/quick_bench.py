"""Scratch timing of the block matcher on device-resident frames (wall clock around a sync;
the real measurement lives in bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ofps_b200 import capi, synth

def run(ctx, w, h, block, search, n_pairs, iters=20, metric=0):
    prev, cur = synth.make_batch(min(n_pairs, 4), w, h, search)
    reps = (n_pairs + len(prev) - 1) // len(prev)
    prev = np.concatenate([prev] * reps)[:n_pairs]
    cur = np.concatenate([cur] * reps)[:n_pairs]
    fb = w * h
    dp, dc = ctx.dev_alloc(prev.nbytes), ctx.dev_alloc(cur.nbytes)
    nb = (w // block) * (h // block)
    de = ctx.dev_alloc(nb * n_pairs * 16)
    ctx.to_device(dp, prev); ctx.to_device(dc, cur)
    for _ in range(3):
        ctx.block_match_dev(dp, dc, w, h, w, fb, n_pairs, block, search, metric, None, None, de)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        ctx.block_match_dev(dp, dc, w, h, w, fb, n_pairs, block, search, metric, None, None, de)
    ctx.sync()
    dt = (time.perf_counter() - t0) / iters
    mpix = w * h * n_pairs / dt / 1e6
    ops = nb * n_pairs * (2 * search + 1) ** 2 * block * block
    print(f"{w}x{h} b{block} r{search} m{metric} pairs={n_pairs}: {dt*1e6:9.1f} us/launch  {dt/n_pairs*1e6:8.2f} us/pair "
          f"{mpix:10.0f} Mpix/s  {ops/dt/1e12:6.2f} T absdiff/s  {2*fb*n_pairs/dt/1e9:7.1f} GB/s", flush=True)
    for p in (dp, dc, de):
        ctx.dev_free(p)

if __name__ == "__main__":
    ctx = capi.Context(0)
    print(capi.version(), ctx.device_info())
    for chunk in (0, 4, 8, 16, 32, 64):
        ctx.set_option("block_match_chunk_pairs", chunk)
        print("== chunk", chunk)
        run(ctx, 1920, 1080, 16, 16, 64)
    ctx.set_option("block_match_chunk_pairs", 0)
    run(ctx, 3840, 2160, 8, 32, 8)
    run(ctx, 7680, 4320, 16, 16, 4)
