"""Latency of launches that do not fill the machine (1 GPU): a single 1080p pair, and one strip of an 8K frame cut in 8
(ranks emulated on one device with ofpsb_tiled_connect_local) — tile height 64 vs 32, per-kernel times from the
library's events, whole call from events on the stream (graph replay for the strip)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ofps_b200 import capi, synth

ctx = capi.Context(0)
stream = torch.cuda.ExternalStream(ctx.get_stream(), device=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")


def timed(fn, n=20, do_flush=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        if do_flush:
            with torch.cuda.stream(stream):
                flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3


def kernels(fn):
    ctx.set_option("block_match_profile", 1)
    ks = []
    for _ in range(5):
        fn()
        ks.append(ctx.block_match_kernel_ms())
    ctx.set_option("block_match_profile", 0)
    return [round(float(np.median([k[i] for k in ks])) * 1e3, 1) for i in (0, 1)]


# ---- single 1080p pair
W, H = 1920, 1080
fr = synth.make_stream(2, W, H, 16)
d = ctx.dev_alloc(2 * W * H)
de = ctx.dev_alloc((W // 16) * (H // 16) * 16)
ctx.to_device(d, fr)
one = lambda: ctx.block_match_dev(d, d + W * H, W, H, W, W * H, 1, 16, 16, 0, None, None, de)
for th in (64, 32, 0):
    ctx.set_option("block_match_tile_h", th)
    print(json.dumps({"case": "single 1080p pair", "tile_h": th, "us": timed(one), "sea_list_us": kernels(one)}), flush=True)

# ---- strips of an 8K frame, 8 ranks on one device
W, H = 7680, 4320
fr = synth.make_stream(2, W, H, 16)
ts = [capi.Tiled(ctx, r, 8, W, H, 16, 16, 2) for r in range(8)]
for r, t in enumerate(ts):
    t.connect_local(ts[r - 1] if r > 0 else None, ts[r + 1] if r < 7 else None)
for t in ts:
    for s in (0, 1):
        t.upload(s, fr[s, t.y0:t.y0 + t.own_rows])
        t.publish(s)
outs = [ctx.dev_alloc(t.n_blocks * 16) for t in ts]
for th in (64, 32, 0):
    ctx.set_option("block_match_tile_h", th)
    res = {"case": "one strip of an 8K pair cut in 8 (rank 3)", "tile_h": th}
    t, o = ts[3], outs[3]
    # a tile-height change needs a fresh graph: different output pointer per setting keeps the cache keys apart
    o2 = ctx.dev_alloc(t.n_blocks * 16)
    f = lambda: t.match(0, 1, o2, wait=True)
    res["us"] = timed(f)
    res["sea_list_us"] = kernels(f)
    print(json.dumps(res), flush=True)
    ctx.dev_free(o2)
ctx.set_option("block_match_tile_h", 0)
