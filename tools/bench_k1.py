"""K1 iteration bench (GPU box): the headline stream (64 pairs of 1080p, 16x16/+-16 SAD) and its noisy twin through
the default (fused SEA) path, the round-1 pruning pipeline and the exhaustive kernel — device-resident frames,
CUDA events on the launching stream, L2 flushed between steps, results checked bit-for-bit against the exhaustive
kernel before timing.  One JSON object per line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ofps_b200 import capi, synth

W, H, B, R = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (1920, 1080, 16, 16)
PAIRS = int(os.environ.get("PAIRS", "64"))
STEPS = int(os.environ.get("STEPS", "10"))
ctx = capi.Context(0)
stream = torch.cuda.ExternalStream(ctx.get_stream(), device=0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
fb = W * H
nb = (W // B) * (H // B)
d = ctx.dev_alloc((PAIRS + 1) * fb)
de = ctx.dev_alloc(PAIRS * nb * 16)


def step():
    ctx.block_match_dev(d, d + fb, W, H, W, fb, PAIRS, B, R, 0, None, None, de)


def timed():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    for a, b in evs:
        with torch.cuda.stream(stream):
            flush.zero_()
            a.record(stream)
            step()
            b.record(stream)
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


def entries():
    out = np.empty((PAIRS, nb, 4), np.float32)
    ctx.to_host(out, de)
    return out


for noise in (0, 2, 6):
    frames = synth.make_stream(PAIRS + 1, W, H, R, noise_lsb=noise)
    ctx.to_device(d, frames)
    ctx.set_option("block_match_prune", 0)
    step(); ctx.sync()
    ref = entries()
    t_exh = timed()
    ctx.set_option("block_match_prune", 1)
    res = {"frame": [W, H], "block": B, "search": R, "pairs": PAIRS, "noise_lsb": noise,
           "exhaustive_us_per_pair": 1e3 * t_exh / PAIRS}
    for name, pruner in (("sea", 0), ("r1_pipeline", 1)):
        ctx.set_option("block_match_pruner", pruner)
        ctx.set_option("block_match_stats", 1)
        step(); ctx.sync()
        st = ctx.block_match_stats()
        ctx.set_option("block_match_stats", 0)
        same = bool(entries().tobytes() == ref.tobytes())
        t = timed()
        res[name] = {"us_per_pair": 1e3 * t / PAIRS, "Gpix_s": W * H * PAIRS / t / 1e6, "bit_equal_to_exhaustive": same,
                     "stats": {k: int(v) for k, v in st.items()}}
    ctx.set_option("block_match_pruner", 0)
    # SEA kernel / work-list kernel split (events around each inside the library) and the tile-height variants
    for th in (0, 32, 64):
        ctx.set_option("block_match_tile_h", th)
        ctx.set_option("block_match_adaptive", 0)
        t = timed()
        ctx.set_option("block_match_profile", 1)
        ks = []
        for _ in range(3):
            step(); ctx.sync()
            ks.append(ctx.block_match_kernel_ms())
        ctx.set_option("block_match_profile", 0)
        res[f"tile_h_{th}"] = {"us_per_pair": 1e3 * t / PAIRS, "sea_kernel_us_per_pair": 1e3 * ks[-1][0] / PAIRS,
                               "list_kernel_us_per_pair": 1e3 * ks[-1][1] / PAIRS}
    ctx.set_option("block_match_tile_h", 0)
    ctx.set_option("block_match_adaptive", 1)
    print(json.dumps(res), flush=True)

# L2 prefetch distance of the SEA kernel (tiles ahead), noise-free stream
frames = synth.make_stream(PAIRS + 1, W, H, R)
ctx.to_device(d, frames)
for dist in (0, 148, 444, 888, 1776):
    ctx.set_option("block_match_prefetch_tiles", dist)
    print(json.dumps({"prefetch_tiles": dist, "us_per_pair": 1e3 * timed() / PAIRS}), flush=True)
ctx.set_option("block_match_prefetch_tiles", -1)
