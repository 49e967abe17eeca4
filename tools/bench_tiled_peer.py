"""Tiled single frame pair over N GPUs with peer-mapped halo rows (ofpsb_tiled_*): correctness against the whole-frame
result on every rank, then latency of ONE pair (max over ranks, median over iterations) and of a stream of pairs.

    python -m torch.distributed.run --nproc-per-node N tools/bench_tiled_peer.py [W H BLOCK SEARCH]     (or plain python: N=1)
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from ofps_b200 import capi, synth
from ofps_b200 import dist as odist

W, H, B, R = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (7680, 4320, 16, 16)
ITERS = int(os.environ.get("ITERS", "30"))
STREAM = int(os.environ.get("STREAM_PAIRS", "16"))
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


ctx = capi.Context(local)
stream = torch.cuda.ExternalStream(ctx.get_stream(), device=local)
frames = synth.make_stream(STREAM + 1, W, H, R) if STREAM > 1 else None
pair = frames[:2] if frames is not None else np.stack(synth.make_pair(W, H, R, index=3)[:2])
m = odist.PeerTiledMatcher(ctx, W, H, B, R, rank, world, n_slots=STREAM + 1)
t = m.t
for s in range(STREAM + 1 if frames is not None else 2):
    m.load(s, frames[s] if frames is not None else pair[s])
barrier()

# ---- correctness: this rank's strip against the whole-frame result computed on this GPU
m.match(0, 1)
m.match(0, 1)
m.match(0, 1)        # eager, capture, replay
ctx.sync()
whole = ctx.block_match(pair[0], pair[1], B, R, 0, want=("entries",))["entries"].reshape(-1, 4)
mine = m.entries.cpu().numpy()
ok = mine.tobytes() == whole[t.y0 // B * t.nbx:(t.y0 // B + t.nby) * t.nbx].tobytes()
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
bit_equal = bool(flag.item())
gathered = m.gather_entries()
gather_ok = gathered.tobytes() == whole.tobytes()

flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")


def time_single(wait, flush_l2):
    ts = []
    for _ in range(ITERS):
        if flush_l2:
            with torch.cuda.stream(stream):
                flush.zero_()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        m.match(0, 1, wait=wait)
        b.record(stream)
        torch.cuda.synchronize()
        v = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ts.append(float(v.item()))
    return float(np.median(ts)) * 1e3


def time_stream(batched):
    out = torch.zeros((STREAM, t.n_blocks, 4), dtype=torch.float32, device=f"cuda:{local}")

    def go():
        if batched:
            t.match_stream(0, STREAM, out.data_ptr(), wait=True)
        else:
            for i in range(STREAM):
                t.match(i, i + 1, out[i].data_ptr(), wait=True)

    for _ in range(3):
        go()
    ts = []
    for _ in range(max(ITERS // 3, 5)):
        with torch.cuda.stream(stream):
            flush.zero_()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        go()
        b.record(stream)
        torch.cuda.synchronize()
        v = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ts.append(float(v.item()))
    return float(np.median(ts)) * 1e3, out


res = {"frame": [W, H], "block": B, "search": R, "n_gpus": world, "bit_equal_strips": bit_equal, "bit_equal_gathered": gather_ok,
       "pair_us_l2_flushed": time_single(True, True), "pair_us_l2_flushed_nowait": time_single(False, True),
       "pair_us_l2_warm": time_single(True, False)}
if STREAM > 1:
    us1, o1 = time_stream(False)
    us, o2 = time_stream(True)
    res["stream_pairs"] = STREAM
    res["stream_us_pair_by_pair"] = us1
    res["stream_us"] = us
    res["stream_Gpix_s"] = W * H * STREAM / us / 1e3
    res["stream_batched_equals_pair_by_pair"] = bool(torch.equal(o1, o2))
res["pair_Gpix_s"] = W * H / res["pair_us_l2_flushed"] / 1e3
# per-kernel device times of this rank's strip (events inside the library; eager launches, no graph)
ctx.set_option("block_match_profile", 1)
ks = []
for _ in range(5):
    barrier()
    m.match(0, 1, wait=False)
    ks.append(ctx.block_match_kernel_ms())
ctx.set_option("block_match_profile", 0)
kt = torch.tensor([float(np.median([k[0] for k in ks])), float(np.median([k[1] for k in ks]))], dtype=torch.float64, device=f"cuda:{local}")
allk = [torch.zeros_like(kt) for _ in range(world)]
if world > 1:
    dist.all_gather(allk, kt)
else:
    allk = [kt]
res["sea_us_per_rank"] = [round(float(k[0]) * 1e3, 1) for k in allk]
res["list_us_per_rank"] = [round(float(k[1]) * 1e3, 1) for k in allk]
if rank == 0:
    print(json.dumps(res), flush=True)
m.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
