#!/usr/bin/env python
"""Headline benchmark: Mpix/s of dense block-matching flow at 1080p, 16x16 blocks, +-16 search (SAD).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch: a synthetic 1080p stream of PAIRS+1 frames
(-> PAIRS frame pairs, BASELINE.json configs[1]) through the block matcher, which emits one
MotionEntry per 16x16 block (K1+K2).  For N > 1 (launched by torchrun, one rank per GPU) every rank
owns its own stream — frames shard across GPUs with no data-path collective ("weak" scaling);
torch.distributed is used only for the barrier and the max-over-ranks of the device time.

Printed JSON (rank 0, one line):
  value      device-resident throughput (frames already in HBM), CUDA-event time, max over ranks
  e2e        same metric through the host C ABI (ofpsb_block_match_batch): pinned host frames in,
             H2D + kernels + D2H of the MotionEntry lists inside the timed region
  roofline   HBM roofline of the dominant kernel (algorithmic bytes / live kernel time) + the
             integer-ALU roofline that actually bounds an exhaustive SAD search
  cpu_baseline  the CPU oracle (plain-C restatement + psadbw, all host threads) on a bounded sample

`--impl reference` times that CPU path only (the reference is Rust and cannot be built here; the
oracle restates it — see DESIGN.md) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, BLOCK, SEARCH, METRIC = 1920, 1080, 16, 16, 0
PAIRS = 64                       # frame pairs per step per GPU (65-frame stream, 135 MB > L2)
NBLOCKS = (W // BLOCK) * (H // BLOCK)
METRIC_NAME = "Mpix/s dense flow @1080p 16x16/+-16 (block-matching SAD)"
WORKLOAD = "1080p block matching 16x16/+-16 SAD, 65-frame synthetic stream = 64 frame pairs per step per GPU"
# algorithmic traffic per frame pair: two u8 planes read once + 16 B per block out
BYTES_PER_PAIR = 2 * W * H + 16 * NBLOCKS
ABSDIFF_PER_PAIR = NBLOCKS * (2 * SEARCH + 1) ** 2 * BLOCK * BLOCK
# ncu --set full captures (profiles/): dram__bytes_read.sum + dram__bytes_write.sum per step of 64 pairs.
#  * pruned path (the default, what `value` times): window_sum 148.8+206.6 MB, prune 397.6+11.1 MB, work list
#    135.6+5.0 MB = 904.7 MB (profiles/r1_pruned_path_ncu.md) — 3.3x the algorithmic bytes: the u16 window-sum plane
#    is written and re-read through HBM and each of the three kernels reads the frames once;
#  * exhaustive kernel: every frame read exactly once, entries written (profiles/r1_block_match_ncu.md).
NCU_TRAFFIC_BYTES_PRUNED_STEP = 904_700_000      # round-1 pipeline (kept for reference)
# round 2, fused SEA path: sea_kernel 135.0 + 8.0 MB, work-list kernel 51.2 + 0.8 MB per 64-pair step
# (profiles/r2_k1_sea_ncu.md) — below the algorithmic bytes because a frame is fetched once for its two pairs
NCU_TRAFFIC_BYTES_SEA_STEP = 195_000_000
NOISE_LSB = 2
TILED_STREAM_PAIRS = 16
NCU_TRAFFIC_BYTES_EXHAUSTIVE_STEP = (PAIRS + 1) * W * H + PAIRS * NBLOCKS * 16


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:  # NVML unavailable: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_block_match_throughput(frames: np.ndarray, target_seconds: float):
    """Oracle (port of the path; SIMD SAD, all host threads) on a bounded sample of the workload."""
    import oracle as orc
    orc.build()
    # every host core this process may run on — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which made
    # the round-1 reference arm single-threaded at N > 1; the oracle takes the count explicitly
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = 0
    t0 = time.perf_counter()
    while True:
        i = n % (len(frames) - 1)
        orc.block_match(frames[i], frames[i + 1], BLOCK, SEARCH, METRIC, threads=threads, fast=True)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= target_seconds or n >= 4096:
            break
    return W * H * n / dt / 1e6, threads, n, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The
    reference is Rust (no rustc in the image) -> the oracle port is what can be timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ofps_b200 import synth
    frames = synth.make_stream(5, W, H, SEARCH)
    for _ in range(max(args.warmup, 1)):
        cpu_block_match_throughput(frames, 0.2)
    vals, total_pairs, total_t, threads = [], 0, 0.0, 1
    for _ in range(args.steps):
        v, threads, n, dt = cpu_block_match_throughput(frames, 1.0)
        vals.append(v)
        total_pairs += n
        total_t += dt
    value = W * H * total_pairs / total_t / 1e6
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{total_pairs} frame pairs in {total_t:.1f} s on the host cores"},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": threads, "kind": "port",
                         "sample": f"{total_pairs} 1080p pairs, exhaustive SAD, psadbw inner loop, OpenMP over blocks"},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_numa(local_rank: int):
    """Pin this rank's threads to the CPUs of its GPU's NUMA node BEFORE any pinned memory is allocated, so that the
    staging buffers are first-touched on the right node (all ranks staged from one node, 21 GB/s each at
    N = 8).  Returns a small description for the JSON line."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = bus.lower()
        if len(dev.split(":")[0]) == 8:      # NVML prints an 8-digit PCI domain, sysfs uses 4
            dev = dev[4:]
        with open(f"/sys/bus/pci/devices/{dev}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return probe_numa(local_rank)
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:
        return {"numa_node": None, "note": type(e).__name__}


def _node_cpus():
    import glob
    out = {}
    for d in glob.glob("/sys/devices/system/node/node[0-9]*"):
        cpus = set()
        with open(os.path.join(d, "cpulist")) as f:
            txt = f.read().strip()
        for part in txt.split(",") if txt else []:
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            out[int(os.path.basename(d)[4:])] = cpus
    return out


def probe_numa(local_rank: int):
    """sysfs does not know the GPU's NUMA node (virtualised PCI topology): when the host has several nodes, time a
    pinned H2D copy first-touched from each of them and bind to the fastest — only if it is clearly (> 10 %) faster."""
    nodes = _node_cpus()
    if len(nodes) < 2:
        return {"numa_node": None, "host_nodes": len(nodes)}
    import torch
    all_cpus = os.sched_getaffinity(0)
    torch.cuda.set_device(local_rank)
    dst = torch.empty(64 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    rates = {}
    for n, cpus in sorted(nodes.items()):
        os.sched_setaffinity(0, cpus)
        src = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        src.fill_(1)                                           # first touch on this node
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        rates[n] = 4 * src.numel() / (a.elapsed_time(b) * 1e-3) / 1e9
        del src
    best = max(rates, key=rates.get)
    info = {"numa_node": None, "host_nodes": len(nodes), "probed_GBps": {str(k): round(v, 1) for k, v in rates.items()}}
    if rates[best] > 1.1 * min(rates.values()):
        os.sched_setaffinity(0, nodes[best])
        info.update(numa_node=best, cpus=len(nodes[best]), how="probed")
    else:
        os.sched_setaffinity(0, all_cpus)
    return info


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ofps_b200 import capi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the ofps_b200 hot path has no CPU fallback")
    numa = bind_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(*vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    ctx = capi.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.get_stream(), device=local_rank)
    frame_bytes = W * H
    # ---- inputs: pinned host stream + device-resident copy (each rank its own stream: weak scaling)
    host = capi.PinnedArray((PAIRS + 1, H, W), np.uint8)
    host.array[:] = synth.make_stream(PAIRS + 1, W, H, SEARCH, first_index=rank)
    host_entries = capi.PinnedArray((PAIRS, NBLOCKS, 4), np.float32)
    d_frames = ctx.dev_alloc((PAIRS + 1) * frame_bytes)
    d_entries = ctx.dev_alloc(PAIRS * NBLOCKS * 16)
    ctx.to_device(d_frames, host.array)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def device_step():
        ctx.block_match_dev(d_frames, d_frames + frame_bytes, W, H, W, frame_bytes, PAIRS, BLOCK, SEARCH, METRIC,
                            None, None, d_entries)

    def e2e_step():
        base = host.array.ctypes.data
        ctx.block_match_raw(base, base + frame_bytes, W, H, W, frame_bytes, PAIRS, BLOCK, SEARCH, METRIC, None, None,
                            host_entries.array)

    # ---- device-resident timing: CUDA events on the launching stream, L2 flushed between steps
    def timed_device_steps(n_steps, step=device_step):
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        l0 = ctx.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for a, b in evs:
            with torch.cuda.stream(stream):
                flush.zero_()
                a.record(stream)
                step()
                b.record(stream)
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs), ctx.launch_count() - l0

    def kernel_times(n_steps):
        """Per-kernel device time (events recorded inside the library around each kernel), L2 flushed per step."""
        ctx.set_option("block_match_profile", 1)
        sea, lst = [], []
        for _ in range(n_steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            device_step()
            a, b = ctx.block_match_kernel_ms()
            sea.append(a)
            lst.append(b)
        ctx.set_option("block_match_profile", 0)
        return float(np.mean(sea)), float(np.mean(lst))

    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_ms, launches = timed_device_steps(args.steps)            # default path: fused SEA kernel + exhaustive work list
    sea_ms, list_ms = kernel_times(max(3, min(args.steps, 10)))
    ctx.set_option("block_match_stats", 1)
    device_step()
    prune_stats = ctx.block_match_stats()
    ctx.set_option("block_match_stats", 0)
    ctx.set_option("block_match_prune", 0)
    exh_ms, exh_launches = timed_device_steps(args.steps)        # every candidate of every block evaluated
    ctx.set_option("block_match_prune", 1)

    # ---- the same stream with sensor noise (+-2 LSB per frame): no pair has an exact match any more
    ctx.to_device(d_frames, synth.make_stream(PAIRS + 1, W, H, SEARCH, first_index=rank, noise_lsb=NOISE_LSB))
    noisy_ms, _ = timed_device_steps(args.steps)
    ctx.set_option("block_match_stats", 1)
    device_step()
    noisy_stats = ctx.block_match_stats()
    ctx.set_option("block_match_stats", 0)
    ctx.to_device(d_frames, host.array)
    device_step()
    ctx.sync()

    # ---- end-to-end timing through the host C ABI (pinned host buffers in, entries out)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.result()

    # ---- what the node can deliver: every rank copies its pinned stream to its GPU at the same time, nothing else running
    # (the ceiling of `e2e` at this N: if per-rank e2e H2D rate sits on it, the host side is saturated, not the GPUs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        ctx.to_device(d_frames, host.array)
    ctx.sync()
    barrier()
    h2d_ceiling = 4 * (PAIRS + 1) * frame_bytes / (time.perf_counter() - t0) / 1e9
    if world > 1:
        t = torch.tensor([h2d_ceiling], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        h2d_ceiling = float(t[0])

    # sanity: e2e results equal the device-resident ones (same frames) — not timed
    chk = np.empty((NBLOCKS, 4), np.float32)
    ctx.to_host(chk, d_entries + (PAIRS - 1) * NBLOCKS * 16)
    if chk.tobytes() != host_entries.array[PAIRS - 1].tobytes():
        raise SystemExit("bench.py: host-API and device-API results differ")

    # ---- frame-by-frame decoder path (ofpsb_stream_*): PAGEABLE caller frames, one frame per call
    pageable = np.array(host.array)                               # plain numpy memory, not page-locked
    if world > 1 and "OFPSB_COPY_THREADS" not in os.environ:
        # ranks share the host: the staging-copy threads of a rank (default: up to 8) are cut to its share of the cores
        # (8 ranks on the pool's 32-core boxes: 2 each instead of 64 spinning threads on 32 cores)
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        os.environ["OFPSB_COPY_THREADS"] = str(max(1, min(8, len(os.sched_getaffinity(0)) // (2 * local_world))))
    st = capi.FrameStream(ctx, W, H, BLOCK, SEARCH, METRIC, depth=6)
    out = np.empty((NBLOCKS, 4), np.float32)
    last = None

    def frame_pass(pipelined, src=None):
        nonlocal last
        src = pageable if src is None else src
        n_out = 0
        for i in range(PAIRS + 1):
            if pipelined:
                st.submit(src[i])
                if i >= 3 and st.collect(out) is not None:
                    n_out += 1
            elif st.push(src[i], out) is not None:
                n_out += 1
        while pipelined and st.collect(out) is not None:
            n_out += 1
        last = out.copy()
        return n_out

    frame_pass(True)
    barrier()
    t0 = time.perf_counter()
    n_pipe = frame_pass(True)
    t_pipe = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    n_sync = frame_pass(False)
    t_sync = time.perf_counter() - t0
    frame_pass(True, host.array)
    barrier()
    t0 = time.perf_counter()
    n_pin = frame_pass(True, host.array)                           # page-locked caller frames: no staging copy
    t_pin = time.perf_counter() - t0
    frame_pass(True)                                               # leave `last` = result of a pageable pass
    st.close()
    # (every pass restarts the stream at frame 0 after frame PAIRS: that extra pair is counted, its result unused)
    if last.tobytes() != host_entries.array[PAIRS - 1].tobytes():
        raise SystemExit("bench.py: streaming-API and batch-API results differ")
    dev_ms, e2e_ms, exh_ms, noisy_ms, sea_ms, list_ms, t_pipe, t_sync, t_pin = rank_max(dev_ms, e2e_ms, exh_ms, noisy_ms, sea_ms,
                                                                                        list_ms, t_pipe, t_sync, t_pin)

    extra = {}
    if world > 1:
        extra["strong_c5"] = strong_c5(ctx, torch, dist, stream, flush, rank, world, dev, barrier, rank_max, args)
        extra["tiled_8k"] = tiled_8k(ctx, torch, dist, stream, flush, rank, world, dev, barrier, rank_max)
        extra["tiled_c4"] = tiled_8k(ctx, torch, dist, stream, flush, rank, world, dev, barrier, rank_max, w=3840, h=2160, BLOCK=8,
                                     SEARCH=32, n_stream=8,
                                     label=f"BASELINE config 4: ONE 3840x2160 pair, 8x8/+-32 SAD, {world} strips; halo rows read from the "
                                           "neighbours' HBM inside the +-32 SEA kernel (peer-mapped tensor maps), no exchange step")

    if rank == 0:
        hbm_peak, peak_src, sm_max = _peaks()
        pix_per_step = W * H * PAIRS * world
        value = pix_per_step * args.steps / (dev_ms * 1e-3) / 1e6
        e2e_value = pix_per_step * args.steps / (e2e_ms * 1e-3) / 1e6
        step_s = dev_ms * 1e-3 / args.steps                        # one pass of the hot path = SEA kernel + work list
        achieved = BYTES_PER_PAIR * PAIRS / step_s / 1e9
        sm_count = ctx.device_info()["sm_count"]
        clk = (clocks["sm_mhz"] or sm_max) * 1e6
        alu_peak = sm_count * 64 * 4 * clk / 1e12                  # VABSDIFF4: 64 lanes/clk/SM, 4 px each (measured)
        exh_s = exh_ms * 1e-3 / args.steps
        alu_ach = ABSDIFF_PER_PAIR * PAIRS / exh_s / 1e12
        h2d = (PAIRS + 1) * frame_bytes
        line = {
            "metric": METRIC_NAME, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [W, H], "block": BLOCK, "search": SEARCH, "metric": "SAD",
                       "pairs_per_step_per_gpu": PAIRS, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                       "l2": "512 MB buffer written between timed steps (L2 flush); per-step CUDA events",
                       "content": "noise-free synthetic motion: most blocks have an exact match, the best case of the exact "
                                  "pruning; `noisy` below is the same stream with +-2 LSB sensor noise per frame",
                       "search_mode": "fused four-term successive elimination (SEA) + exhaustive search of undecided blocks "
                                      "(bit-identical to exhaustive search; data-dependent)"},
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": PAIRS * NBLOCKS * 16, "ms_per_step": e2e_ms / args.steps,
                    "h2d_GBps_per_rank": h2d / (e2e_ms * 1e-3 / args.steps) / 1e9,
                    "h2d_ceiling_GBps_per_rank": h2d_ceiling, "numa": numa,
                    "ceiling_note": "h2d_ceiling = all ranks copying the same pinned frames concurrently with no kernels "
                                    "(min over ranks): e2e cannot exceed it",
                    "api": "ofpsb_block_match_batch (pinned host frames -> MotionEntry lists)"},
            "e2e_frame": {"value": W * H * n_pipe * world / t_pipe / 1e6, "unit": "Mpix/s", "frames_per_s_per_gpu": n_pipe / t_pipe,
                          "sync_push": {"value": W * H * n_sync * world / t_sync / 1e6, "us_per_frame": 1e6 * t_sync / n_sync},
                          "pinned_frames": {"value": W * H * n_pin * world / t_pin / 1e6, "us_per_frame": 1e6 * t_pin / n_pin},
                          "h2d_bytes_per_frame": frame_bytes, "d2h_bytes_per_frame": NBLOCKS * 16,
                          "api": "ofpsb_stream_submit / _collect, one PAGEABLE 1080p frame per call (the drop-in Decoder path); "
                                 "sync_push = ofpsb_stream_push, result returned by the same call; pinned_frames = the "
                                 "pipelined form on page-locked caller frames (no staging copy)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": NCU_TRAFFIC_BYTES_SEA_STEP, "peak_source": peak_src,
                         "kernel": "hot-path pass = sea_kernel<16,16> (dominant) + block_match_list_kernel<16,16,17,4,288,SAD>",
                         "bytes_per_launch": BYTES_PER_PAIR * PAIRS, "kernel_ms": step_s * 1e3,
                         "dominant_kernel": {"name": "sea_kernel<16,16>", "ms": sea_ms, "share_of_pass": sea_ms / (sea_ms + list_ms),
                                             "achieved": BYTES_PER_PAIR * PAIRS / (sea_ms * 1e-3) / 1e9,
                                             "frac": BYTES_PER_PAIR * PAIRS / (sea_ms * 1e-3) / 1e9 / hbm_peak,
                                             "timed": "CUDA events recorded by the library around the kernel on its stream"},
                         "worklist_kernel_ms": list_ms,
                         "pruning": {"blocks": prune_stats["blocks"], "decided_by_bounds": prune_stats["decided"],
                                     "exhaustive_worklist": prune_stats["worklist"], "exact_evals": prune_stats["exact_evals"]}},
            "noisy": {"value": pix_per_step * args.steps / (noisy_ms * 1e-3) / 1e6, "unit": "Mpix/s", "noise_lsb": NOISE_LSB,
                      "ms_per_step": noisy_ms / args.steps,
                      "pruning": {"blocks": noisy_stats["blocks"], "decided_by_bounds": noisy_stats["decided"],
                                  "exhaustive_worklist": noisy_stats["worklist"], "exact_evals": noisy_stats["exact_evals"]}},
            "exhaustive": {"value": pix_per_step * args.steps / (exh_ms * 1e-3) / 1e6, "unit": "Mpix/s",
                           "ms_per_step": exh_ms / args.steps, "gpu_launches": int(exh_launches),
                           "kernel": "block_match_tma_kernel<16,16,17,4,288,SAD> (every candidate of every block)",
                           "hbm": {"achieved": BYTES_PER_PAIR * PAIRS / exh_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": BYTES_PER_PAIR * PAIRS / exh_s / 1e9 / hbm_peak,
                                   "traffic": NCU_TRAFFIC_BYTES_EXHAUSTIVE_STEP},
                           "alu": {"bound": "int-alu (VABSDIFF4.U8.ACC 64 lanes/clk/SM; exhaustive SAD is ~540 abs-diff/byte)",
                                   "achieved": alu_ach, "peak": alu_peak, "unit": "T absdiff/s", "frac": alu_ach / alu_peak}},
            "clocks": clocks,
        }
        line.update(extra)
        if world == 1:
            v, threads, n, dt = cpu_block_match_throughput(host.array[:5], 10.0)
            line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": threads, "kind": "port",
                                    "sample": f"{n} 1080p pairs of the same stream in {dt:.1f} s, exhaustive SAD "
                                              "(psadbw, OpenMP over blocks)"}
        print(json.dumps(line), flush=True)
    ctx.dev_free(d_frames)
    ctx.dev_free(d_entries)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def strong_c5(ctx, torch, dist, stream, flush, rank, world, dev, barrier, rank_max, args):
    """BASELINE config 5 as written: ONE 64-pair 1080p batch, pairs sharded round-robin over the GPUs (strong scaling).
    Every rank checks its pairs bit-for-bit against the same pairs matched in stream layout on its own GPU."""
    from ofps_b200 import capi, synth
    from ofps_b200 import dist as odist
    frames = synth.make_stream(PAIRS + 1, W, H, SEARCH, first_index=1000)     # the same stream on every rank
    mine = odist.shard_round_robin(PAIRS, rank, world)
    fb = W * H
    n = len(mine)
    d_pairs = ctx.dev_alloc(2 * n * fb)                                        # [prev pairs | cur pairs]
    d_out = ctx.dev_alloc(max(n, 1) * NBLOCKS * 16)
    ctx.to_device(d_pairs, np.ascontiguousarray(np.concatenate([frames[mine], frames[[i + 1 for i in mine]]])))

    def step():
        if n:
            ctx.block_match_dev(d_pairs, d_pairs + n * fb, W, H, W, fb, n, BLOCK, SEARCH, METRIC, None, None, d_out)

    step()
    ctx.sync()
    got = np.empty((n, NBLOCKS, 4), np.float32)
    ctx.to_host(got, d_out)
    d_all = ctx.dev_alloc((PAIRS + 1) * fb)
    d_ref = ctx.dev_alloc(PAIRS * NBLOCKS * 16)
    ctx.to_device(d_all, frames)
    ctx.block_match_dev(d_all, d_all + fb, W, H, W, fb, PAIRS, BLOCK, SEARCH, METRIC, None, None, d_ref)
    ref = np.empty((PAIRS, NBLOCKS, 4), np.float32)
    ctx.to_host(ref, d_ref)
    ok = got.tobytes() == ref[mine].tobytes()

    def timed(fn, n_steps):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n_steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(rank_max(a.elapsed_time(b))[0])
        return float(np.median(ts))

    ms = timed(step, args.steps)
    # the whole batch on ONE GPU (rank 0 measures, the others wait at the barriers inside `timed`)
    one_ms = timed((lambda: ctx.block_match_dev(d_all, d_all + fb, W, H, W, fb, PAIRS, BLOCK, SEARCH, METRIC, None, None, d_ref))
                   if rank == 0 else (lambda: None), max(3, args.steps // 2))
    okf = rank_max(0.0 if ok else 1.0)[0] == 0.0
    for p in (d_pairs, d_out, d_all, d_ref):
        ctx.dev_free(p)
    return {"workload": "ONE batch of 64 1080p pairs, 16x16/+-16 SAD, pair i -> rank i % N (BASELINE config 5)", "scaling": "strong",
            "value": W * H * PAIRS / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms": ms, "one_gpu_ms": one_ms, "speedup_vs_1gpu": one_ms / ms,
            "bit_equal_to_single_gpu": okf, "timing": "device events per rank, L2 flushed, max over ranks, median of steps"}


def tiled_8k(ctx, torch, dist, stream, flush, rank, world, dev, barrier, rank_max, w=7680, h=4320, BLOCK=BLOCK, SEARCH=SEARCH,
             n_stream=None, label=None):
    """North-star geometry: ONE 7680x4320 pair, 16x16/+-16, cut into N strips; halo rows are read from the neighbours'
    HBM inside the kernel (ofpsb_tiled_*, CUDA IPC + NVLink), no exchange step.  Also a stream of tiled frames in one
    launch sequence.  Bit-equality against the whole frame matched on each rank's own GPU comes first.
    (Also used for BASELINE config 4: 3840x2160, 8x8/+-32.)"""
    from ofps_b200 import capi, synth
    from ofps_b200 import dist as odist
    n_stream = TILED_STREAM_PAIRS if n_stream is None else n_stream
    frames = synth.make_stream(n_stream + 1, w, h, SEARCH, first_index=2000)
    m = odist.PeerTiledMatcher(ctx, w, h, BLOCK, SEARCH, rank, world, n_slots=n_stream + 1)
    t = m.t
    for s in range(n_stream + 1):
        m.load(s, frames[s])
    barrier()
    for _ in range(3):
        m.match(0, 1)
    ctx.sync()
    whole = ctx.block_match(frames[0], frames[1], BLOCK, SEARCH, 0, want=("entries",))["entries"].reshape(-1, 4)
    mine = m.entries.cpu().numpy()
    ok = mine.tobytes() == whole[t.y0 // BLOCK * t.nbx:(t.y0 // BLOCK + t.nby) * t.nbx].tobytes()
    gathered_ok = m.gather_entries().tobytes() == whole.tobytes()
    out = torch.zeros((n_stream, t.n_blocks, 4), dtype=torch.float32, device=dev)

    def timed(fn, n_steps=15):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n_steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(rank_max(a.elapsed_time(b))[0])
        return float(np.median(ts)) * 1e3

    pair_us = timed(lambda: m.match(0, 1, wait=True))
    stream_us = timed(lambda: t.match_stream(0, n_stream, out.data_ptr(), wait=True), 8)
    # the same pair / stream on ONE GPU (rank 0 measures; a world-1 strip object is the whole frame)
    one_pair_us = one_stream_us = 0.0
    solo = capi.Tiled(ctx, 0, 1, w, h, BLOCK, SEARCH, n_stream + 1) if rank == 0 else None
    if solo:
        for s in range(n_stream + 1):
            solo.upload(s, frames[s])
        d1 = ctx.dev_alloc(n_stream * solo.n_blocks * 16)
    one_pair_us = timed((lambda: solo.match(0, 1, d1, wait=False)) if solo else (lambda: None))
    one_stream_us = timed((lambda: solo.match_stream(0, n_stream, d1, wait=False)) if solo else (lambda: None), 8)
    if solo:
        ctx.dev_free(d1)
        solo.close()
    okf = rank_max(0.0 if (ok and gathered_ok) else 1.0)[0] == 0.0
    m.close()
    return {"workload": label or f"ONE 7680x4320 pair, 16x16/+-16 SAD, {world} strips of whole block rows; halo rows read from the "
                                 "neighbours' HBM inside the SEA kernel (peer-mapped tensor maps), no exchange step",
            "pair_us": pair_us, "pair_one_gpu_us": one_pair_us, "pair_speedup_vs_1gpu": one_pair_us / pair_us,
            "pair_value": w * h / pair_us, "stream_pairs": n_stream, "stream_us": stream_us, "stream_one_gpu_us": one_stream_us,
            "stream_speedup_vs_1gpu": one_stream_us / stream_us, "stream_value": w * h * n_stream / stream_us, "unit": "Mpix/s",
            "bit_equal_to_single_gpu": okf,
            "timing": "device events per rank around the launch sequence (neighbour-ready wait included), L2 flushed, "
                      "max over ranks, median of steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--impl", args.impl]
        return subprocess.call(cmd)
    # stdout carries exactly one JSON line: everything else written to fd 1 while the job runs (NCCL prints its
    # version banner there from C, whatever NCCL_DEBUG says) is sent to stderr, the line goes to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w", buffering=1)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
