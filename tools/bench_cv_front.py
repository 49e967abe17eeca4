"""Measurements of the cv-decoder dense-flow front end (K7 frame_convert, K7b frame_resize, K8 contrast_mask,
K9 flow_entries): device-resident kernel time by CUDA events on the launching stream with the L2 flushed before
every timed launch, algorithmic bytes / time against the measured HBM peak, the host-API (end-to-end) time, and
the CPU oracle (one thread — the reference's cv-decoder loop is single-threaded apart from OpenCV's own pool).

    python tools/bench_cv_front.py            # on a B200 box; one JSON line per case -> gpurun_out/bench_cv_front.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import oracle
from ofps_b200 import capi


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_time(fn, min_time=1.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= min_time:
            return dt / n


def scene_gray(w, h, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    g = (110 + 40 * np.sin(xx / 37.0) * np.cos(yy / 29.0)).astype(np.uint8)
    for _ in range(max(40, w * h // 6000)):
        rw, rh = int(rng.integers(2, 90)), int(rng.integers(2, 90))
        x0, y0 = int(rng.integers(-rw // 2, w)), int(rng.integers(-rh // 2, h))
        g[max(y0, 0):y0 + rh, max(x0, 0):x0 + rw] = rng.integers(0, 256)
    return g


def main():
    oracle.build()
    ctx = capi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.get_stream(), device=0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    peak, peak_src = hbm_peak()
    out = []

    def dev_time(fn, iters=20):
        """Average device time of fn() (launches on ctx's stream), L2 flushed before each timed call."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        l0 = ctx.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in evs:
            with torch.cuda.stream(stream):
                flush.zero_()
                a.record(stream)
                fn()
                b.record(stream)
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / iters * 1e-3, (ctx.launch_count() - l0) // iters

    def emit(**kw):
        out.append(kw)
        print(json.dumps(kw), flush=True)

    def roof(bytes_, t):
        return {"bound": "hbm", "achieved": bytes_ / t / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_ / t / 1e9 / peak,
                "bytes_per_launch": bytes_, "peak_source": peak_src}

    for name, (w, h) in (("1080p", (1920, 1080)), ("4K", (3840, 2160)), ("8K", (7680, 4320))):
        npix = w * h
        rng = np.random.default_rng(w)
        bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        gray = scene_gray(w, h, 5)
        flow = ((rng.random((h, w, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(6)).astype(np.float32)
        gw, gh = capi.mfield_size(w, h)
        d_bgr, d_gray, d_rgba = ctx.dev_alloc(npix * 3), ctx.dev_alloc(npix), ctx.dev_alloc(npix * 4)
        d_mask, d_flow, d_ent = ctx.dev_alloc(npix), ctx.dev_alloc(npix * 8), ctx.dev_alloc(max(gw * gh, 1) * 16)
        d_small = ctx.dev_alloc(gw * gh * 3)
        ctx.to_device(d_bgr, bgr)
        ctx.to_device(d_gray, gray)
        ctx.to_device(d_flow, flow)
        ctx.sync()

        # K7: BGR -> gray (3 B in + 1 B out per pixel)
        t, nl = dev_time(lambda: ctx.frame_convert_dev(d_bgr, w, h, 3 * w, 3, False, d_gray, w))
        t_cpu = cpu_time(lambda: oracle.bgr_to_gray(bgr)) if name != "8K" else None
        emit(case=f"K7 frame_convert BGR->gray {name}", us=t * 1e6, launches=nl, mpix_s=npix / t / 1e6, roofline=roof(4 * npix, t),
             cpu_oracle_ms=t_cpu and t_cpu * 1e3)
        # K7 with the RGBA out_frame as well (3 in + 1 + 4 out)
        t, nl = dev_time(lambda: ctx.frame_convert_dev(d_bgr, w, h, 3 * w, 3, False, d_gray, w, d_rgba))
        emit(case=f"K7 frame_convert BGR->gray+RGBA {name}", us=t * 1e6, launches=nl, mpix_s=npix / t / 1e6, roofline=roof(8 * npix, t))
        ctx.to_device(d_gray, gray)   # the scene (sparse mask) for K8 / K9
        # K7b: resize to the motion-field size ("Process Fullres" off): touches 4 source pixels per output pixel
        t, nl = dev_time(lambda: lib_resize(ctx, d_bgr, w, h, d_small, gw, gh))
        emit(case=f"K7b frame_resize {name} -> {gw}x{gh}", us=t * 1e6, launches=nl, note="latency-bound: 12,600 output pixels")
        # K8: gray -> mask (1 B in + 1 B out per pixel)
        t, nl = dev_time(lambda: ctx.contrast_mask_dev(d_gray, w, h, w, d_mask, w))
        t_cpu = cpu_time(lambda: oracle.contrast_mask(gray)) if name != "8K" else None
        mask = np.empty((h, w), np.uint8)
        ctx.to_host(mask, d_mask)
        ctx.sync()
        emit(case=f"K8 contrast_mask {name}", us=t * 1e6, launches=nl, mpix_s=npix / t / 1e6, roofline=roof(2 * npix, t),
             mask_fraction=float((mask > 0).mean()), cpu_oracle_ms=t_cpu and t_cpu * 1e3)
        # K9: masked flow -> gw x gh cells -> entries (8 B flow + 1 B mask per pixel in, 16 B per touched cell out)
        n = ctx.flow_entries_dev(d_flow, 2 * w, d_mask, w, w, h, gw, gh, d_ent, gw * gh)
        t, nl = dev_time(lambda: launch_flow(ctx, d_flow, w, h, d_mask, gw, gh, d_ent))
        t_cpu = cpu_time(lambda: oracle.flow_entries(flow, mask, gw, gh)) if name != "8K" else None
        emit(case=f"K9 flow_entries (mask, {gw}x{gh} densifier) {name}", us=t * 1e6, launches=nl, mpix_s=npix / t / 1e6,
             roofline=roof(9 * npix + 16 * n, t), entries=n, cpu_oracle_ms=t_cpu and t_cpu * 1e3,
             note="two launches; time includes the 8-byte count read-back the C ABI performs (one stream sync)")
        # K9 without mask (RLOF path)
        n = ctx.flow_entries_dev(d_flow, 2 * w, None, 0, w, h, gw, gh, d_ent, gw * gh)
        t, nl = dev_time(lambda: launch_flow(ctx, d_flow, w, h, None, gw, gh, d_ent))
        emit(case=f"K9 flow_entries (no mask, {gw}x{gh} densifier) {name}", us=t * 1e6, launches=nl, mpix_s=npix / t / 1e6,
             roofline=roof(8 * npix + 16 * n, t), entries=n)
        if name == "1080p":
            # end to end through the host C ABI: gray + flow in pageable host memory -> entries
            t_host = cpu_time(lambda: ctx.cv_flow_frame(gray, flow, True, gw, gh), min_time=0.5)
            t_cpu = cpu_time(lambda: oracle.flow_entries(flow, oracle.contrast_mask(gray), gw, gh))
            emit(case="cv_flow_frame 1080p host API (gray + flow in host memory -> entries)", gpu_ms=t_host * 1e3,
                 cpu_oracle_ms=t_cpu * 1e3, h2d_bytes=9 * npix, note="dominated by the 16.6 MB pageable H2D copy of the flow")
        for p in (d_bgr, d_gray, d_rgba, d_mask, d_flow, d_ent, d_small):
            ctx.dev_free(p)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_cv_front.jsonl"), "w") as f:
        for o in out:
            f.write(json.dumps(o) + "\n")


def lib_resize(ctx, d_src, w, h, d_dst, dw, dh):
    capi.check(capi.lib().ofpsb_frame_resize_dev(ctx._h, d_src, w, h, 3 * w, 3, d_dst, dw, dh, 3 * dw))


def launch_flow(ctx, d_flow, w, h, d_mask, gw, gh, d_ent):
    ctx.flow_entries_dev(d_flow, 2 * w, d_mask, w if d_mask else 0, w, h, gw, gh, d_ent, gw * gh)


if __name__ == "__main__":
    main()
