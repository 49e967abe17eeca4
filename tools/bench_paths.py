"""Secondary measurements of the hot path beyond the headline block matcher: detector (K3+K4), fused
Detection-tab frame, Almeida estimator (K5/K6) and the densifier's sort path, each next to the CPU
oracle (single thread — every reference plugin call is single-threaded).  Prints one JSON line per case.

    python tools/bench_paths.py            # on a B200 box
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from ofps_b200 import capi, synth


def timeit(fn, min_time=0.5, min_iters=5):
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= min_time and n >= min_iters:
            return dt / n


def main():
    oracle.build()
    ctx = capi.Context(0)
    out = []

    def emit(**kw):
        out.append(kw)
        print(json.dumps(kw), flush=True)

    # ---- detector on a 1080p block-match field (8,040 entries), defaults -> 14x14 grid
    prev, cur, _ = synth.make_pair(1920, 1080, 16, index=0)
    ent = ctx.block_match(prev, cur, 16, 16, 0)["entries"]
    t_gpu = timeit(lambda: ctx.detect_block_motion(ent))
    t_cpu = timeit(lambda: oracle.detect_block_motion(ent))
    d_ent = ctx.dev_alloc(ent.nbytes)
    ctx.to_device(d_ent, ent)
    t_dev = timeit(lambda: ctx.detect_block_motion(None, d_entries=d_ent, n=len(ent)))
    emit(case="detect_block_motion 8040 entries -> 14x14 (K3+K4)", gpu_host_api_us=t_gpu * 1e6, gpu_device_entries_us=t_dev * 1e6,
         cpu_oracle_us=t_cpu * 1e6, note="latency-bound: 2 kernels + 1.6 KB D2H; host API adds the 129 KB H2D")
    big = np.concatenate([ent] * 16)
    t_gpu = timeit(lambda: ctx.detect_block_motion(big, min_size=0.01, subdivide=16))
    t_cpu = timeit(lambda: oracle.detect_block_motion(big, min_size=0.01, subdivide=16))
    emit(case="detect_block_motion 128640 entries -> 160x160 (UI bounds)", gpu_host_api_us=t_gpu * 1e6, cpu_oracle_us=t_cpu * 1e6)

    # ---- fused Detection-tab frame: host frames -> entries + verdict
    t_gpu = timeit(lambda: ctx.frame_detect(prev, cur, 16, 16))
    t_cpu = timeit(lambda: oracle.detect_block_motion(oracle.block_match(prev, cur, 16, 16, 0, threads=oracle.max_threads(), fast=True)[2]), min_time=2)
    emit(case="frame_detect 1080p 16x16/+-16 (frames in host memory -> verdict)", gpu_us=t_gpu * 1e6, cpu_oracle_us=t_cpu * 1e6,
         cpu_threads=oracle.max_threads(), note="GPU time is dominated by the 4.1 MB pageable H2D copy of the two frames")

    # ---- Almeida estimator
    for name, (w, h) in (("2500 entries (reference test size)", (50, 50)), ("12600 entries (150x84, reference workloads)", (150, 84)),
                         ("2073600 entries (1080p dense, BASELINE config 2)", (1920, 1080))):
        field, q_truth = synth.rotation_field(w, h, 16 / 9, 22.275, (0.3, -0.2, 0.1))
        n = len(field)
        d = ctx.dev_alloc(field.nbytes)
        ctx.to_device(d, field)
        t_host = timeit(lambda: ctx.almeida(field, 16 / 9, 22.275))
        t_dev = timeit(lambda: ctx.almeida(None, 16 / 9, 22.275, d_entries=d, n=n))
        t_cpu = timeit(lambda: oracle.almeida_lsq_f32(field, 16 / 9, 22.275), min_time=1, min_iters=1)
        q = ctx.almeida(field, 16 / 9, 22.275)
        err = float(min(np.abs(q - q_truth).max(), np.abs(q + q_truth).max()))
        emit(case=f"almeida LSQ {name} (K5)", gpu_host_api_ms=t_host * 1e3, gpu_device_entries_ms=t_dev * 1e3,
             cpu_oracle_ms=t_cpu * 1e3, max_quat_err_vs_truth=err, algorithmic_gbs=16 * n / t_dev / 1e9,
             flop_model_tflops=30 * n * 700 / t_dev / 1e12)
        ctx.dev_free(d)
    field, q_truth = synth.rotation_field(150, 84, 16 / 9, 22.275, (0.5, 0.3, -0.4))
    bad = synth.corrupt_field(field, 0.2)
    t_gpu = timeit(lambda: ctx.almeida(bad, 16 / 9, 22.275, use_ransac=True, num_iters=200, seed=1))
    t_cpu = timeit(lambda: oracle.almeida_ransac_f32(bad, 16 / 9, 22.275, 200, 0.05, 1000, seed=1), min_time=1, min_iters=1)
    emit(case="almeida RANSAC 200x1000 on 12600 entries, 20% outliers (K6)", gpu_ms=t_gpu * 1e3, cpu_oracle_ms=t_cpu * 1e3,
         reference_published_ms="27.9-31.7 (Ryzen 9 3950X, docs/statistics/perf.csv)")

    # ---- densifier sort path: per-pixel field -> 150x84 grid (cv-decoder's case)
    field, _ = synth.rotation_field(1920, 1080, 16 / 9, 22.275, (0.3, -0.2, 0.1))
    t_gpu = timeit(lambda: ctx.densify(field, 150, 84))
    t_cpu = timeit(lambda: oracle.densify(field, 150, 84), min_time=1, min_iters=1)
    emit(case="densify 2073600 entries -> 150x84 (sort path, K3)", gpu_host_api_ms=t_gpu * 1e3, cpu_oracle_ms=t_cpu * 1e3,
         note="host API time includes the 33 MB H2D copy")
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bench_paths.jsonl"), "w") as f:
        for o in out:
            f.write(json.dumps(o) + "\n")


if __name__ == "__main__":
    main()
