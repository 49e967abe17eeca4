#!/bin/bash
# One GPU-box session (1 GPU): smoke, the whole -m gpu suite, the headline bench and the reference arm.
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
tail -14 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cut -c1-3000 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
if [ "$1" != "quick" ]; then
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  cut -c1-600 gpurun_out/bench_reference.json
fi
if [ "$2" = "sanitize" ]; then
  timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_k1.py > gpurun_out/sanitize_memcheck.txt 2>&1; tail -3 gpurun_out/sanitize_memcheck.txt
  timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_k1.py > gpurun_out/sanitize_racecheck.txt 2>&1; tail -3 gpurun_out/sanitize_racecheck.txt
fi
