#!/bin/bash
# One GPU-box session: the whole -m gpu suite, the headline bench (+ reference arm), the secondary benches and the
# ncu captures of the cv-front kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time timeout 720 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python tools/bench_cv_front.py > gpurun_out/bench_cv_front.log 2>&1
tail -3 gpurun_out/bench_cv_front.log | cut -c1-300
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cut -c1-400 gpurun_out/bench_ours.json
if [ "$1" != "quick" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on \
      -k regex:"frame_convert|contrast_mask|flow_cells|flow_emit|flow_pixels|cell_bounds|tile_scan" -c 8 \
      -f -o gpurun_out/r1_cv_front python tools/ncu_cv_front.py > gpurun_out/ncu_cv_front.log 2>&1
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/r1_launches_cv_front.csv python tools/ncu_cv_front.py 1920 1080 > /dev/null 2>&1
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  python tools/bench_paths.py > gpurun_out/bench_paths.log 2>&1
fi
ls -la gpurun_out
