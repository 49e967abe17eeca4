#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_almeida.py tests/test_gpu_densify_detect.py tests/test_gpu_stream.py -m gpu -x -q -p no:cacheprovider ) 2>&1 | tail -4
python tools/bench_paths.py > gpurun_out/bench_paths.log 2>&1; cat gpurun_out/bench_paths.jsonl | cut -c1-400
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size \
   --clock-control none --csv --log-file gpurun_out/r2_launches_paths.csv python tools/ncu_paths.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"densify_scan|seg_|detect_kernel|almeida_lsq|ransac_" -c 40 \
   -f -o gpurun_out/r2_paths python tools/ncu_paths.py > gpurun_out/ncu_paths.log 2>&1
tail -2 gpurun_out/ncu_paths.log
