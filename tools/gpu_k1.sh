#!/bin/bash
# One GPU-box session for K1 work: block-matching parity tests, the K1 bench, a launch list and a full ncu capture.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_block_match.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/pytest_k1.log 2>&1
tail -5 gpurun_out/pytest_k1.log
timeout 600 python tools/bench_k1.py > gpurun_out/bench_k1.jsonl 2> gpurun_out/bench_k1.err
cat gpurun_out/bench_k1.jsonl; tail -3 gpurun_out/bench_k1.err
if [ "$1" != "quick" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 4 -c 8 --csv \
      --log-file gpurun_out/r2_launches_k1.csv python tools/ncu_target.py 64 > /dev/null 2>&1
  cut -d, -f5,12- gpurun_out/r2_launches_k1.csv | tail -26
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sea_kernel|block_match_list" -s 2 -c 2 \
      -f -o gpurun_out/r2_k1 python tools/ncu_target.py 64 > gpurun_out/ncu_k1.log 2>&1
  tail -3 gpurun_out/ncu_k1.log
fi
