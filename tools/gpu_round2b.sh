#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step')})
print('e2e',round(d['e2e']['value']),'e2e_frame',round(d['e2e_frame']['value']),d['e2e_frame']['sync_push'])
print('roofline frac',d['roofline']['frac'],'dominant',d['roofline']['dominant_kernel'])
print('noisy',round(d['noisy']['value']),'exh',round(d['exhaustive']['value']))
print('cpu',d.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cut -c1-400 gpurun_out/bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size \
   --clock-control none --csv --log-file gpurun_out/r2_launches_paths.csv python tools/ncu_paths.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"densify_scan|detect_|almeida_lsq|ransac_" -c 30 \
   -f -o gpurun_out/r2_paths python tools/ncu_paths.py > gpurun_out/ncu_paths.log 2>&1
tail -2 gpurun_out/ncu_paths.log
