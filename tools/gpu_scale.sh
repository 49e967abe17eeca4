#!/bin/bash
# multi-GPU session: almeida re-check, then the headline bench at N = 8 and N = 2 (strong_c5 / tiled_8k objects included)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( timeout 300 python -m pytest tests/test_gpu_almeida.py -m gpu -x -q -p no:cacheprovider ) 2>&1 | tail -3
python tools/bench_paths.py 2>&1 | grep almeida | cut -c1-200
for n in 8 2; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_ours_n$n.json 2> gpurun_out/bench_ours_n$n.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_ours_n$n.json'))
    print('N=$n value',round(d['value']),'e2e',round(d['e2e']['value']),'h2d GB/s/rank',round(d['e2e']['h2d_GBps_per_rank'],1),d['e2e']['numa'],'e2e_frame',round(d['e2e_frame']['value']))
    print('  strong_c5',d.get('strong_c5'))
    print('  tiled_8k',d.get('tiled_8k'))
except Exception as e:
    print('N=$n failed',e)
PY
    tail -3 gpurun_out/bench_ours_n$n.err
  fi
done
