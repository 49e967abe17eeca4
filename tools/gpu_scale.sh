#!/bin/bash
# multi-GPU session: the headline bench at every N the box has out of 8 / 2 (strong_c5 / tiled_8k / tiled_c4 objects
# included); stdout of the launch must be exactly one JSON line.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
lscpu | grep -E "^CPU\(s\)|Socket|NUMA node" ; nvidia-smi topo -m 2>/dev/null | head -11 | cut -c1-150
for n in ${@:-8 2}; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_ours_n$n.json 2> gpurun_out/bench_ours_n$n.err
    echo "stdout lines: $(wc -l < gpurun_out/bench_ours_n$n.json)"
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_ours_n$n.json'))
    print('N=$n value',round(d['value']),'e2e',round(d['e2e']['value']),'h2d GB/s/rank',round(d['e2e']['h2d_GBps_per_rank'],1),d['e2e']['numa'],'ceiling',d['e2e'].get('h2d_ceiling_GBps_per_rank'),'e2e_frame',round(d['e2e_frame']['value']))
    for k in ('strong_c5','tiled_8k','tiled_c4'):
        print('  ',k,d.get(k))
except Exception as e:
    print('N=$n failed',e)
PY
    tail -3 gpurun_out/bench_ours_n$n.err
  fi
done
