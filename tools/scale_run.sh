#!/bin/bash
# 1/2/4/8-GPU runs of the headline bench (frames sharded) and the tiled large-frame benchmark.
mkdir -p gpurun_out
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/scale_bench_$n.json
  for wl in 8k 4k; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) tools/bench_tiled.py $wl 2>/dev/null | grep '^{' >> gpurun_out/scale_tiled.jsonl
  done
done
python - <<'PY'
import json,glob
for n in (1,2,4,8):
    try:
        d=json.load(open(f"gpurun_out/scale_bench_{n}.json"))
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "exhaustive", round(d["exhaustive"]["value"]))
    except Exception as e: print(n, "failed", e)
print(open("gpurun_out/scale_tiled.jsonl").read())
PY
