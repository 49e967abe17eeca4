#!/bin/bash
# multi-GPU session: tiled 8K pair at 1 and N GPUs (N = number of visible GPUs)
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
python tools/bench_tiled_peer.py > gpurun_out/tiled_peer_1.json 2> gpurun_out/tiled_peer_1.err; cat gpurun_out/tiled_peer_1.json; tail -3 gpurun_out/tiled_peer_1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/bench_tiled_peer.py > gpurun_out/tiled_peer_$n.json 2> gpurun_out/tiled_peer_$n.err
    cat gpurun_out/tiled_peer_$n.json; tail -3 gpurun_out/tiled_peer_$n.err
  fi
done
