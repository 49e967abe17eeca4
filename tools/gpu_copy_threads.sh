#!/bin/bash
# host copy of the streaming decoder: non-temporal stores against plain memcpy, and OFPSB_COPY_THREADS
mkdir -p gpurun_out
for mode in nt plain nt plain; do
  if [ $mode = plain ]; then export OFPSB_COPY_PLAIN=1; else unset OFPSB_COPY_PLAIN; fi
  echo "mode $mode"; timeout 200 python tools/bench_stream.py 2>&1 | cut -c1-200
done
unset OFPSB_COPY_PLAIN
for n in 4 6 12; do
  export OFPSB_COPY_THREADS=$n
  echo "nt, copy threads $n"; timeout 200 python tools/bench_stream.py 2>&1 | grep pageable | cut -c1-200
done
( timeout 300 python -m pytest tests/test_gpu_stream.py -m gpu -x -q -p no:cacheprovider ) 2>&1 | tail -2
