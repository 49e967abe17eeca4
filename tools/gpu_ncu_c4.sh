#!/bin/bash
# ncu --set full of the +-32 SEA kernel on BASELINE config 4 (4K, 8x8/+-32), 4 pairs per launch.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sea_kernel" -s 2 -c 1 \
    -f -o gpurun_out/r2_c4 python tools/ncu_target.py 4 3840 2160 8 32 > gpurun_out/ncu_c4.log 2>&1
tail -3 gpurun_out/ncu_c4.log
