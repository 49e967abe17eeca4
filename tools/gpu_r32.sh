#!/bin/bash
# One GPU-box session for the +-32 SEA instances: parity tests, then BASELINE config 4 (4K, 8x8/+-32) and 1080p 16x16/+-32
# through the SEA path, the round-1 pipeline and the exhaustive kernel.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_block_match.py tests/test_gpu_tiled.py tests/test_gpu_tiled_peer.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/pytest_r32.log 2>&1
tail -2 gpurun_out/pytest_r32.log
PAIRS=8 STEPS=5 timeout 600 python tools/bench_k1.py 3840 2160 8 32 > gpurun_out/bench_r32_c4.jsonl 2> gpurun_out/bench_r32.err
PAIRS=16 STEPS=5 timeout 600 python tools/bench_k1.py 1920 1080 16 32 > gpurun_out/bench_r32_1080.jsonl 2>> gpurun_out/bench_r32.err
tail -3 gpurun_out/bench_r32.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_r32_c4.jsonl','gpurun_out/bench_r32_1080.jsonl'):
    for l in open(f):
        d=json.loads(l)
        if 'noise_lsb' in d:
            print(d['block'],d['search'],d['noise_lsb'],'exh %.1f sea %.1f r1 %.1f'%(d['exhaustive_us_per_pair'],d['sea']['us_per_pair'],d['r1_pipeline']['us_per_pair']),d['sea']['bit_equal_to_exhaustive'],d['sea']['stats'],{k:{a:round(b,1) for a,b in v.items()} for k,v in d.items() if k=='tile_h_0'})
PY
