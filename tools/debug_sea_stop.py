"""Phase timing of the SEA kernel by early exit (OFPSB_DEBUG_SEA_STOP = 1 loads, 2 + window sums, 3 + tile predictor, 0 whole
kernel); PAIRS=1 gives the latency of a launch that does not fill the machine."""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ofps_b200 import capi, synth
ctx = capi.Context(0)
W,H,P=1920,1080,int(os.environ.get('PAIRS','64'))
fr = synth.make_stream(P+1, W, H, 16)
d = ctx.dev_alloc(fr.nbytes); de = ctx.dev_alloc(P*8040*16)
ctx.to_device(d, fr)
ctx.set_option("block_match_profile", 1)
ctx.set_option("block_match_adaptive", 0)
for _ in range(4):
    ctx.block_match_dev(d, d+W*H, W, H, W, W*H, P, 16, 16, 0, None, None, de)
    k = ctx.block_match_kernel_ms()
print(os.environ.get("OFPSB_DEBUG_SEA_STOP","0"), "sea_us_per_pair", round(k[0]*1e3/P,2), "list", round(k[1]*1e3/P,2))
