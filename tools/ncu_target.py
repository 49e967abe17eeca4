"""Short target for `ncu --set full`: a few launches of the headline block-matching kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ofps_b200 import capi, synth

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w, h, block, search = (int(a) for a in sys.argv[2:6]) if len(sys.argv) > 5 else (1920, 1080, 16, 16)
ctx = capi.Context(0)
frames = synth.make_stream(pairs + 1, w, h, search)
fb = w * h
d = ctx.dev_alloc(frames.nbytes)
de = ctx.dev_alloc(pairs * (w // block) * (h // block) * 16)
ctx.to_device(d, frames)
for _ in range(4):
    ctx.block_match_dev(d, d + fb, w, h, w, fb, pairs, block, search, 0, None, None, de)
ctx.sync()
print("done")
