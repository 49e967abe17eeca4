"""Short target for `ncu --set full` of the secondary kernels: K3 densify (scan and sort paths), K4 detect, K5 Almeida
LSQ (single CTA / persistent grid), K6 RANSAC."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ofps_b200 import capi, synth

ctx = capi.Context(0)
prev, cur, _ = synth.make_pair(1920, 1080, 16, index=0)
ent = ctx.block_match(prev, cur, 16, 16, 0)["entries"].reshape(-1, 4)
for _ in range(2):
    ctx.detect_block_motion(ent)                                   # scan densify 8040 -> 14x14 + detect
    ctx.detect_block_motion(np.concatenate([ent] * 16), min_size=0.01, subdivide=16)   # 128640 -> 160x160
field, _ = synth.rotation_field(1920, 1080, 16 / 9, 22.275, (0.3, -0.2, 0.1))
for _ in range(2):
    ctx.densify(field, 150, 84)                                    # sort path, 2.07 M entries
    ctx.almeida(field, 16 / 9, 22.275)                             # persistent grid, 2.07 M entries
small, _ = synth.rotation_field(150, 84, 16 / 9, 22.275, (0.5, 0.3, -0.4))
bad = synth.corrupt_field(small, 0.2)
for _ in range(2):
    ctx.almeida(small, 16 / 9, 22.275)                             # persistent grid, 12,600 entries
    ctx.almeida(small[:2000], 16 / 9, 22.275)                      # single CTA
    ctx.almeida(bad, 16 / 9, 22.275, use_ransac=True, num_iters=200, seed=1)
ctx.sync()
print("done")
