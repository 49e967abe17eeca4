import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ofps_b200 import capi, synth
ctx = capi.Context(0)
for w, h in ((30, 20), (150, 84)):
    field, _ = synth.rotation_field(w, h, 16 / 9, 22.275, (0.4, -0.1, 0.25))
    n = len(field)
    d = ctx.dev_alloc(field.nbytes)
    ctx.to_device(d, field)
    for v in (1, 0):
        ctx.set_option("almeida_cluster", v)
        for _ in range(3):
            ctx.almeida(None, 16 / 9, 22.275, d_entries=d, n=n)
