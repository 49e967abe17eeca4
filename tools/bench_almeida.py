"""K5 latency (B200): one LSQ estimate with the entries already on the device, one-cluster solver against the multi-CTA
grid, for the reference's field sizes.  Wall time of the synchronous C-ABI call (it returns the quaternion)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ofps_b200 import capi, synth

ctx = capi.Context(0)


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n


for w, h in ((30, 20), (50, 50), (128, 64), (150, 84), (128, 128)):
    field, _ = synth.rotation_field(w, h, 16 / 9, 22.275, (0.4, -0.1, 0.25))
    n = len(field)
    d = ctx.dev_alloc(field.nbytes)
    ctx.to_device(d, field)
    res = {"entries": n}
    for name, v in (("cluster_us", 1), ("grid_us", 0)):
        ctx.set_option("almeida_cluster", v)
        res[name] = 1e6 * timeit(lambda: ctx.almeida(None, 16 / 9, 22.275, d_entries=d, n=n))
    ctx.set_option("almeida_cluster", 1)
    print(json.dumps(res), flush=True)
    ctx.dev_free(d)
