"""Target for compute-sanitizer (memcheck / racecheck / initcheck-free paths) over the cv-front kernels: every
kernel and both of its staging paths (16-byte asynchronous copies / element loads) on small, odd-sized inputs,
border tiles, interior tiles and the chunked wide-cell case; results are compared with the oracle as they go.

    compute-sanitizer --tool memcheck  python tools/sanitize_cv_front.py
    compute-sanitizer --tool racecheck python tools/sanitize_cv_front.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from ofps_b200 import capi


def main():
    oracle.build()
    ctx = capi.Context(0)
    rng = np.random.default_rng(3)
    checks = 0
    for (w, h) in ((800, 96), (257, 33), (31, 9), (1, 1), (6, 40), (640, 360)):
        bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        gray, rgba = ctx.frame_convert(bgr, want_rgba=True)
        assert np.array_equal(gray, oracle.bgr_to_gray(bgr)) and np.array_equal(rgba, oracle.bgr_to_rgba(bgr))
        g = np.full((h, w), 120, np.uint8)
        for _ in range(max(4, w * h // 3000)):
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            g[y:y + int(rng.integers(1, 20)), x:x + int(rng.integers(1, 20))] = rng.integers(0, 256)
        mask = ctx.contrast_mask(g)
        assert np.array_equal(mask, oracle.contrast_mask(g))
        flow = ((rng.random((h, w, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(6)).astype(np.float32)
        for grid in ((0, 0), capi.mfield_size(w, h), (3, 2)):
            for m in (mask, None):
                got = ctx.flow_entries(flow, m, *grid)
                assert got.tobytes() == oracle.flow_entries(flow, m, *grid).tobytes()
                checks += 1
        got = ctx.cv_flow_frame(g, flow, True, *capi.mfield_size(w, h))
        assert got.tobytes() == oracle.flow_entries(flow, oracle.contrast_mask(g), *capi.mfield_size(w, h)).tobytes()
        if w >= 150:
            dw, dh = capi.mfield_size(w, h)
            assert np.array_equal(ctx.frame_resize(bgr, dw, dh), oracle.resize_linear(bgr, dw, dh))
    # wide cells: rows walked in chunks; unaligned device pointers: the element-load staging paths
    flow = ((rng.random((6, 9000, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(8)).astype(np.float32)
    mask = (rng.random((6, 9000)) < 0.4).astype(np.uint8) * 255
    for grid in ((2, 2), (40, 3)):
        assert ctx.flow_entries(flow, mask, *grid).tobytes() == oracle.flow_entries(flow, mask, *grid).tobytes()
    w, h = 301, 40
    g = rng.integers(0, 256, (h, w), dtype=np.uint8)
    d_g, d_m = ctx.dev_alloc(w * h + 64), ctx.dev_alloc(w * h + 64)
    buf = np.zeros(w * h + 64, np.uint8)
    buf[3:3 + w * h] = g.ravel()
    ctx.to_device(d_g, buf)
    ctx.contrast_mask_dev(d_g + 3, w, h, w, d_m + 5, w)
    out = np.zeros(w * h + 64, np.uint8)
    ctx.to_host(out, d_m)
    ctx.sync()
    assert np.array_equal(out[5:5 + w * h].reshape(h, w), oracle.contrast_mask(g))
    ctx.dev_free(d_g)
    ctx.dev_free(d_m)
    print(f"sanitize_cv_front: ok, {checks} flow cases, {ctx.launch_count()} launches")
    ctx.close()


if __name__ == "__main__":
    main()
