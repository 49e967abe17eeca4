"""ncu target for the cv-decoder front-end kernels: a few launches of K7 / K8 / K9 on device-resident 4K inputs
(larger than nothing in particular: ncu serialises and flushes caches per replayed launch anyway).

    ncu --set full --clock-control none --import-source on -k regex:"frame_convert|contrast_mask|flow_cells|flow_emit|flow_pixels" \
        -o gpurun_out/r1_cv_front python tools/ncu_cv_front.py [W H]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ofps_b200 import capi


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
    ctx = capi.Context(0)
    npix = w * h
    rng = np.random.default_rng(1)
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    gray = np.full((h, w), 120, np.uint8)
    for _ in range(npix // 6000):
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        gray[y:y + int(rng.integers(2, 90)), x:x + int(rng.integers(2, 90))] = rng.integers(0, 256)
    flow = ((rng.random((h, w, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(6)).astype(np.float32)
    gw, gh = capi.mfield_size(w, h)
    d_bgr, d_gray, d_mask = ctx.dev_alloc(npix * 3), ctx.dev_alloc(npix), ctx.dev_alloc(npix)
    d_flow, d_ent, d_g2 = ctx.dev_alloc(npix * 8), ctx.dev_alloc(npix * 16), ctx.dev_alloc(npix)
    ctx.to_device(d_bgr, bgr)
    ctx.to_device(d_gray, gray)
    ctx.to_device(d_flow, flow)
    for _ in range(2):
        ctx.frame_convert_dev(d_bgr, w, h, 3 * w, 3, False, d_g2, w)
        ctx.contrast_mask_dev(d_gray, w, h, w, d_mask, w)
        n1 = ctx.flow_entries_dev(d_flow, 2 * w, d_mask, w, w, h, gw, gh, d_ent, gw * gh)
        n2 = ctx.flow_entries_dev(d_flow, 2 * w, d_mask, w, w, h, 0, 0, d_ent, npix)
    ctx.sync()
    print("entries", n1, n2, "launches", ctx.launch_count())
    ctx.close()


if __name__ == "__main__":
    main()
