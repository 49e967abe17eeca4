"""Where the per-frame time of the streaming decoder goes (host side): submit / collect separately, pageable vs pinned frames."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ofps_b200 import capi, synth

W, H = 1920, 1080
ctx = capi.Context(0)
frames = synth.make_stream(33, W, H, 16)
pin = capi.PinnedArray(frames.shape, np.uint8)
pin.array[:] = frames
out = np.empty(((W // 16) * (H // 16), 4), np.float32)
for name, src in (("pageable", frames), ("pinned", pin.array)):
    st = capi.FrameStream(ctx, W, H, 16, 16, 0, depth=6)
    for rep in range(3):
        ts, tc = 0.0, 0.0
        t_all = time.perf_counter()
        for i in range(len(src)):
            t0 = time.perf_counter()
            st.submit(src[i])
            t1 = time.perf_counter()
            if i >= 3:
                st.collect(out)
            t2 = time.perf_counter()
            ts += t1 - t0
            tc += t2 - t1
        while st.collect(out) is not None:
            pass
        t_all = time.perf_counter() - t_all
    # synchronous form: ofpsb_stream_push returns the result of the frame it was given
    st2 = capi.FrameStream(ctx, W, H, 16, 16, 0, depth=4)
    for rep in range(3):
        t_sync = time.perf_counter()
        for i in range(len(src)):
            st2.push(src[i])
        t_sync = time.perf_counter() - t_sync
    st2.close()
    n = len(src)
    print(json.dumps({"frames": name, "sync_push_us": 1e6 * t_sync / n, "submit_us": 1e6 * ts / n, "collect_us": 1e6 * tc / n, "per_frame_us": 1e6 * t_all / n,
                      "Gpix_s": W * H * n / t_all / 1e9}), flush=True)
    st.close()
