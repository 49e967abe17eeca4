"""Streaming decoder entry points (ofpsb_stream_*): frame by frame == the batch path == the oracle."""
import numpy as np
import pytest

from ofps_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,block,search,metric,depth", [(640, 360, 16, 8, 0, 3), (1920, 1080, 16, 16, 0, 4),
                                                           (322, 200, 8, 8, 1, 5), (648, 364, 8, 16, 0, 4), (768, 432, 8, 32, 0, 4),
                                                           (1280, 720, 16, 32, 0, 3)])
def test_stream_push_equals_batch(ctx, oracle, w, h, block, search, metric, depth):
    frames = synth.make_stream(9, w, h, search, noise_lsb=1)
    st = capi.FrameStream(ctx, w, h, block, search, metric, depth)
    try:
        assert st.push(frames[0]) is None                 # Ok(false): the first frame has no predecessor
        got = [st.push(frames[i]).copy() for i in range(1, len(frames))]
    finally:
        st.close()
    for i, g in enumerate(got):
        _, _, ent = oracle.block_match(frames[i], frames[i + 1], block, search, metric, threads=oracle.max_threads(), fast=metric == 0)
        assert g.tobytes() == ent.tobytes(), i


def test_stream_pipelined_and_pageable_views(ctx):
    """submit runs ahead of collect; strided (non-contiguous rows) and pinned inputs take the other upload paths."""
    w, h = 1280, 720
    frames = synth.make_stream(12, w, h, 16)
    want = ctx.block_match(frames[:-1], frames[1:], 16, 16, 0, want=("entries",))["entries"].reshape(11, -1, 4)
    st = capi.FrameStream(ctx, w, h, 16, 16, 0, depth=5)
    try:
        padded = np.zeros((12, h, w + 48), np.uint8)
        padded[:, :, :w] = frames
        pin = capi.PinnedArray((h, w), np.uint8)
        got = []
        st.submit(padded[0, :, :w])
        for i in range(1, 12):
            if i % 3 == 0:
                pin.array[:] = frames[i]
                st.submit(pin.array)
                r = st.collect()                           # pinned buffer is reused: drain before overwriting it
                if r is not None:
                    got.append(r.copy())
                while True:
                    r = st.collect()
                    if r is None:
                        break
                    got.append(r.copy())
                ctx.sync()
            else:
                st.submit(padded[i, :, :w])
            if i >= 2 and len(got) < i - 1:
                got.append(st.collect().copy())
        while True:
            r = st.collect()
            if r is None:
                break
            got.append(r.copy())
        with pytest.raises(capi.OfpsError):
            for i in range(8):
                st.submit(frames[i])                        # too many outstanding results
    finally:
        st.close()
    assert len(got) == 11
    for i in range(11):
        assert got[i].tobytes() == want[i].tobytes(), i
