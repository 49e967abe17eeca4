"""Spatial tiling behind the C ABI (ofpsb_tiled_*): strips that read their halo rows from the neighbours' memory inside
the matching kernel.  On the 1-GPU box every "rank" is a strip object on the same device joined with
ofpsb_tiled_connect_local — the same kernels, tensor maps and flags as across GPUs; the IPC / NVLink variant is
exercised by bench.py --gpus N (tiled_8k) and tests/test_gpu_tiled.py::test_peer_tiled_over_ipc on multi-GPU boxes."""
import numpy as np
import pytest

from ofps_b200 import capi, synth

pytestmark = pytest.mark.gpu


def run_tiled(ctx, prev, cur, block, search, world, repeats=3, wait=True):
    h, w = prev.shape
    ts = [capi.Tiled(ctx, r, world, w, h, block, search, 2) for r in range(world)]
    try:
        for r, t in enumerate(ts):
            t.connect_local(ts[r - 1] if r > 0 else None, ts[r + 1] if r + 1 < world else None)
        outs = []
        for t in ts:
            t.upload(0, prev[t.y0:t.y0 + t.own_rows])
            t.upload(1, cur[t.y0:t.y0 + t.own_rows])
            t.publish(0)
            t.publish(1)
        bufs = [(ctx.dev_alloc(t.n_blocks * 16), ctx.dev_alloc(t.n_blocks * 4), ctx.dev_alloc(t.n_blocks * 4)) for t in ts]
        for _ in range(repeats):                  # eager, capture, graph replay
            for t, (de, dm, dc) in zip(ts, bufs):
                t.match(0, 1, de, dm, dc, wait=wait)
        ctx.sync()
        ent, mv, cost = [], [], []
        for t, (de, dm, dc) in zip(ts, bufs):
            e = np.empty((t.n_blocks, 4), np.float32); ctx.to_host(e, de)
            m = np.empty((t.n_blocks, 2), np.int16); ctx.to_host(m, dm)
            c = np.empty(t.n_blocks, np.uint32); ctx.to_host(c, dc)
            ent.append(e); mv.append(m); cost.append(c)
            for p in (de, dm, dc):
                ctx.dev_free(p)
        return np.concatenate(ent), np.concatenate(mv), np.concatenate(cost)
    finally:
        for t in ts:
            t.close()


@pytest.mark.parametrize("w,h,block,search,world,noise", [
    (640, 360, 16, 8, 2, 0), (640, 360, 16, 16, 3, 2), (1920, 1080, 16, 16, 8, 0), (1920, 1080, 16, 16, 5, 2),
    (648, 360, 8, 16, 4, 1), (512, 288, 8, 8, 6, 0), (768, 432, 8, 32, 3, 0), (400, 304, 16, 32, 2, 1),
    (1280, 720, 16, 16, 1, 0),
])
def test_peer_tiled_equals_whole_frame(ctx, oracle, w, h, block, search, world, noise):
    prev, cur, _ = synth.make_pair(w, h, search, index=21 + world, noise_lsb=noise)
    ent, mv, cost = run_tiled(ctx, prev, cur, block, search, world)
    whole = ctx.block_match(prev, cur, block, search, 0)
    np.testing.assert_array_equal(mv, whole["mv"].reshape(-1, 2))
    np.testing.assert_array_equal(cost, whole["cost"].reshape(-1))
    assert ent.tobytes() == whole["entries"].tobytes()
    omv, ocost, _ = oracle.block_match(prev, cur, block, search, 0, threads=oracle.max_threads(), fast=True)
    np.testing.assert_array_equal(mv, omv.reshape(-1, 2))


def test_peer_tiled_8k_strips(ctx):
    """North-star geometry: 7680x4320, 16x16/+-16, eight strips (34/34/34/34/34/34/33/33 block rows)."""
    prev, cur, _ = synth.make_pair(7680, 4320, 16, index=3)
    ent, mv, cost = run_tiled(ctx, prev, cur, 16, 16, 8)
    whole = ctx.block_match(prev, cur, 16, 16, 0)
    np.testing.assert_array_equal(mv, whole["mv"].reshape(-1, 2))
    np.testing.assert_array_equal(cost, whole["cost"].reshape(-1))
    assert ent.tobytes() == whole["entries"].tobytes()


def test_peer_tiled_motion_across_seams(ctx, oracle):
    """Vertical pan of exactly +-range: every seam block's match lies wholly inside the neighbour's rows."""
    base = synth.textured_plane(77, 640, 392)
    for dy in (16, -16, 9):
        prev, cur = base[16:376], base[16 + dy:376 + dy]            # cur[y] = prev[y + dy]
        prev, cur = np.ascontiguousarray(prev), np.ascontiguousarray(cur)
        ent, mv, cost = run_tiled(ctx, prev, cur, 16, 16, 4)
        omv, ocost, _ = oracle.block_match(prev, cur, 16, 16, 0, threads=oracle.max_threads(), fast=True)
        np.testing.assert_array_equal(mv, omv.reshape(-1, 2))
        np.testing.assert_array_equal(cost, ocost.reshape(-1))
        inner = mv.reshape(22, 40, 2)[2:-2]
        assert (inner[..., 1] == dy).all() and (inner[..., 0] == 0).all()


def test_tiled_argument_checks(ctx):
    with pytest.raises(capi.OfpsError):
        capi.Tiled(ctx, 0, 4, 64, 48, 16, 8)            # 3 block rows cannot feed 4 ranks
    t = capi.Tiled(ctx, 0, 2, 640, 360, 16, 8)
    try:
        with pytest.raises(capi.OfpsError):
            t.connect_local(None, None)                   # rank 0 of 2 needs a lower neighbour
        with pytest.raises(capi.OfpsError):
            t.connect(None, t.export())                   # a handle of this process / of the wrong rank
        assert t.y0 == 0 and t.rows == 176 and t.own_rows == 176 and (t.nbx, t.nby) == (40, 11)
    finally:
        t.close()


def test_peer_tiled_stream_batch(ctx):
    """ofpsb_tiled_match_stream: consecutive slots in one launch == pair by pair == whole frames."""
    w, h, block, search, world, n = 640, 360, 16, 16, 3, 4
    frames = synth.make_stream(n + 1, w, h, search, noise_lsb=1)
    whole = ctx.block_match(frames[:-1], frames[1:], block, search, 0, want=("entries",))["entries"].reshape(n, -1, 4)
    ts = [capi.Tiled(ctx, r, world, w, h, block, search, n + 1) for r in range(world)]
    try:
        for r, t in enumerate(ts):
            t.connect_local(ts[r - 1] if r > 0 else None, ts[r + 1] if r + 1 < world else None)
        for t in ts:
            for s in range(n + 1):
                t.upload(s, frames[s, t.y0:t.y0 + t.own_rows])
                t.publish(s)
        parts = []
        for t in ts:
            de = ctx.dev_alloc(n * t.n_blocks * 16)
            for _ in range(3):
                t.match_stream(0, n, de)
            ctx.sync()
            e = np.empty((n, t.n_blocks, 4), np.float32)
            ctx.to_host(e, de)
            ctx.dev_free(de)
            parts.append(e)
        got = np.concatenate(parts, axis=1)
        assert got.tobytes() == whole.tobytes()
    finally:
        for t in ts:
            t.close()
