"""K1/K2 parity: the CUDA block matcher against the oracle's exhaustive search (bit-exact).

The reference has no SAD search; the oracle's orc_block_match IS the
specification, so these tests pin the kernel to the spec: integer motion vectors and costs
bit-exact, MotionEntry floats bit-exact (same operation order as av-decoder/src/lib.rs:404-419).
"""
import numpy as np
import pytest

from ofps_b200 import synth

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, prev, cur, block, search, metric):
    got = ctx.block_match(prev, cur, block, search, metric)
    mv, cost, ent = oracle.block_match(prev, cur, block, search, metric, threads=oracle.max_threads(), fast=False)
    assert got["n_blocks"] == (prev.shape[0] // block) * (prev.shape[1] // block)
    np.testing.assert_array_equal(got["cost"], cost)
    np.testing.assert_array_equal(got["mv"], mv)
    assert got["entries"].tobytes() == ent.tobytes()
    return got


@pytest.mark.parametrize("metric", [0, 1])
def test_c1_640x360_b16_r8(ctx, oracle, metric):
    """BASELINE config 0: 640x360, 16x16 blocks, +-8 (40x22 full blocks, 8 remainder rows ignored)."""
    prev, cur, truth = synth.make_pair(640, 360, 8, index=0)
    got = _check(ctx, oracle, prev, cur, 16, 8, metric)
    assert got["mv"].shape == (22, 40, 2)
    # most blocks follow the global pan: block matcher reports d = -(content motion)
    gx, gy = truth["global"]
    frac = np.mean((got["mv"][..., 0] == -gx) & (got["mv"][..., 1] == -gy))
    assert frac > 0.5


@pytest.mark.parametrize("block,search,w,h", [
    (16, 16, 640, 360), (16, 16, 1920, 1080), (8, 32, 512, 288), (8, 16, 320, 200), (8, 8, 328, 200),
    (16, 32, 400, 300), (16, 8, 1000, 100),
])
def test_tuned_instances(ctx, oracle, block, search, w, h):
    prev, cur, _ = synth.make_pair(w, h, search, index=3)
    _check(ctx, oracle, prev, cur, block, search, 0)


@pytest.mark.parametrize("block,search,w,h,metric", [
    (4, 3, 64, 48, 0), (12, 5, 200, 120, 1), (32, 7, 256, 160, 0), (16, 0, 128, 64, 0), (64, 2, 256, 128, 1),
    (16, 63, 300, 200, 0), (20, 9, 333, 127, 0),
])
def test_generic_geometries(ctx, oracle, block, search, w, h, metric):
    prev, cur, _ = synth.make_pair(w, h, max(search, 1), index=5)
    _check(ctx, oracle, prev, cur, block, search, metric)


def test_generic_kernel_equals_tuned(ctx):
    prev, cur, _ = synth.make_pair(1920, 1080, 16, index=2, noise_lsb=2)
    a = ctx.block_match(prev, cur, 16, 16, 0)
    ctx.set_option("block_match_kernel", 1)
    try:
        b = ctx.block_match(prev, cur, 16, 16, 0)
    finally:
        ctx.set_option("block_match_kernel", 0)
    for k in ("mv", "cost", "entries"):
        np.testing.assert_array_equal(a[k], b[k])


def test_ties_prefer_short_vectors(ctx, oracle):
    """Flat frames: every candidate has cost 0 -> the zero vector wins (tie-break on d^2)."""
    prev = np.full((96, 160), 77, np.uint8)
    cur = prev.copy()
    got = _check(ctx, oracle, prev, cur, 16, 16, 0)
    assert not got["mv"].any() and not got["cost"].any()
    # periodic texture: many exact matches, the nearest one must win
    x = (np.arange(160) % 8 * 30).astype(np.uint8)
    prev = np.tile(x, (96, 1))
    _check(ctx, oracle, prev, prev.copy(), 16, 16, 0)
    _check(ctx, oracle, prev, np.roll(prev, 3, axis=1), 8, 16, 0)


def test_noise_and_extremes(ctx, oracle):
    rng = np.random.default_rng(11)
    prev = rng.integers(0, 256, (144, 208), dtype=np.uint8)
    cur = rng.integers(0, 256, (144, 208), dtype=np.uint8)
    _check(ctx, oracle, prev, cur, 16, 16, 0)
    _check(ctx, oracle, prev, cur, 16, 16, 1)
    # maximum cost: all-0 vs all-255
    z = np.zeros((64, 64), np.uint8)
    f = np.full((64, 64), 255, np.uint8)
    got = _check(ctx, oracle, z, f, 16, 8, 0)
    assert (got["cost"] == 16 * 16 * 255).all()
    got = _check(ctx, oracle, z, f, 16, 8, 1)
    assert (got["cost"] == 16 * 16 * 255 * 255).all()


def test_small_and_empty_frames(ctx, oracle):
    prev, cur, _ = synth.make_pair(16, 16, 4, index=1, n_rects=0)
    _check(ctx, oracle, prev, cur, 16, 16, 0)     # a single block, only (0,0) is legal
    got = ctx.block_match(prev[:8], cur[:8], 16, 16, 0)   # no full block at all
    assert got["n_blocks"] == 0 and got["mv"].size == 0


def test_batch_and_stream_mode(ctx, oracle):
    from ofps_b200 import capi
    frames = np.stack([synth.textured_plane(100 + i, 320, 192) for i in range(6)])
    # independent pairs
    got = ctx.block_match(frames[:-1], frames[1:], 16, 8, 0)
    for i in range(5):
        mv, cost, ent = oracle.block_match(frames[i], frames[i + 1], 16, 8, 0)
        np.testing.assert_array_equal(got["mv"][i], mv)
        np.testing.assert_array_equal(got["cost"][i], cost)
    # stream mode: cur == prev + one frame, each frame uploaded once, small chunks
    ctx.set_option("batch_chunk_pairs", 2)
    try:
        nb = 12 * 20
        mv = np.empty((5, 12, 20, 2), np.int16)
        n = ctx.block_match_raw(frames.ctypes.data, frames.ctypes.data + 320 * 192, 320, 192, 320, 320 * 192, 5, 16, 8,
                                0, mv, None, None)
        assert n == nb
        np.testing.assert_array_equal(mv, got["mv"])
    finally:
        ctx.set_option("batch_chunk_pairs", 0)


def test_strided_rows(ctx, oracle):
    """Row stride larger than the width (and not a multiple of 4: the unaligned load path)."""
    import ctypes as C
    from ofps_b200 import capi
    prev, cur, _ = synth.make_pair(322, 100, 8, index=9)
    pad_p = np.zeros((100, 331), np.uint8)
    pad_c = np.zeros((100, 331), np.uint8)
    pad_p[:, :322], pad_c[:, :322] = prev, cur
    mv = np.empty((6, 20, 2), np.int16)
    cost = np.empty((6, 20), np.uint32)
    nb = C.c_size_t()
    capi.check(capi.lib().ofpsb_block_match(ctx._h, pad_p.ctypes.data, pad_c.ctypes.data, 322, 100, 331, 16, 8, 0,
                                            mv.ctypes.data, cost.ctypes.data, None, C.byref(nb)))
    omv, ocost, _ = oracle.block_match(prev, cur, 16, 8, 0)
    np.testing.assert_array_equal(mv, omv)
    np.testing.assert_array_equal(cost, ocost)


def test_invalid_arguments(ctx):
    from ofps_b200 import capi
    prev = np.zeros((64, 64), np.uint8)
    with pytest.raises(capi.OfpsError) as e:
        ctx.block_match(prev, prev, 16, 64, 0)      # range > 63
    assert e.value.code == capi.E_INVALID
    with pytest.raises(capi.OfpsError):
        ctx.block_match(prev, prev, 6, 4, 0)        # block not a multiple of 4
    with pytest.raises(capi.OfpsError):
        ctx.block_match(prev, prev, 16, 4, 7)       # unknown metric


@pytest.mark.parametrize("noise", [0, 1, 3])
@pytest.mark.parametrize("block,search,w,h", [(16, 16, 1920, 1080), (8, 32, 768, 432), (16, 8, 640, 360), (8, 8, 320, 208),
                                              (16, 32, 640, 368), (8, 16, 648, 360), (8, 32, 3840, 2160)])
def test_pruned_equals_exhaustive(ctx, oracle, block, search, w, h, noise):
    """The successive-elimination front end must not change a single output bit."""
    prev, cur, _ = synth.make_pair(w, h, search, index=11 + noise, noise_lsb=noise)
    ctx.set_option("block_match_stats", 1)
    try:
        a = ctx.block_match(prev, cur, block, search, 0)
        st = ctx.block_match_stats()
        ctx.set_option("block_match_prune", 0)
        b = ctx.block_match(prev, cur, block, search, 0)
    finally:
        ctx.set_option("block_match_prune", 1)
        ctx.set_option("block_match_stats", 0)
    for k in ("mv", "cost", "entries"):
        np.testing.assert_array_equal(a[k], b[k])
    if w % 16 == 0:     # the pruned path needs 16-byte aligned rows (TMA); otherwise it is skipped
        assert st["blocks"] == a["n_blocks"] and st["decided"] + st["worklist"] == st["blocks"]
    if noise == 0 and w % 16 == 0:
        assert st["decided"] > 0.5 * st["blocks"]       # noise-free synthetic motion is mostly decided by bounds
    mv, cost, _ = oracle.block_match(prev, cur, block, search, 0, threads=oracle.max_threads(), fast=True)
    np.testing.assert_array_equal(a["mv"], mv)
    np.testing.assert_array_equal(a["cost"], cost)


@pytest.mark.parametrize("block", [8, 16])
def test_wide_range_edge_cases(ctx, oracle, block):
    """+-32 SEA instances: frames smaller than the search window, periodic and flat content (ties over 65 x 65
    candidates), content panned by exactly the range and beyond it, saturated pixels, a batch."""
    cases = []
    a = synth.textured_plane(3, 64, 48)
    cases.append((a, np.roll(a, 5, axis=1)))                                    # 64 x 48: every block is a border block
    per = ((np.arange(160)[:, None] % 6) * 40 + (np.arange(256)[None] % 5) * 9).astype(np.uint8)
    cases.append((per, np.roll(np.roll(per, 2, axis=0), -4, axis=1)))           # many zero-cost candidates: the shortest wins
    cases.append((np.full((96, 160), 9, np.uint8), np.full((96, 160), 9, np.uint8)))
    t = synth.textured_plane(1, 384, 224)
    cases.append((t, np.roll(np.roll(t, 32, axis=1), -32, axis=0)))             # exactly dx = +R / dy = -R
    cases.append((t, np.roll(t, 45, axis=1)))                                   # beyond the range
    hi = np.random.default_rng(5).integers(0, 2, (128, 256)).astype(np.uint8) * 255
    cases.append((hi, 255 - hi))
    for prev, cur in cases:
        got = ctx.block_match(prev, cur, block, 32, 0)
        mv, cost, ent = oracle.block_match(prev, cur, block, 32, 0, threads=oracle.max_threads(), fast=True)
        np.testing.assert_array_equal(got["cost"], cost)
        np.testing.assert_array_equal(got["mv"], mv)
        assert got["entries"].tobytes() == ent.tobytes()
        assert np.abs(got["mv"]).max() <= 32
    fr = synth.make_stream(4, 272, 144, 32, noise_lsb=1)
    got = ctx.block_match(fr[:-1], fr[1:], block, 32, 0)
    for i in range(3):
        mv, cost, _ = oracle.block_match(fr[i], fr[i + 1], block, 32, 0, threads=oracle.max_threads(), fast=True)
        np.testing.assert_array_equal(got["cost"][i], cost)
        np.testing.assert_array_equal(got["mv"][i], mv)


def test_pruned_randomized(ctx):
    """Seeded random frame sizes, contents, batch sizes and geometries: the SEA path against the exhaustive kernel, bit for
    bit (the exhaustive kernel is checked against the oracle above)."""
    rng = np.random.default_rng(77)
    ctx.set_option("block_match_adaptive", 0)
    try:
        for case in range(24):
            block = int(rng.choice([8, 16]))
            search = int(rng.choice([8, 16, 32]))
            w = int(rng.integers(4, 60)) * 16
            h = int(rng.integers(3, 40)) * 8
            n = int(rng.integers(1, 4))
            frames = synth.make_stream(n + 1, w, h, search, noise_lsb=int(rng.integers(0, 4)))
            if rng.random() < 0.3:
                frames[-1] = rng.integers(0, 256, frames[-1].shape, dtype=np.uint8)      # a cut: one pair goes to the work list
            ctx.set_option("block_match_tile_h", int(rng.choice([0, 32, 64])))
            a = ctx.block_match(frames[:-1], frames[1:], block, search, 0)
            ctx.set_option("block_match_prune", 0)
            b = ctx.block_match(frames[:-1], frames[1:], block, search, 0)
            ctx.set_option("block_match_prune", 1)
            for k in ("mv", "cost"):
                np.testing.assert_array_equal(a[k], b[k], err_msg=f"case {case}: {w}x{h} b{block} r{search} n{n}")
            assert a["entries"].tobytes() == b["entries"].tobytes()
    finally:
        ctx.set_option("block_match_prune", 1)
        ctx.set_option("block_match_tile_h", 0)
        ctx.set_option("block_match_adaptive", 1)


def test_content_feedback_needs_two_bad_launches(ctx):
    """Content feedback of the SEA path (block_match_sea.cu): a launch that leaves most blocks to the exhaustive kernel
    is remembered; two in a row send the next 15 launches straight to the exhaustive kernel (1 kernel launch instead
    of 2), a single one — a scene cut in a frame-by-frame stream — does not.  Results are identical either way."""
    rng = np.random.default_rng(11)
    a, b = (rng.integers(0, 256, (360, 640), dtype=np.uint8) for _ in range(2))     # unrelated: nothing is decided by bounds
    good = synth.make_stream(3, 640, 360, 16)
    ctx.set_option("block_match_adaptive", 1)                                       # also clears the state

    def launches(prev, cur):
        n0 = ctx.launch_count()
        r = ctx.block_match(prev, cur, 16, 16, 0)
        ctx.sync()
        return ctx.launch_count() - n0, r

    ref_bad = None
    seen = []
    for i in range(6):
        n, r = launches(a, b)
        seen.append(n)
        if ref_bad is None:
            ref_bad = r
        assert r["entries"].tobytes() == ref_bad["entries"].tobytes()
    assert seen[:2] == [2, 2] and seen[2:] == [1, 1, 1, 1], seen                    # the third launch on is skipped
    ctx.set_option("block_match_adaptive", 1)
    a2 = np.roll(a, 3, axis=1)                                                      # the second scene pans: a good pair
    seen = [launches(*p)[0] for p in ((good[0], good[1]), (good[1], a), (a, a2), (a2, good[0]), (good[0], good[1]))]
    assert seen == [2, 2, 2, 2, 2], seen                                            # isolated bad pairs (cuts) never trigger it


def test_pruned_batch_and_worst_case(ctx, oracle):
    """Unrelated frames (nothing decided by the bounds) and a batch through the work list."""
    rng = np.random.default_rng(5)
    prev = rng.integers(0, 256, (3, 208, 336), dtype=np.uint8)
    cur = rng.integers(0, 256, (3, 208, 336), dtype=np.uint8)
    cur[1] = prev[1]                                    # one pair fully static
    got = ctx.block_match(prev, cur, 16, 16, 0)
    for i in range(3):
        mv, cost, ent = oracle.block_match(prev[i], cur[i], 16, 16, 0)
        np.testing.assert_array_equal(got["mv"][i], mv)
        np.testing.assert_array_equal(got["cost"][i], cost)
        assert got["entries"][i].tobytes() == ent.tobytes()
    assert not got["mv"][1].any()
