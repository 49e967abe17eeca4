"""CPU: the fused SEA block-matching kernel (ofps_b200/csrc/block_match_sea.cu) executed on the thread-per-CUDA-thread
stand-in of tests/emu/cuda_emu.h (TMA box loads replaced by plain copies; everything after them is the product code)
and compared with the oracle's exhaustive search.  Checks kernel LOGIC — window sums, four-term bounds, tie-breaks,
predictors, frame borders, strips — on the GPU-less build container; the parity tests proper are the `-m gpu` ones in
tests/test_gpu_block_match.py.  Test infrastructure only: nothing here is a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ofps_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
LIB = os.path.join(EMU, "_build", "libemu_block_match.so")
CSRC = os.path.join(HERE, "..", "ofps_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU, "emu_block_match.cpp"), os.path.join(EMU, "cuda_emu.h"),
            os.path.join(CSRC, "block_match_sea.cu"), os.path.join(CSRC, "block_match_common.cuh"),
            os.path.join(CSRC, "common.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-I" + EMU, "-o", LIB, srcs[0]])
    L = C.CDLL(LIB)
    L.emu_block_match_sea.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return L


def run_sea(L, prev, cur, block, search, tile_h=64, strip=None):
    """prev/cur: [h, w] or [n, h, w] u8.  strip = (y0, rows): match only cur rows y0 .. y0+rows with the halo rows
    of prev that exist in the frame.  Returns mv, cost, entries, resolved mask, stats."""
    if prev.ndim == 2:
        prev, cur = prev[None], cur[None]
    prev, cur = np.ascontiguousarray(prev), np.ascontiguousarray(cur)
    n, h, w = prev.shape
    y0, rows = (0, h) if strip is None else strip
    top, bot = min(search, y0), min(search, h - y0 - rows)
    nbx, nby = w // block, rows // block
    mv = np.full((n, nby, nbx, 2), -99, np.int16)
    cost = np.full((n, nby, nbx), 0xDEADBEEF, np.uint32)
    ent = np.zeros((n, nby, nbx, 4), np.float32)
    wl = np.zeros(max(n * nby * nbx, 1), np.uint32)
    cnt = np.zeros(1, np.uint32)
    stats = np.zeros(4, np.uint64)
    rc = L.emu_block_match_sea(prev.ctypes.data + y0 * w, cur.ctypes.data + y0 * w, w, rows, w, h * w, n, top, bot, y0, h,
                               block, search, mv.ctypes.data, cost.ctypes.data, ent.ctypes.data, wl.ctypes.data,
                               cnt.ctypes.data, stats.ctypes.data, tile_h)
    assert rc == 0
    resolved = np.ones(n * nby * nbx, bool)
    resolved[wl[:cnt[0]]] = False
    assert len(set(wl[:cnt[0]].tolist())) == cnt[0]
    return mv, cost, ent, resolved.reshape(n, nby, nbx), stats


def check(L, oracle, prev, cur, block, search, min_resolved=0.0, **kw):
    mv, cost, ent, res, stats = run_sea(L, prev, cur, block, search, **kw)
    if prev.ndim == 2:
        prev, cur = prev[None], cur[None]
    for i in range(len(prev)):
        omv, ocost, oent = oracle.block_match(prev[i], cur[i], block, search, 0, threads=oracle.max_threads(), fast=True)
        r = res[i]
        np.testing.assert_array_equal(cost[i][r], ocost[r])
        np.testing.assert_array_equal(mv[i][r], omv[r])
        assert ent[i][r].tobytes() == oent.reshape(ent[i].shape)[r].tobytes()
        assert (cost[i][~r] == 0xDEADBEEF).all()            # undecided blocks are left to the exhaustive kernel
    assert stats[0] == res.size and stats[1] == res.sum()
    assert res.mean() >= min_resolved, (res.mean(), stats)
    return res.mean(), stats


@pytest.mark.parametrize("block,search,w,h,noise", [
    (16, 16, 640, 360, 0), (16, 16, 640, 360, 2), (16, 8, 640, 360, 0), (16, 8, 400, 200, 3),
    (8, 16, 320, 200, 0), (8, 16, 328, 136, 2), (8, 8, 328, 200, 0), (8, 8, 264, 72, 1),
])
def test_sea_matches_oracle(emu, oracle, block, search, w, h, noise):
    prev, cur, _ = synth.make_pair(w, h, search, index=3, noise_lsb=noise)
    frac, stats = check(emu, oracle, prev, cur, block, search, min_resolved=0.5)
    print(f"b{block} r{search} {w}x{h} noise {noise}: resolved {frac:.3f}, exact evals/block {stats[2] / stats[0]:.2f}, "
          f"full scans {stats[3] / stats[0]:.2f}")


@pytest.mark.parametrize("tile_h", [64, 32])
def test_sea_stream_batch(emu, oracle, tile_h):
    fr = synth.make_stream(3, 384, 208, 16)
    check(emu, oracle, fr[:-1], fr[1:], 16, 16, min_resolved=0.6, tile_h=tile_h)


@pytest.mark.parametrize("block,search,w,h,noise", [(16, 16, 640, 360, 1), (16, 8, 400, 200, 0), (8, 16, 328, 136, 2), (8, 8, 264, 72, 0)])
def test_sea_half_height_tiles(emu, oracle, block, search, w, h, noise):
    """The 32-row tile instances (launches that do not fill the machine)."""
    prev, cur, _ = synth.make_pair(w, h, search, index=5, noise_lsb=noise)
    check(emu, oracle, prev, cur, block, search, min_resolved=0.5, tile_h=32)


@pytest.mark.parametrize("block,search,w,h,noise", [
    (8, 32, 328, 136, 0), (8, 32, 328, 136, 2), (16, 32, 384, 208, 0), (16, 32, 400, 176, 3), (8, 32, 136, 72, 1),
])
def test_sea_wide_range(emu, oracle, block, search, w, h, noise):
    """+-32 (BASELINE config 4): two groups of 32 dx columns plus the dx = +32 column, two-pass full scan."""
    prev, cur, _ = synth.make_pair(w, h, search, index=9, noise_lsb=noise)
    frac, stats = check(emu, oracle, prev, cur, block, search, min_resolved=0.4)
    print(f"b{block} r{search} {w}x{h} noise {noise}: resolved {frac:.3f}, exact evals/block {stats[2] / stats[0]:.2f}, "
          f"full scans {stats[3] / stats[0]:.2f}")


def test_sea_wide_range_ties_and_extremes(emu, oracle):
    prev = ((np.arange(160)[:, None] % 6) * 40 + (np.arange(256)[None] % 5) * 9).astype(np.uint8)
    check(emu, oracle, prev, np.roll(np.roll(prev, 2, axis=0), -4, axis=1), 8, 32)
    check(emu, oracle, prev, np.roll(np.roll(prev, 2, axis=0), -4, axis=1), 16, 32)
    check(emu, oracle, np.full((96, 160), 9, np.uint8), np.full((96, 160), 9, np.uint8), 8, 32, min_resolved=1.0)
    a, b = synth.textured_plane(1, 256, 128), synth.textured_plane(2, 256, 128)
    check(emu, oracle, a, b, 8, 32)
    check(emu, oracle, a, np.roll(a, 40, axis=1), 16, 32)             # panned beyond the range
    check(emu, oracle, a, np.roll(np.roll(a, 32, axis=1), -32, axis=0), 8, 32)   # exactly on the dx = +R column / dy = -R row
    rng = np.random.default_rng(5)
    hi = rng.integers(0, 2, (128, 256)).astype(np.uint8) * 255
    check(emu, oracle, hi, 255 - hi, 16, 32)


def test_sea_wide_range_strips(emu, oracle):
    """+-32 strips with 32 halo rows of the previous frame (tiled frames): rows of the whole-frame result."""
    prev, cur, _ = synth.make_pair(264, 224, 32, index=11, noise_lsb=1)
    omv, ocost, _ = oracle.block_match(prev, cur, 8, 32, 0, threads=oracle.max_threads(), fast=True)
    for y0, rows in ((0, 64), (64, 96), (160, 64)):
        mv, cost, ent, res, _ = run_sea(emu, prev, cur, 8, 32, strip=(y0, rows))
        sl = slice(y0 // 8, (y0 + rows) // 8)
        r = res[0]
        np.testing.assert_array_equal(cost[0][r], ocost[sl][r])
        np.testing.assert_array_equal(mv[0][r], omv[sl][r])


def test_sea_ties_and_flat(emu, oracle):
    prev = np.full((96, 160), 77, np.uint8)
    frac, _ = check(emu, oracle, prev, prev.copy(), 16, 16, min_resolved=1.0)
    x = (np.arange(160) % 8 * 30).astype(np.uint8)
    prev = np.tile(x, (96, 1))
    check(emu, oracle, prev, prev.copy(), 16, 16)
    check(emu, oracle, prev, np.roll(prev, 3, axis=1), 8, 16)
    check(emu, oracle, prev, np.roll(prev, 3, axis=1), 16, 8)
    # vertical period too: zero-cost matches at many (dx, dy); the shortest must win
    prev = ((np.arange(96)[:, None] % 6) * 40 + (np.arange(160)[None] % 5) * 9).astype(np.uint8)
    check(emu, oracle, prev, np.roll(np.roll(prev, 2, axis=0), -4, axis=1), 16, 16)
    check(emu, oracle, prev, np.roll(np.roll(prev, 2, axis=0), -4, axis=1), 8, 8)


def test_sea_motion_beyond_range(emu, oracle):
    """content panned further than the search range must never be reported with |d| > range."""
    for block, search, pan in ((16, 8, 12), (8, 8, 11), (16, 16, 20), (16, 8, 16)):
        prev = synth.textured_plane(99, 320, 160)
        cur = np.roll(prev, pan, axis=1)
        mv, cost, ent, res, _ = run_sea(emu, prev, cur, block, search)
        assert not res.any() or np.abs(mv[res]).max() <= search
        check(emu, oracle, prev, cur, block, search)


def test_sea_unrelated_and_extremes(emu, oracle):
    a = synth.textured_plane(1, 256, 128)
    b = synth.textured_plane(2, 256, 128)
    check(emu, oracle, a, b, 16, 16)
    check(emu, oracle, a, b, 8, 8)
    rng = np.random.default_rng(5)
    hi = rng.integers(0, 2, (128, 256)).astype(np.uint8) * 255          # maximal window sums / bounds
    check(emu, oracle, hi, 255 - hi, 16, 16)
    check(emu, oracle, np.zeros((64, 128), np.uint8), np.full((64, 128), 255, np.uint8), 16, 16)


def test_sea_strips(emu, oracle):
    """Strips of a frame with halo rows of the previous frame (multi-GPU tiling): rows of the whole-frame result."""
    prev, cur, _ = synth.make_pair(384, 272, 16, index=7, noise_lsb=1)
    omv, ocost, _ = oracle.block_match(prev, cur, 16, 16, 0, threads=oracle.max_threads(), fast=True)
    for y0, rows in ((0, 96), (96, 96), (192, 80)):
        mv, cost, ent, res, _ = run_sea(emu, prev, cur, 16, 16, strip=(y0, rows), tile_h=32 if y0 else 64)
        sl = slice(y0 // 16, (y0 + rows) // 16)
        r = res[0]
        np.testing.assert_array_equal(cost[0][r], ocost[sl][r])
        np.testing.assert_array_equal(mv[0][r], omv[sl][r])


def test_sea_randomized_geometries(emu, oracle):
    """Seeded random geometries and contents through every instance of the kernel (block 8 / 16, range 8 / 16 / 32, both
    tile heights): widths that are not multiples of the tile, frames smaller than the search window, pans, local motion,
    noise, strips at random offsets."""
    rng = np.random.default_rng(20260)
    for case in range(10):
        block = int(rng.choice([8, 16]))
        search = int(rng.choice([8, 16, 32]))
        w = int(rng.integers(4, 30)) * 16
        h = int(rng.integers(3, 14)) * 16
        base = synth.textured_plane(int(rng.integers(1, 1000)), w + 64, h + 64)
        dx, dy = (int(v) for v in rng.integers(-search, search + 1, 2))
        prev = base[32:32 + h, 32:32 + w].copy()
        cur = base[32 + dy:32 + dy + h, 32 + dx:32 + dx + w].copy()
        if rng.random() < 0.5:                                            # a rectangle with its own motion
            y0, x0 = int(rng.integers(0, h // 2)), int(rng.integers(0, w // 2))
            cur[y0:y0 + h // 3, x0:x0 + w // 3] = np.roll(prev, (int(rng.integers(-5, 6)), int(rng.integers(-5, 6))), (0, 1))[y0:y0 + h // 3, x0:x0 + w // 3]
        noise = int(rng.integers(0, 3))
        if noise:
            cur = np.clip(cur.astype(np.int16) + rng.integers(-noise, noise + 1, cur.shape), 0, 255).astype(np.uint8)
        th = int(rng.choice([32, 64]))
        check(emu, oracle, prev, cur, block, search, tile_h=th)
        rows = (h // block // 2) * block
        if rows >= block and h - rows >= block:                           # the lower part as a strip with its halo
            omv, ocost, _ = oracle.block_match(prev, cur, block, search, 0, threads=oracle.max_threads(), fast=True)
            y0 = h - rows - ((h - rows) % block)
            y0 -= y0 % block
            mv, cost, ent, res, _ = run_sea(emu, prev, cur, block, search, strip=(y0, (h - y0) // block * block), tile_h=th)
            sl = slice(y0 // block, y0 // block + mv.shape[1])
            r = res[0]
            np.testing.assert_array_equal(cost[0][r], ocost[sl][r])
            np.testing.assert_array_equal(mv[0][r], omv[sl][r])
