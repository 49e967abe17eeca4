"""K7-K9 parity: the cv-decoder dense-flow front end on the B200 through the C ABI, bit-exact against
(a) the vectors OpenCV itself produced with the reference's call parameters (tests/golden/golden_cv_v1.npz,
    cv-decoder/src/lib.rs:138, 204-236) and
(b) the oracle (oracle/cv_front.c) on seeded inputs up to 1080p / 4K, plus size-independent properties."""
import os

import numpy as np
import pytest

from ofps_b200 import capi, synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cv_v1.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


def scene_gray(w, h, seed, n_rects=None):
    """Flat regions + rectangles + isolated pixels: the contrast mask covers a part of the frame only."""
    rng = np.random.default_rng(seed)
    n_rects = max(40, w * h // 6000) if n_rects is None else n_rects
    yy, xx = np.mgrid[0:h, 0:w]
    g = (110 + 40 * np.sin(xx / 37.0) * np.cos(yy / 29.0)).astype(np.uint8)
    for _ in range(n_rects):
        rw, rh = int(rng.integers(2, max(3, min(w // 6, 90)))), int(rng.integers(2, max(3, min(h // 6, 90))))
        x0, y0 = int(rng.integers(-rw // 2, w)), int(rng.integers(-rh // 2, h))
        g[max(y0, 0):y0 + rh, max(x0, 0):x0 + rw] = rng.integers(0, 256)
    for _ in range(n_rects):
        g[int(rng.integers(0, h)), int(rng.integers(0, w))] = rng.integers(0, 256)
    return g


def random_flow(w, h, seed, scale=6.0):
    rng = np.random.default_rng(seed)
    return ((rng.random((h, w, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(scale)).astype(np.float32)


# ------------------------------------------------------------------------------------------ K7
def test_frame_convert_matches_opencv(ctx, gold):
    for i in range(len(gold["cases"])):
        gray, rgba = ctx.frame_convert(gold[f"c{i}_bgr"], want_rgba=True)
        assert np.array_equal(gray, gold[f"c{i}_gray"]), f"case {i}"
        bgr = gold[f"c{i}_bgr"]
        assert np.array_equal(rgba[..., :3], bgr[..., ::-1]) and (rgba[..., 3] == 255).all()
    gray, _ = ctx.frame_convert(gold["bgra"])
    assert np.array_equal(gray, gold["bgra_gray"])
    gray, _ = ctx.frame_convert(np.ascontiguousarray(gold["bgra"][..., :3]), rgb_order=True)
    assert np.array_equal(gray, gold["rgb_gray"])


@pytest.mark.parametrize("w,h,ch", [(1920, 1080, 3), (1920, 1080, 4), (1921, 1079, 3), (3840, 2160, 3), (5, 3, 4)])
def test_frame_convert_matches_oracle(ctx, oracle, w, h, ch):
    img = np.random.default_rng(w + h + ch).integers(0, 256, (h, w, ch), dtype=np.uint8)
    gray, rgba = ctx.frame_convert(img, want_rgba=True)
    assert np.array_equal(gray, oracle.bgr_to_gray(img))
    assert np.array_equal(rgba, oracle.bgr_to_rgba(img))
    only_rgba = ctx.frame_convert(img, want_gray=False, want_rgba=True)
    assert only_rgba[0] is None and np.array_equal(only_rgba[1], rgba)


def test_frame_convert_all_colours(ctx, oracle):
    """Every (B, G, R) with B, G on a stride-3 lattice and all 256 R: the fixed-point rounding everywhere."""
    b, g, r = np.meshgrid(np.arange(0, 256, 3), np.arange(0, 256, 3), np.arange(256), indexing="ij")
    img = np.stack([b, g, r], -1).astype(np.uint8).reshape(86 * 86, 256, 3)
    gray, _ = ctx.frame_convert(img)
    assert np.array_equal(gray, oracle.bgr_to_gray(img))


# ------------------------------------------------------------------------------------------ K8
def test_contrast_mask_matches_opencv(ctx, gold):
    for i in range(len(gold["cases"])):
        assert np.array_equal(ctx.contrast_mask(gold[f"c{i}_gray"]), gold[f"c{i}_mask"]), f"case {i} {gold['cases'][i]}"
    assert np.array_equal(ctx.contrast_mask(gold["tex_gray"]), gold["tex_mask"])
    assert np.array_equal(ctx.contrast_mask(gold["fb_gray"]), gold["fb_mask"])


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160), (1919, 1081), (640, 360), (300, 33), (256, 32), (257, 33),
                                 (512, 64), (31, 500), (1, 1), (2, 1), (1, 40), (6, 6)])
def test_contrast_mask_matches_oracle(ctx, oracle, w, h):
    g = scene_gray(w, h, 7 * w + h)
    want = oracle.contrast_mask(g)
    got = ctx.contrast_mask(g)
    assert np.array_equal(got, want)
    assert set(np.unique(got)) <= {0, 255}
    if w >= 256 and h >= 64:
        assert 0.02 < (want > 0).mean() < 0.98   # the case exercises both outcomes


def test_contrast_mask_textured_and_flat(ctx, oracle):
    tex = synth.textured_plane(0x0F950001, 1920, 1080)
    assert np.array_equal(ctx.contrast_mask(tex), oracle.contrast_mask(tex))
    flat = np.full((200, 300), 77, np.uint8)
    assert not ctx.contrast_mask(flat).any()
    # the mask is invariant to adding a constant (pure derivative) as long as nothing saturates
    g = scene_gray(640, 360, 3) // 2
    assert np.array_equal(ctx.contrast_mask(g), ctx.contrast_mask(g + 100))


def test_contrast_mask_dev_strides(ctx, oracle):
    """Device entry point with unaligned base pointers and pitches: the byte paths of the kernel."""
    w, h = 700, 90
    g = scene_gray(w, h, 11)
    want = oracle.contrast_mask(g)
    for gpitch, mpitch, goff, moff in ((w, w, 0, 0), (w + 3, w + 5, 1, 3), (704, 704, 0, 0), (704, 720, 16, 16)):
        gbuf = np.zeros(gpitch * h + 64, np.uint8)
        gbuf[goff:goff + gpitch * h].reshape(h, gpitch)[:, :w] = g
        d_g = ctx.dev_alloc(gbuf.size)
        d_m = ctx.dev_alloc(mpitch * h + 64)
        try:
            ctx.to_device(d_g, gbuf)
            mbuf = np.full(mpitch * h + 64, 9, np.uint8)
            ctx.to_device(d_m, mbuf)
            ctx.contrast_mask_dev(d_g + goff, w, h, gpitch, d_m + moff, mpitch)
            ctx.to_host(mbuf, d_m)
            ctx.sync()
            m = mbuf[moff:moff + mpitch * h].reshape(h, mpitch)
            assert np.array_equal(m[:, :w], want), (gpitch, mpitch, goff, moff)
            assert (m[:, w:] == 9).all() and (mbuf[:moff] == 9).all()   # nothing outside the w columns is written
        finally:
            ctx.dev_free(d_g)
            ctx.dev_free(d_m)


# ------------------------------------------------------------------------------------------ K9
@pytest.mark.parametrize("grid", [(0, 0), (20, 11), (150, 84), (1, 1), (7, 1), (1, 5), (160, 90), (200, 120), (3, 2)])
@pytest.mark.parametrize("use_mask", [True, False])
def test_flow_entries_farneback_fixture(ctx, oracle, gold, grid, use_mask):
    flow = gold["fb_flow"]
    mask = gold["fb_mask"] if use_mask else None
    got = ctx.flow_entries(flow, mask, *grid)
    want = oracle.flow_entries(flow, mask, *grid)
    assert got.shape == want.shape and len(got) > 0
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("w,h,gw,gh", [(1920, 1080, 150, 84), (1920, 1080, 0, 0), (3840, 2160, 150, 84), (1280, 720, 2000, 720),
                                       (1921, 1079, 149, 83), (640, 360, 150, 84), (9000, 6, 2, 2), (9000, 6, 40, 3),
                                       (50, 40, 150, 84), (1, 1, 1, 1), (17, 1, 4, 1)])
def test_flow_entries_matches_oracle(ctx, oracle, w, h, gw, gh):
    flow = random_flow(w, h, w * 3 + h)
    mask = ((np.random.default_rng(w + 5 * h).random((h, w)) < 0.35).astype(np.uint8)) * 255
    for m in (mask, None):
        got = ctx.flow_entries(flow, m, gw, gh)
        want = oracle.flow_entries(flow, m, gw, gh)
        assert got.shape == want.shape
        assert got.tobytes() == want.tobytes()


def test_flow_entries_empty_and_capacity(ctx):
    flow = random_flow(64, 48, 1)
    zero = np.zeros((48, 64), np.uint8)
    assert len(ctx.flow_entries(flow, zero, 0, 0)) == 0
    assert len(ctx.flow_entries(flow, zero, 10, 8)) == 0
    with pytest.raises(capi.OfpsError) as e:
        ctx.flow_entries(flow, None, 0, 0, cap=100)
    assert e.value.code == capi.E_CAPACITY
    with pytest.raises(capi.OfpsError) as e:
        ctx.flow_entries(flow, None, 10, 0)
    assert e.value.code == capi.E_INVALID


def test_flow_entries_properties_1080p(ctx):
    """Size-independent properties at full size: a constant flow gives every touched cell the same mean
    (up to the one-hit 1/(1+eps) quirk), the per-pixel path is the identity map, output order is (x, y)."""
    w, h, gw, gh = 1920, 1080, 150, 84
    flow = np.empty((h, w, 2), np.float32)
    flow[..., 0], flow[..., 1] = 3.0, -1.5
    e = ctx.flow_entries(flow, None, gw, gh)
    assert len(e) == gw * gh
    nx, ny = np.float32(1) / np.float32(w), np.float32(1) / np.float32(h)
    # >= 2 hits per cell: counts are exact integers, sum of n equal values / n differs from the value by rounding only
    assert np.allclose(e[:, 2], np.float32(3.0) * nx, rtol=1e-4) and np.allclose(e[:, 3], np.float32(-1.5) * ny, rtol=1e-4)
    xs = np.rint(e[:, 0] * gw - 0.5).astype(int)
    ys = np.rint(e[:, 1] * gh - 0.5).astype(int)
    assert np.array_equal(xs * gh + ys, np.arange(gw * gh))
    p = ctx.flow_entries(flow, None, 0, 0)
    assert len(p) == w * h
    assert np.array_equal(p[:, 2], np.full(w * h, np.float32(3.0) * nx)) and p[0, 0] == np.float32(0.5) * nx
    assert np.array_equal(p[:, 1].reshape(h, w)[:, 0], ((np.arange(h, dtype=np.float32) + np.float32(0.5)) * ny))


@pytest.mark.parametrize("w,h", [(1920, 1080), (640, 360), (300, 70)])
@pytest.mark.parametrize("use_mask", [True, False])
def test_cv_flow_frame_fused(ctx, oracle, w, h, use_mask):
    """gray + flow -> entries in one call == oracle mask -> oracle flow_entries; then into the estimator's input."""
    g = scene_gray(w, h, w + h)
    flow = random_flow(w, h, 2 * w + h)
    gw, gh = capi.mfield_size(w, h)
    for grid in ((gw, gh), (0, 0)):
        got = ctx.cv_flow_frame(g, flow, use_mask, *grid)
        want = oracle.flow_entries(flow, oracle.contrast_mask(g) if use_mask else None, *grid)
        assert got.shape == want.shape and got.tobytes() == want.tobytes()


def test_flow_entries_dev_feeds_estimator(ctx, oracle):
    """Device-resident chain: flow -> entries (HBM) -> Almeida LSQ, no host round trip of the entries."""
    w, h = 640, 360
    ent, _ = synth.rotation_field(w, h, 16 / 9, 22.275, (0.4, -0.3, 0.2))
    flow = np.ascontiguousarray((ent[:, 2:].reshape(h, w, 2) * np.array([w, h], np.float32)).astype(np.float32))
    gw, gh = capi.mfield_size(w, h)
    d_flow = ctx.dev_alloc(flow.nbytes)
    d_ent = ctx.dev_alloc(gw * gh * 16)
    try:
        ctx.to_device(d_flow, flow)
        n = ctx.flow_entries_dev(d_flow, 2 * w, None, 0, w, h, gw, gh, d_ent, gw * gh)
        assert n == gw * gh
        back = np.empty((n, 4), np.float32)
        ctx.to_host(back, d_ent)
        ctx.sync()
        want = oracle.flow_entries(flow, None, gw, gh)
        assert back.tobytes() == want.tobytes()
        q = ctx.almeida(None, 16 / 9, 22.275, d_entries=d_ent, n=n)
        q_ref = oracle.almeida_lsq_f64(want, 16 / 9, 22.275)
        assert min(np.abs(q - q_ref).max(), np.abs(q + q_ref).max()) < 1e-4
    finally:
        ctx.dev_free(d_flow)
        ctx.dev_free(d_ent)


# ------------------------------------------------------------------------------------------ K7b
def test_frame_resize_matches_opencv(ctx, gold):
    for i, (sw, sh, dw, dh) in enumerate(gold["resize_cases"]):
        got = ctx.frame_resize(gold[f"rz{i}_src"], int(dw), int(dh))
        assert np.array_equal(got, gold[f"rz{i}_dst"]), (sw, sh, dw, dh)


@pytest.mark.parametrize("sw,sh,ch", [(1920, 1080, 3), (1920, 1080, 4), (3840, 2160, 3), (640, 360, 3), (151, 85, 3)])
def test_frame_resize_matches_oracle(ctx, oracle, sw, sh, ch):
    img = np.random.default_rng(sw + ch).integers(0, 256, (sh, sw, ch), dtype=np.uint8)
    dw, dh = capi.mfield_size(sw, sh)
    assert np.array_equal(ctx.frame_resize(img, dw, dh), oracle.resize_linear(img, dw, dh))
    with pytest.raises(capi.OfpsError):
        ctx.frame_resize(img, sw + 1, sh)


def test_process_fullres_off_chain(ctx, oracle, gold):
    """'Process Fullres' off (cv-decoder/src/lib.rs:124-138, 272): resize -> gray -> mask -> per-pixel entries."""
    bgr = np.random.default_rng(4).integers(0, 256, (360, 640, 3), dtype=np.uint8)
    bgr[100:200, 200:400] = 30
    dw, dh = capi.mfield_size(640, 360)
    small = ctx.frame_resize(bgr, dw, dh)
    gray, _ = ctx.frame_convert(small)
    flow = random_flow(dw, dh, 12)
    got = ctx.cv_flow_frame(gray, flow, True, 0, 0)
    o_small = oracle.resize_linear(bgr, dw, dh)
    o_gray = oracle.bgr_to_gray(o_small)
    want = oracle.flow_entries(flow, oracle.contrast_mask(o_gray), 0, 0)
    assert np.array_equal(small, o_small) and np.array_equal(gray, o_gray)
    assert got.shape == want.shape and got.tobytes() == want.tobytes()
