"""CPU: the cv-front kernels (ofps_b200/csrc/cv_front.cu) executed UNMODIFIED on the thread-per-CUDA-thread stand-in of
tests/emu/cuda_emu.h and compared with the oracle / the OpenCV golden vectors.  This checks kernel LOGIC (indices,
barriers, border reflection, summation order) on the GPU-less build container; the real parity tests are the
`-m gpu` ones in tests/test_gpu_cv_front.py.  Test infrastructure only: nothing here is a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
LIB = os.path.join(EMU, "_build", "libemu_cv_front.so")
GOLDEN = os.path.join(HERE, "golden", "golden_cv_v1.npz")


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU, "emu_cv_front.cpp"), os.path.join(EMU, "cuda_emu.h"),
            os.path.join(HERE, "..", "ofps_b200", "csrc", "cv_front.cu"), os.path.join(HERE, "..", "ofps_b200", "csrc", "common.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-I" + EMU,
                               "-o", LIB, srcs[0]])
    L = C.CDLL(LIB)
    u8p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    L.emu_frame_convert.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, u8p]
    L.emu_frame_resize.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
    L.emu_contrast_mask.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
    L.emu_flow_entries.argtypes = [f32p, C.c_size_t, u8p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p,
                                   C.c_size_t, C.POINTER(C.c_ulonglong)]
    return L


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def emu_mask(L, gray, pad=0):
    h, w = gray.shape
    g = np.zeros((h, w + pad), np.uint8)
    g[:, :w] = gray
    m = np.full((h, w + pad), 7, np.uint8)
    assert L.emu_contrast_mask(_u8(g), w, h, w + pad, _u8(m), w + pad) == 0
    assert (m[:, w:] == 7).all()
    return m[:, :w]


def emu_flow_entries(L, flow, mask, gw, gh, cap=None):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    cap = max(w * h, gw * gh, 1) if cap is None else cap
    out = np.zeros((cap, 4), np.float32)
    n = C.c_ulonglong(0)
    mp = _u8(mask) if mask is not None else None
    assert L.emu_flow_entries(flow.ctypes.data_as(C.POINTER(C.c_float)), 2 * w, mp, w, w, h, gw, gh, out.ctypes.data, cap,
                              C.byref(n)) == 0
    return out[:min(n.value, cap)], n.value


def test_emu_frame_convert(emu, gold, oracle):
    for img, rgb in ((gold["c0_bgr"], 0), (gold["c3_bgr"], 0), (gold["bgra"], 0), (gold["bgra"], 1), (gold["c9_bgr"], 0)):
        img = np.ascontiguousarray(img)
        h, w, ch = img.shape
        for gs in (w, (w + 3) & ~3):
            gray = np.zeros((h, gs), np.uint8)
            rgba = np.zeros((h, w, 4), np.uint8)
            assert emu.emu_frame_convert(_u8(img), w, h, w * ch, ch, rgb, _u8(gray), gs, _u8(rgba)) == 0
            assert np.array_equal(gray[:, :w], oracle.bgr_to_gray(img, bool(rgb)))
            assert np.array_equal(rgba, oracle.bgr_to_rgba(img))


@pytest.mark.parametrize("w,h,ch,pitch,gpitch", [(64, 5, 3, 192, 64), (80, 3, 4, 320, 80), (53, 4, 3, 160, 64), (20, 3, 4, 80, 32),
                                                  (16, 2, 3, 48, 16), (100, 2, 3, 304, 112)])
def test_emu_frame_convert_vector_path(emu, oracle, w, h, ch, pitch, gpitch):
    """16-byte aligned pitches: the 16-pixel vector path, its ragged row end and the vector / scalar RGBA stores."""
    rng = np.random.default_rng(w * ch)
    buf = rng.integers(0, 256, (h, pitch), dtype=np.uint8)
    img = np.ascontiguousarray(buf[:, :w * ch].reshape(h, w, ch))
    assert buf.ctypes.data % 16 == 0
    for rgb in (0, 1):
        gray = np.full((h, gpitch), 3, np.uint8)
        rgba = np.zeros((h, w, 4), np.uint8)
        assert emu.emu_frame_convert(_u8(buf), w, h, pitch, ch, rgb, _u8(gray), gpitch, _u8(rgba)) == 0
        assert np.array_equal(gray[:, :w], oracle.bgr_to_gray(img, bool(rgb))) and (gray[:, w:] == 3).all()
        assert np.array_equal(rgba, oracle.bgr_to_rgba(img))


def test_emu_frame_resize(emu, gold, oracle):
    for i, (sw, sh, dw, dh) in enumerate(gold["resize_cases"]):
        src = np.ascontiguousarray(gold[f"rz{i}_src"])
        out = np.zeros((dh, dw, 3), np.uint8)
        assert emu.emu_frame_resize(_u8(src), sw, sh, sw * 3, 3, _u8(out), dw, dh) == 0
        assert np.array_equal(out, gold[f"rz{i}_dst"]), (sw, sh, dw, dh)
    src = np.random.default_rng(3).integers(0, 256, (90, 160, 4), dtype=np.uint8)
    out = np.zeros((37, 80, 4), np.uint8)
    assert emu.emu_frame_resize(_u8(src), 160, 90, 640, 4, _u8(out), 80, 37) == 0
    assert np.array_equal(out, oracle.resize_linear(src, 80, 37))
    assert emu.emu_frame_resize(_u8(src), 160, 90, 640, 4, _u8(out), 161, 37) != 0   # enlarging is refused


def test_emu_contrast_mask_golden(emu, gold):
    for i in range(len(gold["cases"])):
        assert np.array_equal(emu_mask(emu, gold[f"c{i}_gray"]), gold[f"c{i}_mask"]), f"case {i} {gold['cases'][i]}"
    assert np.array_equal(emu_mask(emu, gold["tex_gray"]), gold["tex_mask"])
    # 16-byte aligned pitch takes the vector store path and the word-load path of interior tiles
    assert np.array_equal(emu_mask(emu, gold["c12_gray"], pad=4), gold["c12_mask"])
    assert np.array_equal(emu_mask(emu, gold["c0_gray"], pad=4), gold["c0_mask"])


def test_emu_contrast_mask_interior_tiles(emu, oracle):
    # 3 x 3 tiles: the centre tile is fully interior (word loads, no border fix-up)
    rng = np.random.default_rng(5)
    g = np.full((96, 800), 120, np.uint8)
    for _ in range(60):
        y, x = int(rng.integers(0, 96)), int(rng.integers(0, 800))
        g[y:y + int(rng.integers(1, 6)), x:x + int(rng.integers(1, 9))] = rng.integers(0, 256)
    want = oracle.contrast_mask(g)
    assert 0.05 < (want > 0).mean() < 0.95
    assert np.array_equal(emu_mask(emu, g), want)


@pytest.mark.parametrize("grid", [(0, 0), (20, 11), (150, 84), (1, 1), (7, 1), (1, 5), (200, 120), (3, 2)])
@pytest.mark.parametrize("use_mask", [True, False])
def test_emu_flow_entries(emu, gold, oracle, grid, use_mask):
    flow = gold["fb_flow"]
    mask = np.ascontiguousarray(gold["fb_mask"]) if use_mask else None
    got, n = emu_flow_entries(emu, flow, mask, *grid)
    want = oracle.flow_entries(flow, mask, *grid)
    assert n == len(want) and n > 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_emu_flow_entries_wide_cells_and_capacity(emu, oracle):
    # span of one CTA's cells wider than the staging buffer (4096 px) -> chunked rows
    rng = np.random.default_rng(9)
    flow = (rng.random((6, 9000, 2), dtype=np.float32) - 0.5) * 8
    mask = (rng.random((6, 9000)) < 0.4).astype(np.uint8) * 255
    for grid in ((2, 2), (40, 3)):
        got, n = emu_flow_entries(emu, flow, mask, *grid)
        want = oracle.flow_entries(flow, mask, *grid)
        assert n == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    got, n = emu_flow_entries(emu, flow[:, :300], np.ascontiguousarray(mask[:, :300]), 0, 0, cap=100)
    want = oracle.flow_entries(flow[:, :300], mask[:, :300], 0, 0)
    assert n == len(want) and np.array_equal(got.view(np.uint32), want[:100].view(np.uint32))
