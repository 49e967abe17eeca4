"""CPU, only where OpenCV is importable (the build container; skipped on the GPU box): the oracle's cv-decoder front end
against cv2 itself on fresh random inputs — a wider net than the committed vectors of tests/golden/golden_cv_v1.npz.
Calls and parameters are the reference's (cv-decoder/src/lib.rs:127-138, 204-236)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _scene(rng, w, h):
    kind = rng.integers(0, 3)
    if kind == 0:   # flat + rectangles: sparse mask
        g = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        for _ in range(int(rng.integers(1, 12))):
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            g[y:y + int(rng.integers(1, 12)), x:x + int(rng.integers(1, 12))] = rng.integers(0, 256)
        return g
    if kind == 1:   # noise: dense mask
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]   # smooth ramps: values near the threshold
    return ((xx * int(rng.integers(1, 5)) + yy * int(rng.integers(1, 5)) + (xx * yy) // int(rng.integers(3, 40))) % 256).astype(np.uint8)


def _cv_mask(gray):
    sob = cv2.Sobel(gray, cv2.CV_32F, 1, 1, ksize=5, scale=1.0, delta=0.0, borderType=cv2.BORDER_DEFAULT)
    _, th = cv2.threshold(sob, 20.0, 255.0, cv2.THRESH_BINARY)
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (11, 11), (5, 5))
    return sob, cv2.dilate(th, se, anchor=(-1, -1), iterations=1, borderType=cv2.BORDER_DEFAULT)


def test_contrast_mask_random_images(oracle):
    rng = np.random.default_rng(2025)
    for _ in range(120):
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 70))
        g = _scene(rng, w, h)
        sob, mask = _cv_mask(g)
        omask, osob = oracle.contrast_mask(g, return_sobel=True)
        assert np.array_equal(osob, sob.astype(np.int32)), (w, h)
        assert np.array_equal(omask > 0, mask > 0), (w, h)


def test_bgr_to_gray_and_resize_random(oracle):
    rng = np.random.default_rng(7)
    for _ in range(40):
        w, h = int(rng.integers(2, 200)), int(rng.integers(2, 150))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(oracle.bgr_to_gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
        dw, dh = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
        assert np.array_equal(oracle.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)), (w, h, dw, dh)


def test_flow_entries_random_shapes(oracle):
    import pyref
    rng = np.random.default_rng(11)
    for _ in range(25):
        w, h = int(rng.integers(1, 40)), int(rng.integers(1, 30))
        flow = ((rng.random((h, w, 2), dtype=np.float32) - np.float32(0.5)) * np.float32(9)).astype(np.float32)
        mask = (rng.random((h, w)) < rng.random()).astype(np.uint8) * 255 if rng.random() < 0.7 else None
        gw, gh = (0, 0) if rng.random() < 0.3 else (int(rng.integers(1, 50)), int(rng.integers(1, 40)))
        got, want = oracle.flow_entries(flow, mask, gw, gh), pyref.flow_entries(flow, mask, gw, gh)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (w, h, gw, gh)
