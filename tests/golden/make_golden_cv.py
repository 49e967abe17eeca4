"""Generates tests/golden/golden_cv_v1.npz by running OpenCV ITSELF (cv2, present in the build container)
with the exact call parameters of the reference's cv-decoder (cv-decoder/src/lib.rs:138, 204-236):

    gray   = cvtColor(frame, COLOR_BGR2GRAY)
    sobel  = Sobel(gray, CV_32F, 1, 1, ksize=5, scale=1, delta=0, BORDER_DEFAULT)
    thresh = threshold(sobel, 20, 255, THRESH_BINARY)
    mask   = dilate(thresh, getStructuringElement(MORPH_ELLIPSE, (11,11), (5,5)), (-1,-1), 1, BORDER_DEFAULT)
    flow   = calcOpticalFlowFarneback(old_gray, gray, None, 0.5, 5, 13, 3, 7, 1.5, 0)      (:186-197; input only)

These are the library calls the reference makes, so the vectors pin the oracle (oracle/cv_front.c) and the
CUDA path to the reference's real arithmetic for this stage.  cv2 does not exist on the GPU box: the
vectors are committed.  Re-run with   python tests/golden/make_golden_cv.py   (prints the cv2 version)."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ofps_b200 import synth  # noqa: E402


def scene(w, h, seed, n_rects=6, smooth=True):
    """BGR test frame: smooth gradient background + flat rectangles (corners trigger the mixed
    derivative, large flat areas stay below the threshold) + a few isolated bright pixels."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.empty((h, w, 3), np.float64)
    for c in range(3):
        img[..., c] = 90 + 50 * np.sin(xx / (17.0 + 5 * c)) * np.cos(yy / (23.0 - 3 * c))
    for _ in range(n_rects):
        rw, rh = int(rng.integers(2, max(3, w // 3))), int(rng.integers(2, max(3, h // 3)))
        x0, y0 = int(rng.integers(-rw // 2, w)), int(rng.integers(-rh // 2, h))
        img[max(y0, 0):y0 + rh, max(x0, 0):x0 + rw] = rng.integers(0, 256, 3)
    for _ in range(4):
        img[int(rng.integers(0, h)), int(rng.integers(0, w))] = rng.integers(0, 256, 3)
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if smooth and min(w, h) >= 3:
        img = cv2.GaussianBlur(img, (3, 3), 0)
    return img


def cv_mask(gray):
    sob = cv2.Sobel(gray, cv2.CV_32F, 1, 1, ksize=5, scale=1.0, delta=0.0, borderType=cv2.BORDER_DEFAULT)
    _, th = cv2.threshold(sob, 20.0, 255.0, cv2.THRESH_BINARY)
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (11, 11), (5, 5))
    mask = cv2.dilate(th, se, anchor=(-1, -1), iterations=1, borderType=cv2.BORDER_DEFAULT)
    return sob, th, mask


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    out["ellipse11"] = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (11, 11), (5, 5)).astype(np.uint8)
    # (w, h, seed, smooth): big enough to span several CUDA tiles and word boundaries, odd sizes,
    # sizes below the 11x11 element and the 5x5 kernel, single rows / columns
    cases = [(300, 70, 1, True), (257, 33, 2, False), (96, 64, 3, True), (33, 47, 4, False), (31, 9, 5, False),
             (8, 8, 6, False), (5, 12, 7, False), (4, 3, 8, False), (2, 2, 9, False), (7, 1, 10, False),
             (1, 11, 11, False), (64, 6, 12, False), (540, 40, 13, True)]
    out["cases"] = np.array([(w, h) for w, h, _, _ in cases], np.int32)
    fracs = []
    for i, (w, h, seed, smooth) in enumerate(cases):
        bgr = scene(w, h, seed, smooth=smooth)
        gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
        sob, th, mask = cv_mask(gray)
        assert np.array_equal(sob, np.rint(sob)) and np.abs(sob).max() < 2 ** 23
        out[f"c{i}_bgr"] = bgr
        out[f"c{i}_gray"] = gray
        out[f"c{i}_sobel"] = sob.astype(np.int32)
        out[f"c{i}_thresh"] = (th > 0).astype(np.uint8) * 255
        out[f"c{i}_mask"] = (mask > 0).astype(np.uint8) * 255
        fracs.append(round(float((mask > 0).mean()), 3))
    # textured noise (the block matcher's synthetic frames): dense mask regime
    tex = synth.textured_plane(0x0F950001, 128, 48)
    sob, th, mask = cv_mask(tex)
    out["tex_gray"], out["tex_sobel"], out["tex_mask"] = tex, sob.astype(np.int32), (mask > 0).astype(np.uint8) * 255
    # BGRA input and RGB order
    bgra = np.concatenate([scene(45, 21, 20), np.full((21, 45, 1), 7, np.uint8)], axis=2)
    out["bgra"] = bgra
    out["bgra_gray"] = cv2.cvtColor(bgra, cv2.COLOR_BGRA2GRAY)
    out["rgb_gray"] = cv2.cvtColor(np.ascontiguousarray(bgra[..., :3]), cv2.COLOR_RGB2GRAY)
    # resize(INTER_LINEAR) as cv-decoder calls it when "Process Fullres" is off (cv-decoder/src/lib.rs:127-135)
    rz = [(300, 170, 150, 85), (257, 101, 150, 58), (160, 90, 150, 84), (64, 48, 32, 24), (50, 40, 50, 40), (97, 31, 13, 7),
          (40, 30, 40, 15), (33, 20, 7, 20)]
    out["resize_cases"] = np.array(rz, np.int32)
    rng = np.random.default_rng(77)
    for i, (sw, sh, dw, dh) in enumerate(rz):
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        out[f"rz{i}_src"] = src
        out[f"rz{i}_dst"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    # a real Farneback flow (third-party; used as INPUT of the flow -> MotionEntry stage only)
    a = scene(160, 90, 30, n_rects=10)
    b = np.roll(a, (2, -3), axis=(0, 1))
    ga, gb = cv2.cvtColor(a, cv2.COLOR_BGR2GRAY), cv2.cvtColor(b, cv2.COLOR_BGR2GRAY)
    flow = cv2.calcOpticalFlowFarneback(ga, gb, None, 0.5, 5, 13, 3, 7, 1.5, 0)
    out["fb_gray"] = gb
    out["fb_flow"] = flow.astype(np.float32)
    out["fb_mask"] = (cv_mask(gb)[2] > 0).astype(np.uint8) * 255
    np.savez_compressed(os.path.join(HERE, "golden_cv_v1.npz"), **out)
    print("cv2", cv2.__version__, "mask fractions", fracs, "fb mask", round(float((out["fb_mask"] > 0).mean()), 3))
    print("wrote golden_cv_v1.npz", os.path.getsize(os.path.join(HERE, "golden_cv_v1.npz")), "bytes")


if __name__ == "__main__":
    main()
