"""TEST INFRASTRUCTURE: a stand-in for ofps_b200.capi.Context backed by the CPU emulation of the cv-front kernels
(tests/emu).  Used only to dry-run tests/test_gpu_cv_front.py on the GPU-less container
(OFPSB_EMU_CTX=1 python -m pytest tests/test_gpu_cv_front.py -m gpu) before GPU time is spent on it."""
import ctypes as C
import os

import numpy as np
import pytest

from ofps_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p, _f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float)


class EmuContext:
    def __init__(self):
        L = C.CDLL(os.path.join(HERE, "_build", "libemu_cv_front.so"))
        L.emu_frame_convert.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _u8p]
        L.emu_frame_resize.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
        L.emu_contrast_mask.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.emu_flow_entries.argtypes = [_f32p, C.c_size_t, _u8p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_ulonglong)]
        self.L = L

    def frame_convert(self, img, rgb_order=False, want_gray=True, want_rgba=False):
        img = np.ascontiguousarray(img, np.uint8)
        h, w, ch = img.shape
        gs = (w + 3) & ~3
        gray = np.empty((h, gs), np.uint8) if want_gray else None
        rgba = np.empty((h, w, 4), np.uint8) if want_rgba else None
        rc = self.L.emu_frame_convert(img.ctypes.data_as(_u8p), w, h, w * ch, ch, int(rgb_order),
                                      gray.ctypes.data_as(_u8p) if want_gray else None, gs,
                                      rgba.ctypes.data_as(_u8p) if want_rgba else None)
        assert rc == 0
        return (np.ascontiguousarray(gray[:, :w]) if want_gray else None), rgba

    def frame_resize(self, img, dw, dh):
        img = np.ascontiguousarray(img, np.uint8)
        h, w, ch = img.shape
        out = np.empty((dh, dw, ch), np.uint8)
        if self.L.emu_frame_resize(img.ctypes.data_as(_u8p), w, h, w * ch, ch, out.ctypes.data_as(_u8p), dw, dh) != 0:
            raise capi.OfpsError(capi.E_INVALID, "emu")
        return out

    def contrast_mask(self, gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        gp = (w + 15) & ~15
        g = np.zeros((h, gp), np.uint8)
        g[:, :w] = gray
        m = np.zeros((h, gp), np.uint8)
        assert self.L.emu_contrast_mask(g.ctypes.data_as(_u8p), w, h, gp, m.ctypes.data_as(_u8p), gp) == 0
        return np.ascontiguousarray(m[:, :w])

    def flow_entries(self, flow, mask=None, gw=0, gh=0, cap=None):
        flow = np.ascontiguousarray(flow, np.float32)
        h, w, _ = flow.shape
        if (gw == 0) != (gh == 0):
            raise capi.OfpsError(capi.E_INVALID, "emu")
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        cap = (gw * gh if gw else w * h) if cap is None else cap
        out = np.empty((max(cap, 1), 4), np.float32)
        n = C.c_ulonglong()
        rc = self.L.emu_flow_entries(flow.ctypes.data_as(_f32p), 2 * w, mask.ctypes.data_as(_u8p) if mask is not None else None,
                                     w, w, h, gw, gh, out.ctypes.data, cap, C.byref(n))
        assert rc == 0
        if n.value > cap:
            raise capi.OfpsError(capi.E_CAPACITY, "emu")
        return out[:n.value].copy()

    def cv_flow_frame(self, gray, flow, use_mask=True, gw=0, gh=0):
        return self.flow_entries(flow, self.contrast_mask(gray) if use_mask else None, gw, gh)

    def __getattr__(self, name):
        pytest.skip(f"emulated context has no {name}")
