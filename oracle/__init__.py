"""CPU ORACLE — test infrastructure, not product code.

ctypes/numpy wrapper around ``oracle/libofps_oracle.so`` (the plain-C restatement
of the reference hot path, see ``ofps_oracle.h`` for the file:line citations and
the pinning statement).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this package;
nothing under ``ofps_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libofps_oracle.so")

MV_DTYPE = np.dtype([("px", "<f4"), ("py", "<f4"), ("mx", "<f4"), ("my", "<f4")])


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, see oracle/Makefile)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("ofps_oracle.c", "cv_front.c", "ofps_oracle.h", "camera_almeida.inc", "Makefile")
    )
    if force or src_newer:
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L: C.CDLL) -> None:
    f32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    f64p = C.POINTER(C.c_double)
    szp = C.POINTER(C.c_size_t)
    L.orc_densify.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, f32p, f32p]
    L.orc_densify.restype = None
    L.orc_flow_field.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, f32p]
    L.orc_flow_field.restype = None
    L.orc_interpolate_empty_cells.argtypes = [f32p, f32p, C.c_size_t, C.c_size_t]
    L.orc_interpolate_empty_cells.restype = None
    L.orc_block_dim.argtypes = [C.c_float, C.c_size_t]
    L.orc_block_dim.restype = C.c_size_t
    L.orc_detect_block_motion.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_size_t, C.c_float,
                                          szp, szp, f32p, C.c_size_t, f32p]
    L.orc_detect_block_motion.restype = C.c_int
    L.orc_almeida_lsq_f.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, f32p]
    L.orc_almeida_lsq_f.restype = None
    L.orc_almeida_lsq_d.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, f64p]
    L.orc_almeida_lsq_d.restype = None
    L.orc_almeida_ransac_f.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_size_t, C.c_float,
                                       C.c_size_t, C.c_uint64, f32p, szp, szp]
    L.orc_almeida_ransac_f.restype = None
    L.orc_perm_index.argtypes = [C.c_uint64] * 5
    L.orc_perm_index.restype = C.c_uint64
    L.orc_quat_from_euler_d.argtypes = [C.c_double] * 3 + [f64p]
    L.orc_quat_from_euler_d.restype = None
    L.orc_quat_from_euler_f.argtypes = [C.c_float] * 3 + [f32p]
    L.orc_quat_from_euler_f.restype = None
    L.orc_look_at_rh_view_d.argtypes = [f64p, f64p]
    L.orc_look_at_rh_view_d.restype = None
    L.orc_quat_angle_to_d.argtypes = [f64p, f64p]
    L.orc_quat_angle_to_d.restype = C.c_double
    L.orc_camera_new_f.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.orc_camera_new_d.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.orc_camera_unproject_d.argtypes = [C.c_void_p, C.c_double, C.c_double, f64p, f64p]
    L.orc_camera_project_d.argtypes = [C.c_void_p, f64p, f64p, f64p]
    L.orc_camera_unproject_f.argtypes = [C.c_void_p, C.c_float, C.c_float, f32p, f32p]
    L.orc_camera_project_f.argtypes = [C.c_void_p, f32p, f32p, f32p]
    L.orc_camera_delta_f.argtypes = [C.c_void_p, C.c_float, C.c_float, f32p, f32p]
    L.orc_camera_delta_d.argtypes = [C.c_void_p, C.c_double, C.c_double, f64p, f64p]
    L.orc_camera_point_angle_f.argtypes = [C.c_void_p, C.c_float, C.c_float, f32p]
    L.orc_camera_point_angle_d.argtypes = [C.c_void_p, C.c_double, C.c_double, f64p]
    for name in ("orc_block_match", "orc_block_match_fast"):
        fn = getattr(L, name)
        fn.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                       C.POINTER(C.c_int16), C.POINTER(C.c_uint32), C.c_void_p, C.c_int]
        fn.restype = C.c_long
    L.orc_max_threads.restype = C.c_int
    L.orc_bgr_to_gray.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
    L.orc_bgr_to_gray.restype = None
    L.orc_bgr_to_rgba.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
    L.orc_bgr_to_rgba.restype = None
    L.orc_resize_linear.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
    L.orc_resize_linear.restype = None
    L.orc_mfield_size.argtypes = [C.c_size_t] * 6 + [szp, szp]
    L.orc_mfield_size.restype = None
    L.orc_contrast_mask.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.POINTER(C.c_int32)]
    L.orc_contrast_mask.restype = None
    L.orc_flow_entries.argtypes = [f32p, C.c_size_t, u8p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                   C.c_void_p, C.c_size_t]
    L.orc_flow_entries.restype = C.c_size_t


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def as_mv(entries) -> np.ndarray:
    """Accept an (n,4) float32 array or a structured MV array; return contiguous (n,4) f32."""
    a = np.asarray(entries)
    if a.dtype == MV_DTYPE:
        a = a.view("<f4").reshape(-1, 4)
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
    return a


# ------------------------------------------------------------------ densifier / detector
def densify(entries, w: int, h: int, return_counts: bool = False):
    mv = as_mv(entries)
    field = np.zeros((h, w, 2), np.float32)
    counts = np.zeros((h, w, 2), np.float32)
    lib().orc_densify(mv.ctypes.data, len(mv), w, h, _f32p(field), _f32p(counts))
    return (field, counts) if return_counts else field


def flow_field(entries, w: int, h: int) -> np.ndarray:
    mv = as_mv(entries)
    field = np.zeros((h, w, 2), np.float32)
    lib().orc_flow_field(mv.ctypes.data, len(mv), w, h, _f32p(field))
    return field


def interpolate_empty_cells(sums: np.ndarray, counts: np.ndarray):
    """interpolate_empty_cells (motion_field.rs:193-294) in place on sums / counts f32[h,w,2]."""
    assert sums.dtype == np.float32 and counts.dtype == np.float32 and sums.flags.c_contiguous and counts.flags.c_contiguous
    h, w = sums.shape[:2]
    lib().orc_interpolate_empty_cells(_f32p(sums), _f32p(counts), w, h)


def block_dim(min_size: float, subdivide: int) -> int:
    return int(lib().orc_block_dim(min_size, subdivide))


def detect_block_motion(entries, min_size=0.05, subdivide=3, target_motion=0.003, return_mean=False):
    """Returns (has_motion, area, dim, field[dim,dim,2]) (+ mean field)."""
    mv = as_mv(entries)
    dim = block_dim(min_size, subdivide)
    field = np.zeros((dim, dim, 2), np.float32)
    mean = np.zeros((dim, dim, 2), np.float32)
    area, dim_o = C.c_size_t(0), C.c_size_t(0)
    rc = lib().orc_detect_block_motion(mv.ctypes.data, len(mv), min_size, subdivide, target_motion,
                                       C.byref(area), C.byref(dim_o), _f32p(field), dim * dim, _f32p(mean))
    if rc < 0:
        raise ValueError("oracle: invalid detector geometry")
    out = (bool(rc), int(area.value), int(dim_o.value), field)
    return out + (mean,) if return_mean else out


# ------------------------------------------------------------------ almeida
def almeida_lsq_f32(entries, aspect: float, fov_y: float) -> np.ndarray:
    mv = as_mv(entries)
    q = np.zeros(4, np.float32)
    lib().orc_almeida_lsq_f(mv.ctypes.data, len(mv), aspect, fov_y, _f32p(q))
    return q


def almeida_lsq_f64(entries, aspect: float, fov_y: float) -> np.ndarray:
    mv = as_mv(entries)
    q = np.zeros(4, np.float64)
    lib().orc_almeida_lsq_d(mv.ctypes.data, len(mv), aspect, fov_y, _f64p(q))
    return q


def almeida_ransac_f32(entries, aspect, fov_y, num_iters=200, inlier_angle=0.05, num_samples=1000, seed=0):
    mv = as_mv(entries)
    q = np.zeros(4, np.float32)
    bc, bi = C.c_size_t(0), C.c_size_t(0)
    lib().orc_almeida_ransac_f(mv.ctypes.data, len(mv), aspect, fov_y, num_iters, inlier_angle, num_samples,
                               seed, _f32p(q), C.byref(bc), C.byref(bi))
    return q, int(bc.value), int(bi.value)


def perm_index(seed, it, stream, j, n) -> int:
    return int(lib().orc_perm_index(seed, it, stream, j, n))


class CameraF64:
    """StandardCamera in f64 (for building the reference's own test fields)."""

    def __init__(self, aspect: float, fov_y: float):
        self._buf = (C.c_double * 40)()
        lib().orc_camera_new_d(C.addressof(self._buf), aspect, fov_y)

    def unproject(self, x, y, inv_view):
        iv = np.ascontiguousarray(inv_view, np.float64)
        out = np.zeros(3, np.float64)
        lib().orc_camera_unproject_d(C.addressof(self._buf), x, y, _f64p(iv), _f64p(out))
        return out

    def project(self, world, view):
        v = np.ascontiguousarray(view, np.float64)
        w = np.ascontiguousarray(world, np.float64)
        out = np.zeros(2, np.float64)
        lib().orc_camera_project_d(C.addressof(self._buf), _f64p(w), _f64p(v), _f64p(out))
        return out

    def delta(self, x, y, rot):
        r = np.ascontiguousarray(rot, np.float64)
        out = np.zeros(2, np.float64)
        lib().orc_camera_delta_d(C.addressof(self._buf), x, y, _f64p(r), _f64p(out))
        return out

    def point_angle(self, x, y):
        out = np.zeros(2, np.float64)
        lib().orc_camera_point_angle_d(C.addressof(self._buf), x, y, _f64p(out))
        return out


class CameraF32:
    def __init__(self, aspect: float, fov_y: float):
        self._buf = (C.c_float * 40)()
        lib().orc_camera_new_f(C.addressof(self._buf), aspect, fov_y)

    def delta(self, x, y, rot):
        r = np.ascontiguousarray(rot, np.float32)
        out = np.zeros(2, np.float32)
        lib().orc_camera_delta_f(C.addressof(self._buf), x, y, _f32p(r), _f32p(out))
        return out

    def point_angle(self, x, y):
        out = np.zeros(2, np.float32)
        lib().orc_camera_point_angle_f(C.addressof(self._buf), x, y, _f32p(out))
        return out


def quat_from_euler(roll, pitch, yaw) -> np.ndarray:
    q = np.zeros(4, np.float64)
    lib().orc_quat_from_euler_d(roll, pitch, yaw, _f64p(q))
    return q


def calc_view(q) -> np.ndarray:
    """calc_view(rot, origin) of the reference test module (almeida:280-286)."""
    qq = np.ascontiguousarray(q, np.float64)
    view = np.zeros(16, np.float64)
    lib().orc_look_at_rh_view_d(_f64p(qq), _f64p(view))
    return view.reshape(4, 4)


def quat_angle_to(a, b) -> float:
    aa = np.ascontiguousarray(a, np.float64)
    bb = np.ascontiguousarray(b, np.float64)
    return float(lib().orc_quat_angle_to_d(_f64p(aa), _f64p(bb)))


# ------------------------------------------------------------------ block matcher
def block_match(prev: np.ndarray, cur: np.ndarray, block: int, rng: int, metric: int = 0,
                threads: int = 1, fast: bool = False):
    """Returns (mv_xy int16[nby,nbx,2], cost uint32[nby,nbx], entries f32[n,4])."""
    assert prev.dtype == np.uint8 and cur.dtype == np.uint8 and prev.shape == cur.shape
    prev = np.ascontiguousarray(prev)
    cur = np.ascontiguousarray(cur)
    h, w = prev.shape
    nbx, nby = w // block, h // block
    mv = np.zeros((nby, nbx, 2), np.int16)
    cost = np.zeros((nby, nbx), np.uint32)
    entries = np.zeros((nby * nbx, 4), np.float32)
    fn = lib().orc_block_match_fast if fast else lib().orc_block_match
    u8p = C.POINTER(C.c_uint8)
    rc = fn(prev.ctypes.data_as(u8p), cur.ctypes.data_as(u8p), w, h, w, block, rng, metric,
            mv.ctypes.data_as(C.POINTER(C.c_int16)), cost.ctypes.data_as(C.POINTER(C.c_uint32)),
            entries.ctypes.data, threads)
    if rc < 0:
        raise ValueError("oracle: invalid block-match arguments")
    return mv, cost, entries


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ------------------------------------------------------------------ cv-decoder dense-flow front end
def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def bgr_to_gray(img: np.ndarray, rgb_order: bool = False) -> np.ndarray:
    """cvtColor(COLOR_BGR2GRAY) (cv-decoder/src/lib.rs:138); img = u8[h,w,3|4]."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, ch = img.shape
    gray = np.zeros((h, w), np.uint8)
    lib().orc_bgr_to_gray(_u8p(img), w, h, w * ch, ch, int(rgb_order), _u8p(gray), w)
    return gray


def bgr_to_rgba(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w, ch = img.shape
    out = np.zeros((h, w, 4), np.uint8)
    lib().orc_bgr_to_rgba(_u8p(img), w, h, w * ch, ch, _u8p(out))
    return out


def resize_linear(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """resize(INTER_LINEAR), 8-bit (cv-decoder/src/lib.rs:127-135)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, ch = img.shape
    out = np.zeros((dh, dw, ch), np.uint8)
    lib().orc_resize_linear(_u8p(img), w, h, w * ch, ch, _u8p(out), dw, dh)
    return out


def mfield_size(frame_w, frame_h, ar_x=1, ar_y=1, max_w=150, max_h=150):
    dx, dy = C.c_size_t(0), C.c_size_t(0)
    lib().orc_mfield_size(frame_w, frame_h, ar_x, ar_y, max_w, max_h, C.byref(dx), C.byref(dy))
    return int(dx.value), int(dy.value)


def contrast_mask(gray: np.ndarray, return_sobel: bool = False):
    """Sobel(1,1,5) -> threshold 20 -> dilate 11x11 ellipse (cv-decoder/src/lib.rs:204-236); u8 0/255."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    mask = np.zeros((h, w), np.uint8)
    sob = np.zeros((h, w), np.int32)
    lib().orc_contrast_mask(_u8p(gray), w, h, w, _u8p(mask), sob.ctypes.data_as(C.POINTER(C.c_int32)))
    return (mask, sob) if return_sobel else mask


def flow_entries(flow: np.ndarray, mask=None, gw: int = 0, gh: int = 0) -> np.ndarray:
    """Dense flow f32[h,w,2] (+ u8 mask[h,w]) -> MotionEntry f32[n,4] (cv-decoder/src/lib.rs:238-291)."""
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    cap = max(w * h, gw * gh, 1)
    out = np.zeros((cap, 4), np.float32)
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
        mp = _u8p(mask)
    else:
        mp = None
    n = lib().orc_flow_entries(_f32p(flow), 2 * w, mp, w, w, h, gw, gh, out.ctypes.data, cap)
    return out[:n].copy()
