"""ctypes binding of ``libofps_b200.so`` (the C ABI in ``include/ofps_b200.h``).

This is the only way Python reaches the CUDA kernels: there is no CPU fallback and nothing here
imports ``oracle/``.  If the shared library is missing the import of :func:`lib` raises — build it
with ``python -m ofps_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libofps_b200.so")

OK = 0
E_INVALID, E_CUDA, E_NOMEM, E_NODEVICE, E_IO, E_CAPACITY = -1, -2, -3, -4, -5, -6
METRIC_SAD, METRIC_SSD = 0, 1

MV_DTYPE = np.dtype([("px", "<f4"), ("py", "<f4"), ("mx", "<f4"), ("my", "<f4")])

_u8p = C.POINTER(C.c_uint8)
_i16p = C.POINTER(C.c_int16)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_szp = C.POINTER(C.c_size_t)
_intp = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/ofps_b200.h declares
SIGNATURES = {
    "ofpsb_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ofpsb_destroy": (None, [_vp]),
    "ofpsb_last_error": (C.c_char_p, []),
    "ofpsb_version": (C.c_char_p, []),
    "ofpsb_set_stream": (C.c_int, [_vp, _vp]),
    "ofpsb_get_stream": (_vp, [_vp]),
    "ofpsb_sync": (C.c_int, [_vp]),
    "ofpsb_device_info": (C.c_int, [_vp, _intp, _szp, _szp, _intp, _intp]),
    "ofpsb_launch_count": (C.c_uint64, [_vp]),
    "ofpsb_set_option": (C.c_int, [_vp, C.c_char_p, C.c_longlong]),
    "ofpsb_block_match_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "ofpsb_block_match_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "ofpsb_stream_open": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ofpsb_stream_close": (None, [_vp]),
    "ofpsb_stream_blocks": (C.c_size_t, [_vp]),
    "ofpsb_stream_push": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _szp]),
    "ofpsb_stream_submit": (C.c_int, [_vp, _vp, C.c_size_t]),
    "ofpsb_stream_collect": (C.c_int, [_vp, _vp, _szp]),
    "ofpsb_tiled_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ofpsb_tiled_destroy": (None, [_vp]),
    "ofpsb_tiled_info": (C.c_int, [_vp, _intp, _intp, _intp, _intp, _intp, _intp]),
    "ofpsb_tiled_export": (C.c_int, [_vp, _vp]),
    "ofpsb_tiled_connect": (C.c_int, [_vp, _vp, _vp]),
    "ofpsb_tiled_connect_local": (C.c_int, [_vp, _vp, _vp]),
    "ofpsb_tiled_slot_ptr": (_vp, [_vp, C.c_int]),
    "ofpsb_tiled_upload": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t]),
    "ofpsb_tiled_publish": (C.c_int, [_vp, C.c_int]),
    "ofpsb_tiled_match": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
    "ofpsb_tiled_match_stream": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
    "ofpsb_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "ofpsb_host_free": (None, [_vp]),
    "ofpsb_dev_alloc": (C.c_int, [_vp, C.POINTER(_vp), C.c_size_t]),
    "ofpsb_dev_free": (None, [_vp, _vp]),
    "ofpsb_copy_to_device": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "ofpsb_copy_to_host": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "ofpsb_block_match": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    _vp, _vp, _vp, _szp]),
    "ofpsb_block_match_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                          C.c_int, C.c_int, _vp, _vp, _vp, _szp]),
    "ofpsb_block_match_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                        C.c_int, C.c_int, _vp, _vp, _vp]),
    "ofpsb_block_match_strip_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "ofpsb_block_match_strip_batch_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "ofpsb_densify": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp]),
    "ofpsb_densify_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp]),
    "ofpsb_flow_field": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp]),
    "ofpsb_interpolate_empty_cells": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t]),
    "ofpsb_block_dim": (C.c_int, [C.c_float, C.c_size_t, _szp]),
    "ofpsb_detect_block_motion": (C.c_int, [_vp, _vp, C.c_size_t, C.c_float, C.c_size_t, C.c_float, _intp, _szp,
                                            _szp, _vp, C.c_size_t]),
    "ofpsb_detect_block_motion_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_float, C.c_size_t, C.c_float, _intp, _szp,
                                                _szp, _vp, C.c_size_t]),
    "ofpsb_almeida": (C.c_int, [_vp, _vp, C.c_size_t, C.c_float, C.c_float, C.c_int, C.c_size_t, C.c_float,
                                C.c_size_t, C.c_uint64, _f32p]),
    "ofpsb_almeida_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_float, C.c_float, C.c_int, C.c_size_t, C.c_float,
                                    C.c_size_t, C.c_uint64, _f32p]),
    "ofpsb_frame_detect": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_size_t, C.c_float, _vp, _szp, _intp, _szp, _szp, _vp, C.c_size_t]),
    "ofpsb_mfield_size": (C.c_int, [C.c_size_t] * 6 + [_szp, _szp]),
    "ofpsb_frame_convert": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "ofpsb_frame_convert_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp]),
    "ofpsb_frame_resize": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int]),
    "ofpsb_frame_resize_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int]),
    "ofpsb_contrast_mask": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "ofpsb_contrast_mask_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int]),
    "ofpsb_flow_entries": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _vp, C.c_size_t, _szp]),
    "ofpsb_flow_entries_dev": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                         _vp, C.c_size_t, _szp]),
    "ofpsb_cv_flow_frame": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _vp,
                                      C.c_size_t, _szp]),
    "ofpsb_mvec_append": (C.c_int, [C.c_char_p, _vp, C.c_size_t, C.c_int]),
    "ofpsb_mvec_read": (C.c_int, [C.c_char_p, C.c_size_t, _vp, C.c_size_t, _szp]),
    "ofpsb_flo_write": (C.c_int, [C.c_char_p, _vp, C.c_size_t, C.c_size_t]),
}

_lib = None


class OfpsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libofps_b200 error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    """Load the shared library (raises if it has not been built — there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build the CUDA library with `python -m ofps_b200.build` "
                              "(ofps_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().ofpsb_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OK:
        raise OfpsError(rc, last_error())


def _ptr(a) -> int | None:
    """Host pointer of a numpy array / device pointer given as int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):   # torch tensor (device or pinned host): plumbing only
        return a.data_ptr()
    raise TypeError(f"cannot take a pointer of {type(a)!r}")


def as_mv(entries) -> np.ndarray:
    """(n,4) float32 view of MotionEntry data: (px, py, mx, my) per row."""
    a = np.asarray(entries)
    if a.dtype == MV_DTYPE:
        a = a.view("<f4").reshape(-1, 4)
    return np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)


class PinnedArray:
    """numpy view over page-locked host memory from ``ofpsb_host_alloc``."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = _vp()
        check(lib().ofpsb_host_alloc(C.byref(p), nbytes))
        self._p = p
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            lib().ofpsb_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One ``ofpsb_ctx``: a device, its stream and scratch memory.  Send, not Sync — like the
    reference's plugin objects (ofps/src/plugins/mod.rs:78-85)."""

    def __init__(self, device: int = 0):
        self._h = _vp()
        check(lib().ofpsb_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().ofpsb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- plumbing
    def sync(self):
        check(lib().ofpsb_sync(self._h))

    def set_stream(self, cuda_stream: int | None):
        check(lib().ofpsb_set_stream(self._h, cuda_stream))

    def get_stream(self) -> int:
        return lib().ofpsb_get_stream(self._h) or 0

    def set_option(self, key: str, value: int):
        check(lib().ofpsb_set_option(self._h, key.encode(), int(value)))

    def block_match_stats(self) -> dict:
        out = (C.c_uint64 * 4)()
        check(lib().ofpsb_block_match_stats(self._h, out))
        return {"blocks": out[0], "decided": out[1], "exact_evals": out[2], "worklist": out[3]}

    def block_match_kernel_ms(self):
        out = (C.c_float * 2)()
        check(lib().ofpsb_block_match_kernel_ms(self._h, out))
        return float(out[0]), float(out[1])

    def launch_count(self) -> int:
        return int(lib().ofpsb_launch_count(self._h))

    def device_info(self) -> dict:
        sm, cma, cmi = C.c_int(), C.c_int(), C.c_int()
        l2, mem = C.c_size_t(), C.c_size_t()
        check(lib().ofpsb_device_info(self._h, C.byref(sm), C.byref(l2), C.byref(mem), C.byref(cma), C.byref(cmi)))
        return {"sm_count": sm.value, "l2_bytes": l2.value, "mem_bytes": mem.value, "cc": (cma.value, cmi.value)}

    def dev_alloc(self, nbytes: int) -> int:
        p = _vp()
        check(lib().ofpsb_dev_alloc(self._h, C.byref(p), nbytes))
        return p.value

    def dev_free(self, ptr: int):
        lib().ofpsb_dev_free(self._h, ptr)

    def to_device(self, d_dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        check(lib().ofpsb_copy_to_device(self._h, d_dst, src.ctypes.data, src.nbytes))
        self.sync()

    def to_host(self, dst: np.ndarray, d_src: int):
        assert dst.flags.c_contiguous
        check(lib().ofpsb_copy_to_host(self._h, dst.ctypes.data, d_src, dst.nbytes))
        self.sync()

    # ---- block matcher (host buffers)
    def block_match(self, prev: np.ndarray, cur: np.ndarray, block: int, search: int, metric: int = METRIC_SAD,
                    want=("mv", "cost", "entries")):
        """prev/cur: uint8 [h,w] or [n,h,w].  Returns dict with mv int16[...,nby,nbx,2], cost uint32[...,nby,nbx],
        entries float32[..., nby*nbx, 4]."""
        if prev.dtype != np.uint8 or cur.dtype != np.uint8 or prev.shape != cur.shape:
            raise ValueError("prev/cur must be uint8 arrays of the same shape")
        single = prev.ndim == 2
        if single:
            prev, cur = prev[None], cur[None]
        if not (prev.flags.c_contiguous and cur.flags.c_contiguous):
            prev, cur = np.ascontiguousarray(prev), np.ascontiguousarray(cur)
        n, h, w = prev.shape
        nbx, nby = w // block, h // block
        mv = np.empty((n, nby, nbx, 2), np.int16) if "mv" in want else None
        cost = np.empty((n, nby, nbx), np.uint32) if "cost" in want else None
        ent = np.empty((n, nby * nbx, 4), np.float32) if "entries" in want else None
        nb = C.c_size_t()
        check(lib().ofpsb_block_match_batch(self._h, prev.ctypes.data, cur.ctypes.data, w, h, w, h * w, n, block,
                                            search, metric, _ptr(mv), _ptr(cost), _ptr(ent), C.byref(nb)))
        out = {"n_blocks": nb.value}
        for k, v in (("mv", mv), ("cost", cost), ("entries", ent)):
            if v is not None:
                out[k] = v[0] if single else v
        return out

    def block_match_raw(self, prev_ptr: int, cur_ptr: int, w: int, h: int, stride: int, pair_stride: int, n_pairs: int,
                        block: int, search: int, metric: int, mv_ptr, cost_ptr, entries_ptr) -> int:
        """Host-pointer batched entry point with explicit strides (pinned memory, stream mode)."""
        nb = C.c_size_t()
        check(lib().ofpsb_block_match_batch(self._h, prev_ptr, cur_ptr, w, h, stride, pair_stride, n_pairs, block,
                                            search, metric, _ptr(mv_ptr), _ptr(cost_ptr), _ptr(entries_ptr), C.byref(nb)))
        return nb.value

    # ---- block matcher (device buffers; asynchronous)
    def block_match_dev(self, d_prev, d_cur, w: int, h: int, stride: int, pair_stride: int, n_pairs: int, block: int,
                        search: int, metric: int, d_mv=None, d_cost=None, d_entries=None):
        check(lib().ofpsb_block_match_dev(self._h, _ptr(d_prev), _ptr(d_cur), w, h, stride, pair_stride, n_pairs, block,
                                          search, metric, _ptr(d_mv), _ptr(d_cost), _ptr(d_entries)))

    def block_match_strip_dev(self, d_prev, d_cur, w: int, strip_h: int, stride: int, halo_top: int, halo_bottom: int,
                              y_offset: int, full_h: int, block: int, search: int, metric: int, d_mv=None, d_cost=None,
                              d_entries=None):
        check(lib().ofpsb_block_match_strip_dev(self._h, _ptr(d_prev), _ptr(d_cur), w, strip_h, stride, halo_top,
                                                halo_bottom, y_offset, full_h, block, search, metric, _ptr(d_mv),
                                                _ptr(d_cost), _ptr(d_entries)))

    def block_match_strip_batch_dev(self, d_prev, d_cur, w: int, strip_h: int, stride: int, pair_stride: int, n_pairs: int,
                                    halo_top: int, halo_bottom: int, y_offset: int, full_h: int, block: int, search: int,
                                    metric: int, d_mv=None, d_cost=None, d_entries=None):
        check(lib().ofpsb_block_match_strip_batch_dev(self._h, _ptr(d_prev), _ptr(d_cur), w, strip_h, stride, pair_stride,
                                                      n_pairs, halo_top, halo_bottom, y_offset, full_h, block, search,
                                                      metric, _ptr(d_mv), _ptr(d_cost), _ptr(d_entries)))

    # ---- densifier / detector
    def densify(self, entries, gw: int, gh: int, return_counts: bool = False):
        mv = as_mv(entries)
        field = np.empty((gh, gw, 2), np.float32)
        counts = np.empty((gh, gw, 2), np.float32) if return_counts else None
        check(lib().ofpsb_densify(self._h, mv.ctypes.data, len(mv), gw, gh, field.ctypes.data, _ptr(counts)))
        return (field, counts) if return_counts else field

    def densify_dev(self, d_entries, n: int, gw: int, gh: int, d_field, d_counts=None):
        check(lib().ofpsb_densify_dev(self._h, _ptr(d_entries), n, gw, gh, _ptr(d_field), _ptr(d_counts)))

    def flow_field(self, entries, w: int, h: int) -> np.ndarray:
        """flow-extract's dense field: densify (GPU) -> interpolate_empty_cells (host, sequential) -> mean."""
        mv = as_mv(entries)
        field = np.empty((h, w, 2), np.float32)
        check(lib().ofpsb_flow_field(self._h, mv.ctypes.data, len(mv), w, h, field.ctypes.data))
        return field

    def detect_block_motion(self, entries, min_size: float = 0.05, subdivide: int = 3, target_motion: float = 0.003,
                            d_entries=None, n: int | None = None):
        """Returns (has_motion, area, dim, field[dim,dim,2]).  Pass ``d_entries``/``n`` for device-resident input."""
        dim = block_dim(min_size, subdivide)
        if dim == 0 or dim > 16384:
            raise OfpsError(E_INVALID, f"block_dim {dim} out of range")
        field = np.zeros((dim, dim, 2), np.float32)
        has, area, dim_o = C.c_int(), C.c_size_t(), C.c_size_t()
        if d_entries is not None:
            check(lib().ofpsb_detect_block_motion_dev(self._h, _ptr(d_entries), n, min_size, subdivide, target_motion,
                                                      C.byref(has), C.byref(area), C.byref(dim_o), field.ctypes.data,
                                                      dim * dim))
        else:
            mv = as_mv(entries)
            check(lib().ofpsb_detect_block_motion(self._h, mv.ctypes.data, len(mv), min_size, subdivide, target_motion,
                                                  C.byref(has), C.byref(area), C.byref(dim_o), field.ctypes.data,
                                                  dim * dim))
        return bool(has.value), int(area.value), int(dim_o.value), field

    def frame_detect(self, prev: np.ndarray, cur: np.ndarray, block: int, search: int, metric: int = METRIC_SAD,
                     min_size: float = 0.05, subdivide: int = 3, target_motion: float = 0.003):
        prev, cur = np.ascontiguousarray(prev), np.ascontiguousarray(cur)
        h, w = prev.shape
        dim = block_dim(min_size, subdivide)
        field = np.zeros((dim, dim, 2), np.float32)
        ent = np.empty(((h // block) * (w // block), 4), np.float32)
        has, area, dim_o, nb = C.c_int(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(lib().ofpsb_frame_detect(self._h, prev.ctypes.data, cur.ctypes.data, w, h, w, block, search, metric,
                                       min_size, subdivide, target_motion, ent.ctypes.data, C.byref(nb), C.byref(has),
                                       C.byref(area), C.byref(dim_o), field.ctypes.data, dim * dim))
        return ent, bool(has.value), int(area.value), int(dim_o.value), field

    # ---- estimator
    def almeida(self, entries, aspect: float, fov_y_deg: float, use_ransac: bool = False, num_iters: int = 200,
                inlier_angle_deg: float = 0.05, ransac_samples: int = 1000, seed: int = 0, d_entries=None,
                n: int | None = None) -> np.ndarray:
        q = np.zeros(4, np.float32)
        qp = q.ctypes.data_as(_f32p)
        if d_entries is not None:
            check(lib().ofpsb_almeida_dev(self._h, _ptr(d_entries), n, aspect, fov_y_deg, int(use_ransac), num_iters,
                                          inlier_angle_deg, ransac_samples, seed, qp))
        else:
            mv = as_mv(entries)
            check(lib().ofpsb_almeida(self._h, mv.ctypes.data, len(mv), aspect, fov_y_deg, int(use_ransac), num_iters,
                                      inlier_angle_deg, ransac_samples, seed, qp))
        return q


    # ---- cv-decoder dense-flow front end
    def frame_convert(self, img: np.ndarray, rgb_order: bool = False, want_gray: bool = True, want_rgba: bool = False):
        """u8[h,w,3|4] BGR(A) -> (gray u8[h,w] | None, rgba u8[h,w,4] | None)."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w, ch = img.shape
        gray = np.empty((h, w), np.uint8) if want_gray else None
        rgba = np.empty((h, w, 4), np.uint8) if want_rgba else None
        check(lib().ofpsb_frame_convert(self._h, img.ctypes.data, w, h, w * ch, ch, int(rgb_order), _ptr(gray), _ptr(rgba)))
        return gray, rgba

    def frame_convert_dev(self, d_src, w: int, h: int, stride: int, channels: int, rgb_order: bool, d_gray, gray_stride: int,
                          d_rgba=None):
        check(lib().ofpsb_frame_convert_dev(self._h, _ptr(d_src), w, h, stride, channels, int(rgb_order), _ptr(d_gray),
                                            gray_stride, _ptr(d_rgba)))

    def frame_resize(self, img: np.ndarray, dw: int, dh: int) -> np.ndarray:
        """resize(INTER_LINEAR) of u8[h,w,3|4] to (dh, dw) (reductions only)."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w, ch = img.shape
        out = np.empty((dh, dw, ch), np.uint8)
        check(lib().ofpsb_frame_resize(self._h, img.ctypes.data, w, h, w * ch, ch, out.ctypes.data, dw, dh))
        return out

    def contrast_mask(self, gray: np.ndarray) -> np.ndarray:
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        mask = np.empty((h, w), np.uint8)
        check(lib().ofpsb_contrast_mask(self._h, gray.ctypes.data, w, h, w, mask.ctypes.data))
        return mask

    def contrast_mask_dev(self, d_gray, w: int, h: int, stride: int, d_mask, mask_stride: int):
        check(lib().ofpsb_contrast_mask_dev(self._h, _ptr(d_gray), w, h, stride, _ptr(d_mask), mask_stride))

    def flow_entries(self, flow: np.ndarray, mask=None, gw: int = 0, gh: int = 0, cap: int | None = None) -> np.ndarray:
        """Dense flow f32[h,w,2] (+ u8 mask[h,w]) -> MotionEntry f32[n,4]."""
        flow = np.ascontiguousarray(flow, np.float32)
        h, w, _ = flow.shape
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        cap = (gw * gh if gw else w * h) if cap is None else cap
        out = np.empty((max(cap, 1), 4), np.float32)
        n = C.c_size_t()
        check(lib().ofpsb_flow_entries(self._h, flow.ctypes.data, _ptr(mask), w, h, gw, gh, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def flow_entries_dev(self, d_flow, flow_stride: int, d_mask, mask_stride: int, w: int, h: int, gw: int, gh: int,
                         d_entries, cap: int) -> int:
        n = C.c_size_t()
        check(lib().ofpsb_flow_entries_dev(self._h, _ptr(d_flow), flow_stride, _ptr(d_mask), mask_stride, w, h, gw, gh,
                                           _ptr(d_entries), cap, C.byref(n)))
        return int(n.value)

    def cv_flow_frame(self, gray, flow: np.ndarray, use_mask: bool = True, gw: int = 0, gh: int = 0) -> np.ndarray:
        """gray u8[h,w] + flow f32[h,w,2] -> MotionEntry f32[n,4]: mask, densifier and ordering on the GPU."""
        flow = np.ascontiguousarray(flow, np.float32)
        h, w, _ = flow.shape
        gray = np.ascontiguousarray(gray, np.uint8) if gray is not None else None
        cap = gw * gh if gw else w * h
        out = np.empty((max(cap, 1), 4), np.float32)
        n = C.c_size_t()
        check(lib().ofpsb_cv_flow_frame(self._h, _ptr(gray), w, flow.ctypes.data, w, h, int(use_mask), gw, gh,
                                        out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()


def mfield_size(frame_w: int, frame_h: int, ar_x: int = 1, ar_y: int = 1, max_w: int = 150, max_h: int = 150):
    dx, dy = C.c_size_t(), C.c_size_t()
    check(lib().ofpsb_mfield_size(frame_w, frame_h, ar_x, ar_y, max_w, max_h, C.byref(dx), C.byref(dy)))
    return int(dx.value), int(dy.value)


def interpolate_empty_cells(sums: np.ndarray, counts: np.ndarray):
    """In place on a densifier state, sums / counts f32[h,w,2] (host arithmetic, no device)."""
    assert sums.dtype == np.float32 and counts.dtype == np.float32 and sums.flags.c_contiguous and counts.flags.c_contiguous
    h, w = sums.shape[:2]
    check(lib().ofpsb_interpolate_empty_cells(sums.ctypes.data, counts.ctypes.data, w, h))


def block_dim(min_size: float, subdivide: int) -> int:
    d = C.c_size_t()
    check(lib().ofpsb_block_dim(min_size, subdivide, C.byref(d)))
    return int(d.value)


def version() -> str:
    return lib().ofpsb_version().decode()


# ---- interchange files (no GPU involved)
def mvec_append(path: str, entries, truncate: bool = False):
    mv = as_mv(entries)
    check(lib().ofpsb_mvec_append(os.fsencode(path), mv.ctypes.data, len(mv), int(truncate)))


def mvec_read(path: str, frame_index: int) -> np.ndarray:
    n = C.c_size_t()
    check(lib().ofpsb_mvec_read(os.fsencode(path), frame_index, None, 0, C.byref(n)))
    out = np.empty((n.value, 4), np.float32)
    check(lib().ofpsb_mvec_read(os.fsencode(path), frame_index, out.ctypes.data, n.value, C.byref(n)))
    return out


def flo_write(path: str, field: np.ndarray):
    f = np.ascontiguousarray(field, np.float32)
    h, w = f.shape[:2]
    check(lib().ofpsb_flo_write(os.fsencode(path), f.ctypes.data, w, h))


TILED_HANDLE_BYTES = 128


class Tiled:
    """One rank's strip of a spatially tiled frame (ofpsb_tiled_*, include/ofps_b200.h): halo rows of the previous frame
    are read straight from the neighbours' HBM — no exchange step."""

    def __init__(self, ctx: "Context", rank: int, world: int, w: int, h: int, block: int, search: int, n_slots: int = 2):
        self.ctx, self.rank, self.world = ctx, rank, world
        self._t = _vp()
        check(lib().ofpsb_tiled_create(ctx._h, rank, world, w, h, block, search, n_slots, C.byref(self._t)))
        v = [C.c_int() for _ in range(6)]
        check(lib().ofpsb_tiled_info(self._t, *[C.byref(x) for x in v]))
        self.y0, self.rows, self.own_rows, self.nbx, self.nby, self.stride = (x.value for x in v)
        self.n_blocks = self.nbx * self.nby

    def export(self) -> bytes:
        buf = C.create_string_buffer(TILED_HANDLE_BYTES)
        check(lib().ofpsb_tiled_export(self._t, buf))
        return buf.raw

    def connect(self, up: bytes | None, down: bytes | None):
        check(lib().ofpsb_tiled_connect(self._t, up, down))

    def connect_local(self, up: "Tiled | None", down: "Tiled | None"):
        check(lib().ofpsb_tiled_connect_local(self._t, up._t if up else None, down._t if down else None))

    def slot_ptr(self, slot: int) -> int:
        return int(lib().ofpsb_tiled_slot_ptr(self._t, slot) or 0)

    def upload(self, slot: int, own_rows: np.ndarray):
        """own_rows: this rank's rows of the frame, u8 [own_rows, w] (C-contiguous rows)."""
        a = np.ascontiguousarray(own_rows, np.uint8)
        assert a.shape[0] == self.own_rows
        check(lib().ofpsb_tiled_upload(self._t, slot, a.ctypes.data, a.strides[0]))
        self.ctx.sync()          # `a` may be a temporary

    def publish(self, slot: int):
        check(lib().ofpsb_tiled_publish(self._t, slot))

    def match(self, prev_slot: int, cur_slot: int, d_entries: int = 0, d_mv: int = 0, d_cost: int = 0, wait: bool = True):
        check(lib().ofpsb_tiled_match(self._t, prev_slot, cur_slot, d_entries or None, d_mv or None, d_cost or None, int(wait)))

    def match_stream(self, first_slot: int, n_pairs: int, d_entries: int = 0, d_mv: int = 0, d_cost: int = 0, wait: bool = True):
        check(lib().ofpsb_tiled_match_stream(self._t, first_slot, n_pairs, d_entries or None, d_mv or None, d_cost or None,
                                             int(wait)))

    def close(self):
        if self._t:
            lib().ofpsb_tiled_destroy(self._t)
            self._t = _vp()


class FrameStream:
    """Streaming block-matching decoder (ofpsb_stream_*): frames in one at a time, MotionEntry lists out."""

    def __init__(self, ctx: "Context", w: int, h: int, block: int, search: int, metric: int = 0, depth: int = 4):
        self.ctx = ctx
        self._s = _vp()
        check(lib().ofpsb_stream_open(ctx._h, w, h, block, search, metric, depth, C.byref(self._s)))
        self.n_blocks = int(lib().ofpsb_stream_blocks(self._s))

    def push(self, frame: np.ndarray, out: np.ndarray | None = None):
        """-> entries [n_blocks, 4] f32 of (previous frame, frame), or None for the first frame."""
        out = np.empty((self.n_blocks, 4), np.float32) if out is None else out
        n = C.c_size_t()
        check(lib().ofpsb_stream_push(self._s, frame.ctypes.data, frame.strides[0], out.ctypes.data, C.byref(n)))
        return out if n.value else None

    def submit(self, frame: np.ndarray):
        check(lib().ofpsb_stream_submit(self._s, frame.ctypes.data, frame.strides[0]))

    def collect(self, out: np.ndarray | None = None):
        out = np.empty((self.n_blocks, 4), np.float32) if out is None else out
        n = C.c_size_t()
        check(lib().ofpsb_stream_collect(self._s, out.ctypes.data, C.byref(n)))
        return out if n.value else None

    def close(self):
        if self._s:
            lib().ofpsb_stream_close(self._s)
            self._s = _vp()
