"""Deterministic synthetic inputs — language-independent recipes.

* luma frames: ``n(x,y) = splitmix64(seed ^ (y*W+x)) >> 56`` then a 3x3 integer box
  blur ``(sum+4)//9`` with edge replication;
* frame pairs: ``cur`` = ``prev`` panned by an integer global motion plus K rectangles
  moved by their own integer motions, uncovered pixels from a second noise field,
  optional +-noise LSBs;
* estimator fields: one entry per grid point, motion generated the way the reference's
  own test does (almeida-estimator/src/lib.rs:257-306): un-project at the identity view,
  re-project under ``calc_view(q_truth)``, in f64, rounded to f32.

Pure numpy; no dependency on the CUDA library or on ``oracle/``.
"""
from __future__ import annotations

import math

import numpy as np

_U64 = np.uint64
BASE_SEED = 0x0F950001


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser (wrap-around uint64 arithmetic)."""
    with np.errstate(over="ignore"):
        x = x.astype(_U64) + _U64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> _U64(27))) * _U64(0x94D049BB133111EB)
        return x ^ (x >> _U64(31))


def splitmix64_scalar(x: int) -> int:
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
    return x ^ (x >> 31)


def noise_plane(seed: int, w: int, h: int) -> np.ndarray:
    idx = np.arange(w * h, dtype=_U64)
    return (splitmix64(idx ^ _U64(seed & ((1 << 64) - 1))) >> _U64(56)).astype(np.uint8).reshape(h, w)


def box_blur3(img: np.ndarray) -> np.ndarray:
    p = np.pad(img.astype(np.uint32), 1, mode="edge")
    h, w = img.shape
    acc = np.zeros((h, w), np.uint32)
    for dy in range(3):
        for dx in range(3):
            acc += p[dy:dy + h, dx:dx + w]
    return ((acc + 4) // 9).astype(np.uint8)


def textured_plane(seed: int, w: int, h: int) -> np.ndarray:
    return box_blur3(noise_plane(seed, w, h))


class _Rng:
    """Tiny counter RNG (splitmix64 stream) so recipes are reproducible anywhere."""

    def __init__(self, seed: int):
        self.s = seed & ((1 << 64) - 1)

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
        return splitmix64_scalar(self.s)

    def randint(self, lo: int, hi: int) -> int:
        """uniform integer in [lo, hi]"""
        return lo + self.next() % (hi - lo + 1)


def shift_plane(src: np.ndarray, fill: np.ndarray, gx: int, gy: int) -> np.ndarray:
    """out[y,x] = src[y-gy, x-gx] where defined, else fill[y,x] (content moves by (gx,gy))."""
    h, w = src.shape
    out = fill.copy()
    xs0, xs1 = max(0, gx), min(w, w + gx)
    ys0, ys1 = max(0, gy), min(h, h + gy)
    if xs1 > xs0 and ys1 > ys0:
        out[ys0:ys1, xs0:xs1] = src[ys0 - gy:ys1 - gy, xs0 - gx:xs1 - gx]
    return out


def make_pair(w: int, h: int, search: int, index: int = 0, n_rects: int = 8, noise_lsb: int = 0,
              seed: int = BASE_SEED):
    """One synthetic (prev, cur) luma pair plus its ground-truth motions.

    Content moves prev->cur by ``(gx, gy)`` globally (each in [-search/2, search/2]) and by a
    per-rectangle motion in [-search, search] inside K rectangles (64..256 px).  The block
    matcher's convention (cur block at p matches prev at p+d) therefore reports
    ``d = -(motion)`` for blocks fully inside a moved region.
    """
    rng = _Rng(seed + 1000003 * index)
    s_prev = rng.next()
    s_fill = rng.next()
    prev = textured_plane(s_prev, w, h)
    fill = textured_plane(s_fill, w, h)
    half = max(search // 2, 0)
    gx, gy = rng.randint(-half, half), rng.randint(-half, half)
    cur = shift_plane(prev, fill, gx, gy)
    rects = []
    for _ in range(n_rects):
        rw = rng.randint(min(64, w // 2), min(256, w // 2))
        rh = rng.randint(min(64, h // 2), min(256, h // 2))
        x0 = rng.randint(0, w - rw)
        y0 = rng.randint(0, h - rh)
        dx, dy = rng.randint(-search, search), rng.randint(-search, search)
        # destination pixels x in the rectangle whose source x-dx lies inside prev
        cx0, cx1 = max(x0, dx), min(x0 + rw, w + dx)
        cy0, cy1 = max(y0, dy), min(y0 + rh, h + dy)
        if cx1 > cx0 and cy1 > cy0:
            cur[cy0:cy1, cx0:cx1] = prev[cy0 - dy:cy1 - dy, cx0 - dx:cx1 - dx]
            rects.append((cx0, cy0, cx1 - cx0, cy1 - cy0, dx, dy))
    if noise_lsb > 0:
        nz = noise_plane(rng.next(), w, h).astype(np.int16) % (2 * noise_lsb + 1) - noise_lsb
        cur = np.clip(cur.astype(np.int16) + nz, 0, 255).astype(np.uint8)
    return prev, cur, {"global": (gx, gy), "rects": rects}


def make_batch(n_pairs: int, w: int, h: int, search: int, noise_lsb: int = 0, first_index: int = 0,
               seed: int = BASE_SEED):
    """``n_pairs`` independent pairs as two arrays [n, h, w] uint8 (prev, cur)."""
    prev = np.empty((n_pairs, h, w), np.uint8)
    cur = np.empty((n_pairs, h, w), np.uint8)
    for i in range(n_pairs):
        p, c, _ = make_pair(w, h, search, first_index + i, noise_lsb=noise_lsb, seed=seed)
        prev[i], cur[i] = p, c
    return prev, cur


def make_stream(n_frames: int, w: int, h: int, search: int, n_rects: int = 8, seed: int = BASE_SEED,
                first_index: int = 0, noise_lsb: int = 0) -> np.ndarray:
    """``n_frames`` consecutive luma frames [n, h, w] of one synthetic video: frame i+1 is frame i
    panned by an integer global motion with ``n_rects`` rectangles moved on their own (the
    :func:`make_pair` recipe, chained), so pair (i, i+1) is a real motion pair for every i.
    ``noise_lsb`` > 0 adds independent uniform sensor noise in [-noise_lsb, noise_lsb] to every frame
    (the clean content is what is chained, so the noise does not accumulate): no pair has an exact match."""
    rng = _Rng(seed + 7919 * (first_index + 1))
    frames = np.empty((n_frames, h, w), np.uint8)
    frames[0] = textured_plane(rng.next(), w, h)
    half = max(search // 2, 0)
    for i in range(1, n_frames):
        prev = frames[i - 1]
        fill = textured_plane(rng.next(), w, h)
        gx, gy = rng.randint(-half, half), rng.randint(-half, half)
        cur = shift_plane(prev, fill, gx, gy)
        for _ in range(n_rects):
            rw = rng.randint(min(64, w // 2), min(256, w // 2))
            rh = rng.randint(min(64, h // 2), min(256, h // 2))
            x0 = rng.randint(0, w - rw)
            y0 = rng.randint(0, h - rh)
            dx, dy = rng.randint(-search, search), rng.randint(-search, search)
            cx0, cx1 = max(x0, dx), min(x0 + rw, w + dx)
            cy0, cy1 = max(y0, dy), min(y0 + rh, h + dy)
            if cx1 > cx0 and cy1 > cy0:
                cur[cy0:cy1, cx0:cx1] = prev[cy0 - dy:cy1 - dy, cx0 - dx:cx1 - dx]
        frames[i] = cur
    if noise_lsb > 0:
        nrng = _Rng(seed ^ 0x5EED0000 ^ (first_index + 1))
        for i in range(n_frames):
            nz = noise_plane(nrng.next(), w, h).astype(np.int16) % (2 * noise_lsb + 1) - noise_lsb
            frames[i] = np.clip(frames[i].astype(np.int16) + nz, 0, 255).astype(np.uint8)
    return frames


# ----------------------------------------------------------------------------- estimator fields
def _quat_from_euler(roll: float, pitch: float, yaw: float) -> np.ndarray:
    sr, cr = math.sin(roll * 0.5), math.cos(roll * 0.5)
    sp, cp = math.sin(pitch * 0.5), math.cos(pitch * 0.5)
    sy, cy = math.sin(yaw * 0.5), math.cos(yaw * 0.5)
    return np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])


def _quat_to_mat3(q: np.ndarray) -> np.ndarray:
    w, i, j, k = q
    return np.array([
        [w * w + i * i - j * j - k * k, 2 * (i * j - w * k), 2 * (w * j + i * k)],
        [2 * (w * k + i * j), w * w - i * i + j * j - k * k, 2 * (j * k - w * i)],
        [2 * (i * k - w * j), 2 * (w * i + j * k), w * w - i * i - j * j + k * k]])


def _look_at_rh_rot(q: np.ndarray) -> np.ndarray:
    """Rotation part of the reference test's calc_view(q, origin) (almeida:280-286)."""
    r = _quat_to_mat3(q)
    d = r @ np.array([0.0, -1.0, 0.0])
    up = r @ np.array([0.0, 0.0, 1.0])
    z = -d / np.linalg.norm(d)
    x = np.cross(up, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    y /= np.linalg.norm(y)
    return np.stack([x, y, z])


def rotation_field(width: int, height: int, aspect: float, fov_y_deg: float, euler_deg,
                   centre_offset: float = 0.5) -> tuple[np.ndarray, np.ndarray]:
    """Dense (width*height) MotionEntry field for a pure camera rotation.

    ``pos = ((x+centre_offset)/width, (y+centre_offset)/height)`` (cv-decoder's convention,
    cv-decoder/src/lib.rs:264-269, with the default 0.5); motion = p2 - p1 exactly as the reference
    test builds it, evaluated in f64 and rounded to f32.  Returns (entries[n,4] f32, q_truth[4] f64).
    """
    q = _quat_from_euler(*(math.radians(a) for a in euler_deg))
    t = math.tan(math.radians(fov_y_deg) / 2)
    m11 = 1.0 / t
    m00 = m11 / aspect
    zn, zf = 0.1, 10.0
    m22 = (zf + zn) / (zn - zf)
    m23 = 2 * zf * zn / (zn - zf)
    xs = (np.arange(width, dtype=np.float64) + centre_offset) / width
    ys = (np.arange(height, dtype=np.float64) + centre_offset) / height
    px, py = np.meshgrid(xs, ys)
    px, py = px.ravel(), py.ravel()
    # unproject at NDC z = 1 through inv_proj, then "inv_view" = calc_view(identity) (sic, almeida:272-276)
    cx, cy = px * 2 - 1, py * 2 - 1
    hx, hy, hz = cx / m00, cy / m11, -np.ones_like(cx)
    hw = 1.0 / m23 + (m22 / m23)
    v0 = _look_at_rh_rot(np.array([1.0, 0, 0, 0]))
    pts = np.stack([hx / hw, hy / hw, hz / hw], axis=1) @ v0.T

    def project(points, view):
        v = points @ view.T
        inv = -1.0 / v[:, 2]
        sx, sy, sz = m00 * v[:, 0] * inv, m11 * v[:, 1] * inv, (m22 * v[:, 2] + m23) * inv
        return np.stack([(sx / sz + 1) * 0.5, (sy / sz + 1) * 0.5], axis=1)

    p1 = project(pts, v0)
    p2 = project(pts, _look_at_rh_rot(q))
    entries = np.concatenate([p1, p2 - p1], axis=1).astype(np.float32)
    return entries, q


def corrupt_field(entries: np.ndarray, fraction: float, seed: int = 7, magnitude: float = 0.05) -> np.ndarray:
    """Replace a PRNG-chosen ``fraction`` of the motions by uniform vectors in [-magnitude, magnitude]^2."""
    n = len(entries)
    r = splitmix64(np.arange(n, dtype=_U64) ^ _U64(seed))
    pick = (r >> _U64(40)).astype(np.float64) / float(1 << 24) < fraction
    out = entries.copy()
    u = splitmix64(r)
    v = splitmix64(u)
    ux = ((u >> _U64(40)).astype(np.float64) / float(1 << 24) * 2 - 1) * magnitude
    vy = ((v >> _U64(40)).astype(np.float64) / float(1 << 24) * 2 - 1) * magnitude
    out[pick, 2] = ux[pick].astype(np.float32)
    out[pick, 3] = vy[pick].astype(np.float32)
    return out
