"""Second, independent restatement of the reference path in pure Python / numpy float32 scalars.

Used to (a) cross-check the C oracle and (b) generate the committed golden vectors under
tests/golden/ (see make_golden.py).  It follows the reference sources directly, not the C oracle:
  ofps/src/motion_field.rs:133-190, 297-308      densifier
  block-motion-detector/src/lib.rs:49-119        detector (stack flood fill, literally)
  SURVEY.md §8c                                  block-matching specification (ours)
Slow by design: small cases only."""
import math

import numpy as np

F = np.float32
EPS = np.finfo(np.float32).eps


def _round_half_away(v):
    v = float(v)
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)


def _as_usize(v):
    if v != v or v <= 0:
        return 0
    return int(v)


def clamp_point(x, y):
    # nalgebra::clamp on Point2: `if val > min { if val < max { val } else { max } } else { min }`
    # with the all-components partial order of matrices
    x, y = F(x), F(y)
    if x > 0 and y > 0:
        if x < 1 and y < 1:
            return x, y
        return F(1), F(1)
    return F(0), F(0)


class Densifier:
    def __init__(self, w, h):
        self.w, self.h = w, h
        self.sums = np.zeros((w * h, 2), np.float32)
        self.counts = np.full((w * h, 2), EPS, np.float32)

    def add_vector(self, px, py, mx, my, weight=F(1)):
        px, py = clamp_point(px, py)
        x = _as_usize(_round_half_away(F(px * F(self.w - 1))))
        y = _as_usize(_round_half_away(F(py * F(self.h - 1))))
        idx = y * self.w + x
        self.counts[idx, 0] = F(self.counts[idx, 0] + weight)
        self.counts[idx, 1] = F(self.counts[idx, 1] + weight)
        self.sums[idx, 0] = F(F(F(mx) * weight) + self.sums[idx, 0])
        self.sums[idx, 1] = F(F(F(my) * weight) + self.sums[idx, 1])
        return x, y

    def field(self):
        return (self.sums / self.counts).astype(np.float32).reshape(self.h, self.w, 2)


def densify(entries, w, h):
    d = Densifier(w, h)
    for px, py, mx, my in np.asarray(entries, np.float32).reshape(-1, 4):
        d.add_vector(px, py, mx, my)
    return d.field(), d.counts.reshape(h, w, 2).copy()


def block_dim(min_size, subdivide):
    bw = F(F(math.sqrt(float(F(min_size)))) / F(subdivide))   # f32 sqrt is correctly rounded
    return _as_usize(math.ceil(float(F(F(1) / bw))))


def detect_block_motion(entries, min_size=0.05, subdivide=3, target_motion=0.003):
    dim = block_dim(min_size, subdivide)
    mf, _ = densify(entries, dim, dim)
    mp = [[False] * dim for _ in range(dim)]
    for y in range(dim):
        for x in range(dim):
            mx, my = mf[y, x]
            mag = F(math.sqrt(float(F(F(mx * mx) + F(my * my)))))
            if mag >= F(target_motion):
                mp[y][x] = True
    biggest, best = 0, None
    for y in range(dim):
        for x in range(dim):
            if not mp[y][x]:
                continue
            area = 0
            mf2 = np.zeros((dim, dim, 2), np.float32)
            mp[y][x] = False
            stack = [(x, y)]
            while stack:
                cx, cy = stack.pop()
                area += 1
                for ox in (-1, 0, 1):
                    for oy in (-1, 0, 1):
                        nx, ny = cx + ox, cy + oy
                        if 0 <= nx < dim and 0 <= ny < dim and mp[ny][nx]:
                            mf2[ny, nx] = mf[ny, nx]
                            stack.append((nx, ny))
                            mp[ny][nx] = False
            if area > biggest:
                biggest, best = area, mf2
    if best is not None and F(F(biggest) / F(dim * dim)) >= F(min_size):
        return True, biggest, dim, best
    return False, 0, dim, np.zeros((dim, dim, 2), np.float32)


def block_match(prev, cur, block, search, metric=0):
    h, w = prev.shape
    nbx, nby = w // block, h // block
    mv = np.zeros((nby, nbx, 2), np.int16)
    cost = np.zeros((nby, nbx), np.uint32)
    ent = np.zeros((nby * nbx, 4), np.float32)
    p32, c32 = prev.astype(np.int64), cur.astype(np.int64)
    nx, ny = F(1) / F(w), F(1) / F(h)
    for by in range(nby):
        for bx in range(nbx):
            x0, y0 = bx * block, by * block
            c = c32[y0:y0 + block, x0:x0 + block]
            best = None
            for dy in range(-search, search + 1):
                if y0 + dy < 0 or y0 + dy + block > h:
                    continue
                for dx in range(-search, search + 1):
                    if x0 + dx < 0 or x0 + dx + block > w:
                        continue
                    d = c - p32[y0 + dy:y0 + dy + block, x0 + dx:x0 + dx + block]
                    v = int(np.abs(d).sum()) if metric == 0 else int((d * d).sum())
                    key = (v, dx * dx + dy * dy, dy, dx)
                    if best is None or key < best:
                        best = key
            v, _, dy, dx = best
            mv[by, bx] = (dx, dy)
            cost[by, bx] = v
            sx, sy = x0 + block // 2 + dx, y0 + block // 2 + dy
            ent[by * nbx + bx] = (F(sx) * nx, F(sy) * ny, F(dx) * -nx, F(dy) * -ny)
    return mv, cost, ent


# ---------------------------------------------------------------- cv-decoder dense-flow front end
def mfield_size(frame_w, frame_h, ar_x, ar_y, max_w, max_h):
    """cv-decoder/src/lib.rs:90-118 (usize arithmetic)."""
    ratio = (frame_w * ar_x, frame_h * ar_y)
    w, h = min(max_w, frame_w), min(max_h, frame_h)
    width_based = (w, w * ratio[1] // ratio[0])
    height_based = (h * ratio[0] // ratio[1], h)
    return width_based if width_based[0] < height_based[0] else height_based


def flow_entries(flow, mask=None, gw=0, gh=0):
    """cv-decoder/src/lib.rs:238-291, literally: raster loop, optional mask test, per-pixel push or
    down-sampling densifier + BTreeSet<(x,y)> of touched cells."""
    h, w, _ = flow.shape
    nx, ny = F(F(1) / F(w)), F(F(1) / F(h))
    out = []
    dens = Densifier(gw, gh) if gw else None
    points = set()
    for y in range(h):
        for x in range(w):
            if mask is not None and mask[y, x] < 0.1:
                continue
            pos = (F(F(F(x) + F(0.5)) * nx), F(F(F(y) + F(0.5)) * ny))
            motion = (F(flow[y, x, 0] * nx), F(flow[y, x, 1] * ny))
            if dens is not None:
                points.add(dens.add_vector(pos[0], pos[1], motion[0], motion[1]))
            else:
                out.append((pos[0], pos[1], motion[0], motion[1]))
    if dens is not None:
        field = dens.field()
        gx, gy = F(F(1) / F(gw)), F(F(1) / F(gh))
        for (x, y) in sorted(points):
            out.append((F(F(F(x) + F(0.5)) * gx), F(F(F(y) + F(0.5)) * gy), field[y, x, 0], field[y, x, 1]))
    return np.array(out, np.float32).reshape(-1, 4)
