"""CPU: the host-side hole fill of the C ABI (ofpsb_interpolate_empty_cells, csrc/hole_fill.cu — seven buckets of
hierarchical bitmaps) against the oracle's restatement of MotionFieldDensifier::interpolate_empty_cells
(ofps/src/motion_field.rs:193-294 — lazy binary heap on the reference's key) and against a literal pure-Python
version with a sorted set.  Three independent implementations of one sequential definition, bit for bit."""
import numpy as np
import pytest

from ofps_b200 import capi

EPS = np.finfo(np.float32).eps


def make_state(w, h, fill, seed):
    rng = np.random.default_rng(seed)
    hit = rng.random((h, w)) < fill
    n = rng.integers(1, 5, (h, w)).astype(np.float32)
    counts = np.where(hit, n, np.float32(EPS)).astype(np.float32)
    counts = np.repeat(counts[..., None], 2, axis=2).copy()
    sums = (np.where(hit[..., None], rng.standard_normal((h, w, 2)) * 0.01, 0.0)).astype(np.float32)
    return sums, counts


def py_interpolate(sums, counts):
    """Literal restatement with a sorted list as the BTreeSet (small fields only)."""
    import bisect
    F = np.float32
    h, w = sums.shape[:2]
    s = sums.reshape(-1, 2)
    c = counts.reshape(-1, 2)
    nb = [(-1, 0), (0, -1), (-1, -1), (1, 0), (0, 1), (1, 1)]

    def calc(i):
        x, y = i % w, i // w
        return sum(1 for ox, oy in nb if 0 <= x + ox < w and 0 <= y + oy < h and c[(x + ox) + (y + oy) * w, 0] > F(0.1))

    q = sorted((-calc(i), i) for i in range(w * h) if c[i, 0] < F(0.5))
    if len(q) == w * h:
        return
    while q:
        key, i = q.pop(0)
        x, y = i % w, i // w
        added = False
        for ox, oy in nb:
            if 0 <= x + ox < w and 0 <= y + oy < h:
                j = (x + ox) + (y + oy) * w
                cnt = c[j, 0]
                if cnt > F(0.1):
                    scale = F(F(1) - F(F(np.sqrt(F(ox * ox + oy * oy))) * F(0.5)))
                    f = F(scale * F(F(1) / cnt))
                    v = (F(f * s[j, 0]), F(f * s[j, 1]))
                    c[i, 0] = F(c[i, 0] + scale)
                    c[i, 1] = F(c[i, 1] + scale)
                    s[i, 0] = F(F(v[0] * scale) + s[i, 0])
                    s[i, 1] = F(F(v[1] * scale) + s[i, 1])
                    added = True
        assert added
        for ox, oy in nb:
            if 0 <= x + ox < w and 0 <= y + oy < h:
                j = (x + ox) + (y + oy) * w
                old = (-calc(j) + 1, j)
                k = bisect.bisect_left(q, old)
                if k < len(q) and q[k] == old:
                    q.pop(k)
                    bisect.insort(q, (old[0] - 1, j))


@pytest.mark.parametrize("w,h,fill,seed", [(9, 7, 0.2, 1), (16, 16, 0.05, 2), (1, 12, 0.3, 3), (13, 1, 0.3, 4), (20, 15, 0.6, 5),
                                          (6, 6, 0.03, 6), (3, 3, 1.0, 7), (24, 10, 0.01, 11)])
def test_three_implementations_agree(oracle, w, h, fill, seed):
    sums, counts = make_state(w, h, fill, seed)
    if not (counts[..., 0] >= 0.5).any():
        counts[h // 2, w // 2] = 1.0
        sums[h // 2, w // 2] = (0.01, -0.02)
    s1, c1 = sums.copy(), counts.copy()
    s2, c2 = sums.copy(), counts.copy()
    s3, c3 = sums.copy(), counts.copy()
    capi.interpolate_empty_cells(s1, c1)
    oracle.interpolate_empty_cells(s2, c2)
    py_interpolate(s3, c3)
    assert s1.tobytes() == s2.tobytes() == s3.tobytes()
    assert c1.tobytes() == c2.tobytes() == c3.tobytes()
    assert (c1[..., 0] > 0.1).all()   # every cell ends up filled


@pytest.mark.parametrize("w,h,fill,seed", [(150, 84, 0.1, 1), (640, 360, 0.004, 2), (1920, 1080, 0.004, 3), (300, 200, 0.9, 4),
                                          (4100, 3, 0.01, 5)])
def test_matches_oracle_large(oracle, w, h, fill, seed):
    sums, counts = make_state(w, h, fill, seed)
    s1, c1 = sums.copy(), counts.copy()
    s2, c2 = sums.copy(), counts.copy()
    capi.interpolate_empty_cells(s1, c1)
    oracle.interpolate_empty_cells(s2, c2)
    assert s1.tobytes() == s2.tobytes() and c1.tobytes() == c2.tobytes()


def test_all_empty_and_all_full_are_untouched(oracle):
    sums, counts = make_state(12, 9, 0.0, 1)
    s, c = sums.copy(), counts.copy()
    capi.interpolate_empty_cells(s, c)
    assert s.tobytes() == sums.tobytes() and c.tobytes() == counts.tobytes()   # early-out (:243-245)
    sums, counts = make_state(12, 9, 1.1, 2)
    s, c = sums.copy(), counts.copy()
    capi.interpolate_empty_cells(s, c)
    assert s.tobytes() == sums.tobytes() and c.tobytes() == counts.tobytes()
    with pytest.raises(capi.OfpsError):
        capi.check(capi.lib().ofpsb_interpolate_empty_cells(None, None, 3, 3))
